# full ncu captures of the comb-round kernels inside one R1CS prove at 2^16 multipliers (folded generators, n_eff = 2048)
# and of the materialisation kernel
ncu --set full --clock-control none --import-source on -k regex:k_comb_round -s 12 -c 1 -o gpurun_out/r2g_comb_round -f python tools/prove_profile.py 16 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_comb_final -s 12 -c 1 -o gpurun_out/r2g_comb_final -f python tools/prove_profile.py 16 0 > /dev/null 2>&1
ls -la gpurun_out/r2g*.ncu-rep
