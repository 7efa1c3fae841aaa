# full ncu captures of the comb-round kernels inside one R1CS prove at 2^10 multipliers
ncu --set full --clock-control none --import-source on -k regex:k_comb_round -s 12 -c 1 -o gpurun_out/r2e_comb_round -f python tools/prove_profile.py 10 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_comb_final -s 12 -c 1 -o gpurun_out/r2e_comb_final -f python tools/prove_profile.py 10 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
