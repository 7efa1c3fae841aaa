# full ncu capture of the materialisation kernel inside one R1CS prove at 2^16 multipliers
ncu --set full --clock-control none --import-source on -k regex:k_comb_materialize -s 2 -c 1 -o gpurun_out/r2f_materialize -f python tools/prove_profile.py 16 0 > /dev/null 2>&1
ls -la gpurun_out/r2f*.ncu-rep
