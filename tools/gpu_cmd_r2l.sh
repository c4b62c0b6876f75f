python tools/profile_run.py 64 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 70 -c 1 -o gpurun_out/r2_apply -f python tools/profile_run.py 64 1 > gpurun_out/ncu_apply.log 2>&1
tail -3 gpurun_out/ncu_apply.log; ls -la gpurun_out/*.ncu-rep
