ncu --set full --clock-control none --import-source on -k regex:k_front --launch-skip 13 --launch-count 1 -o gpurun_out/r3_k_front -f python tools/profile_run.py 64 1 > gpurun_out/r3b_ncu.log 2>&1
tail -2 gpurun_out/r3b_ncu.log
