ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_launches.csv python tools/profile_run.py 64 1 > gpurun_out/r2x_ncu.log 2>&1
tail -3 gpurun_out/r2x_ncu.log
