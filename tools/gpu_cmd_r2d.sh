timeout 300 python tools/dbg_defer.py 2>&1 | tail -20
B200LDLT_TIMELINE=1 timeout 300 python tools/profile_run.py 64 2 > gpurun_out/r2d_timeline.txt 2>&1; grep -c timeline gpurun_out/r2d_timeline.txt
