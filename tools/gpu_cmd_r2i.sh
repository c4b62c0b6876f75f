timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -32
B200LDLT_TIMELINE=1 timeout 300 python tools/profile_run.py 64 2 > gpurun_out/r2i_timeline.txt 2>&1; grep -c timeline gpurun_out/r2i_timeline.txt
