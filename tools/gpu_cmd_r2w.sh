timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2w_pytest.log; cat gpurun_out/r2w_pytest.log
timeout 600 python tools/front_check.py 2>&1 | tail -4
for cfg in "B200LDLT_FRONT=1" "B200LDLT_FRONT=0"; do echo "== $cfg"; env $cfg B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -30 | awk '{printf "%s:%s ", $3, $9} END {print ""}'; done
timeout 300 python tools/run_config.py c4 2>&1 | tail -4
