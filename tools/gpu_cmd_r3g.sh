timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/front_check.py 2>&1 | grep -v "^OK" | tail -8
B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -30 | awk '{printf "%s:%s ", $3, $9} END {print ""}'
B200LDLT_LEVELTIMES=1 timeout 300 python tools/run_config.py c4 2>&1 | tail -14
