timeout 600 python tools/front_check.py quick 2>&1 | grep -v "^OK" | tail -8
B200LDLT_TRACE=2 timeout 300 python tools/profile_run.py 64 2 2>&1 | grep k_front | tail -6
B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -30 | awk '{printf "%s:%s ", $3, $9} END {print ""}'
B200LDLT_LEVELTIMES=1 timeout 300 python tools/run_config.py c4 2>&1 | tail -16
