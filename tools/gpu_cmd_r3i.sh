timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "B200LDLT_SMALL_STRIPS=1" "B200LDLT_SMALL_STRIPS=0"; do echo "== $cfg"; env $cfg B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -13 | awk '{printf "%s:%s ", $3, $9} END {print ""}'; done
B200LDLT_TIMELINE=1 timeout 300 python tools/profile_run.py 64 2 2>&1 | grep "L24" | tail -40 > gpurun_out/r3i_timeline.txt
timeout 600 python tools/run_config.py c3 2>&1 | tail -4
