timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "B200LDLT_SOLVE_INV=1" "B200LDLT_SOLVE_INV=0"; do echo "== $cfg"; env $cfg timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -3; done
env B200LDLT_SOLVE_INV=1 timeout 300 python tools/run_config.py c3s 2>&1 | tail -2
env B200LDLT_SOLVE_INV=0 timeout 300 python tools/run_config.py c3s 2>&1 | tail -1
