timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 > gpurun_out/r2c_levels.txt 2>&1; tail -32 gpurun_out/r2c_levels.txt
for d in 1 2; do echo "DEFER=$d"; B200LDLT_DEFER=$d timeout 300 python tools/profile_run.py 64 3 2>&1 | grep -E "factor ms|berr" ; done
echo "UPD3=0"; B200LDLT_UPD3=0 timeout 300 python tools/profile_run.py 64 3 2>&1 | grep -E "factor ms|berr"
