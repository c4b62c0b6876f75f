timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_pytest.log
for cfg in "B200LDLT_PDL=1" "B200LDLT_PDL=0" "B200LDLT_PDL=1 B200LDLT_UPD3_MINM=1024"; do echo "== $cfg"; env $cfg B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -13 | awk '{printf "%s:%s ", $3, $9} END {print ""}'; done
