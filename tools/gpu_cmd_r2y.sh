timeout 600 python tools/front_check.py 2>&1 | grep -v "^OK" | tail -8
timeout 600 python tools/front_check.py 2>&1 | grep -E "lap7_64|qpband_1000000|lap7_48"
for cfg in "B200LDLT_FRONT=1"; do echo "== $cfg"; env $cfg B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 3 2>&1 | tail -30 | awk '{printf "%s:%s ", $3, $9} END {print ""}'; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_front --csv --log-file gpurun_out/r2y_front.csv python tools/profile_run.py 64 1 > /dev/null 2>&1
grep k_front gpurun_out/r2y_front.csv | awk -F'","' '{print $5, $(NF)}' | tr -d '"' | head -20
