# ncu --set full capture of the largest k_update launch (single-stream schedule so that launch indices are deterministic)
export B200LDLT_LOOKAHEAD=0 B200LDLT_GROUPS=1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_update --csv --log-file gpurun_out/upd_list.csv python tools/profile_run.py 64 1 > /dev/null 2>&1
IDX=$(python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/upd_list.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
best=max(range(len(rows)), key=lambda i: float(rows[i]['Metric Value'].replace(',','')))
print(best)
PY
)
echo "largest k_update launch index $IDX"
ncu --set full --clock-control none --import-source on -k regex:k_update --launch-skip $IDX --launch-count 1 -o gpurun_out/r1c_k_update -f python tools/profile_run.py 64 1 > gpurun_out/ncu_full_upd.log 2>&1
ls -la gpurun_out/*.ncu-rep
