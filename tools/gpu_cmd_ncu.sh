# launch list of one factorisation + solve of C2, then ncu --set full of the longest k_upd3 launch
# (B200LDLT_LOOKAHEAD=0: every launch on the engine stream, so that launch indices are deterministic)
set -x
python tools/profile_run.py 64 1 > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_run.py 64 1 > gpurun_out/ncu_list.log 2>&1
export B200LDLT_LOOKAHEAD=0
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_upd3 --csv --log-file gpurun_out/upd3_list.csv python tools/profile_run.py 64 1 > /dev/null 2>&1
IDX=$(python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/upd3_list.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
best=max(range(len(rows)), key=lambda i: float(rows[i]['Metric Value'].replace(',','')))
print(best)
PY
)
echo "longest k_upd3 launch index $IDX"
ncu --set full --clock-control none --import-source on -k regex:k_upd3 --launch-skip $IDX --launch-count 1 -o gpurun_out/r2_k_upd3 -f python tools/profile_run.py 64 1 > gpurun_out/ncu_full_upd3.log 2>&1
tail -2 gpurun_out/ncu_full_upd3.log
ls -la gpurun_out/*.ncu-rep
