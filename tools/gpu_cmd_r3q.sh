ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fwd|k_bwd|k_permute|k_solve" --csv --log-file gpurun_out/r3q_solve.csv python tools/profile_run.py 64 1 > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/r3q_solve.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
tot=0
for r in rows:
    t=float(r['Metric Value'].replace(',',''))/1e3; tot+=t
    print(r['Kernel Name'][:40].replace('void ',''), r['Grid Size'], "%.1f us"%t)
print("total %.1f us"%tot)
PY
