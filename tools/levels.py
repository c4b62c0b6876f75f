"""Per-level summary of an ncu launch list (last factorisation in the file): python tools/levels.py launches.csv"""
import csv, re, sys, collections
FL = [1.199e+05,3.007e+05,3.161e+06,8.271e+06,3.172e+07,7.623e+07,1.590e+08,2.955e+08,4.126e+08,5.023e+08,6.062e+08,7.083e+08,8.144e+08,9.444e+08,1.074e+09,1.194e+09,3.210e+09,7.284e+09,1.620e+10,2.395e+10,4.219e+10,7.768e+10,9.015e+10,9.498e+10,7.733e+10]
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [(re.sub(r"\(.*", "", r["Kernel Name"]), float(r["Metric Value"]), r["Grid Size"]) for r in csv.DictReader(lines)]
segs, cur = [], []
for r in rows:
    cur.append(r)
    if r[0] == 'k_stats': segs.append(cur); cur = []
lv, cur = [], []
for r in segs[-1]:
    if r[0] == 'k_setup' and cur: lv.append(cur); cur = []
    cur.append(r)
lv.append(cur)
tot_all = 0
for i, l in enumerate(lv):
    t = collections.defaultdict(float); c = collections.Counter()
    for n, v, g in l: t[n] += v / 1e3; c[n] += 1
    tot = sum(t.values()); tot_all += tot
    eff = FL[i] / (t.get('k_update', 0) * 1e-6) / 1e12 if len(lv) == 25 and t.get('k_update') else 0
    print(i, l[0][2], "tot %.0fus upd %.1fTF/s" % (tot, eff), " ".join("%s:%d/%.0f" % (k.replace('k_', ''), c[k], t[k]) for k in sorted(t, key=lambda k: -t[k])))
print("sum of kernel times (factor) %.1f ms" % (tot_all / 1e3))
