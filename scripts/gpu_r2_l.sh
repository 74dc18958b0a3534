#!/bin/bash
# grid bands v2: per-kernel, per-step durations (ncu launch list, one GPU, whole grid = one band)
O=gpurun_out/r2l
mkdir -p $O
G=${1:-4096}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:grid_shard -c 200 --csv --log-file $O/launches_$G.csv python scripts/profile_grid_bands.py $G 20 > $O/prof.log 2>&1
tail -2 $O/prof.log
python - <<PY
import csv, collections
rows = []
with open('$O/launches_$G.csv') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append((r['Kernel Name'].split('(')[0].replace('jxb::','').replace('void ',''), float(r['Metric Value'].replace(',','')), r['Metric Unit']))
per = collections.OrderedDict()
for k, v, u in rows:
    per.setdefault(k, []).append(v / (1000.0 if u in ('ns','nsecond') else 1.0))
with open('$O/per_step_$G.txt', 'w') as out:
    for k, v in per.items():
        line = '%-45s n=%d sum=%.1f us : ' % (k[:45], len(v), sum(v)) + ' '.join('%.1f' % x for x in v[:20])
        print(line); out.write(line + '\n')
PY
