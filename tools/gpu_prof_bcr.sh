#!/bin/bash
# launch list of one BCR build (5 M x 150 bp): which kernels the cycle time goes to
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bcr_launches.csv python tools/bench_bcr.py --reads ${READS:-3000000} --len 150 --ref-reads 0 > gpurun_out/bcr_prof.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/bcr_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try:
        name = r[4].split('(')[0][-60:]; v = float(r[-1].replace(',', ''))
    except Exception:
        continue
    unit = r[-2]
    if unit in ('us', 'usecond'): v /= 1e3
    elif unit in ('ns', 'nsecond'): v /= 1e6
    elif unit in ('s', 'second'): v *= 1e3
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%6.1f%% %10.1f ms %6d  %s" % (100 * v[1] / tot, v[1], v[0], k))
PY
