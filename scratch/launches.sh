#!/bin/bash
# launch list of the DAA bench: scratch/launches.sh <out.csv>
MOPOE_BENCH_SKIP_TRAIN=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $1 python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
python - $1 <<'PY'
import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
c=collections.defaultdict(list)
for r in rows[1:]: c[r[ik][:60]].append(float(r[iv].replace(",","")))
tot=sum(sum(v) for v in c.values())
for k,v in c.items(): print("%-62s n=%3d avg %9.1f us  %5.1f%%" % (k,len(v),sum(v)/len(v)/1000, 100*sum(v)/tot))
PY
