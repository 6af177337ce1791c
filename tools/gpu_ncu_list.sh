#!/bin/bash
# launch list (warm L2) of one update of $1 -> gpurun_out/$2_launches_$1_warm.csv
mkdir -p gpurun_out
A=$1; TAG=${2:-r2f}
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s ${3:-500} -c ${4:-60} --csv \
    --log-file gpurun_out/${TAG}_launches_${A}_warm.csv python bench.py --algo $A --steps 6 --warmup 4 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_launches_${A}_warm.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
print(' '.join(f"{r[ki].replace('void ','').replace('oprl::','')[:14]}:{r[gi].strip('()').split(',')[0]}:{float(r[vi].replace(',',''))/1000:.1f}" for r in rows[1:]))
PY
