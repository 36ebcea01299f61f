#!/bin/bash
# ncu launch list of an arbitrary bench command line: tools/ncu_list.sh <tag> <bench args...>
TAG=$1; shift
O=gpurun_out; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches.csv python bench.py "$@" > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/${TAG}_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:44]].append(float(r[vi].replace(',','')))
for k,v in d.items():
    v2=v[len(v)//2:]
    print("%-46s n=%4d  median(last half)=%9.1f %s"%(k,len(v),sorted(v2)[len(v2)//2],rows[1][ui]))
PY
