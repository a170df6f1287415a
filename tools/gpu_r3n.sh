#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/r3n_beam_launches.csv python tools/list_beam_launches.py > $out/r3n_beam.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv, collections, re
rows=list(csv.DictReader([l for l in open("$out/r3n_beam_launches.csv") if l.startswith('"')]))
agg=collections.OrderedDict()
for r in rows:
    n=re.sub(r"\(.*","",r["Kernel Name"])[:90]
    us=float(r["Metric Value"].replace(",",""))*{"ns":1e-3,"us":1.0,"usecond":1.0,"ms":1e3}.get(r["Metric Unit"],1.0)
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=us
tot=sum(v[1] for v in agg.values()); cnt=sum(v[0] for v in agg.values())
print(f"one forward_beam_search call (B=64, beam 5, 19 decode steps, eager): {cnt} launches, {tot/1e3:.2f} ms of kernel time (serialised under ncu)")
print("| kernel | launches | us | us/launch |\n|---|---|---|---|")
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"| \`{k}\` | {v[0]} | {v[1]:.0f} | {v[1]/v[0]:.1f} |")
PY
