#!/bin/bash
tag=r2t
mkdir -p gpurun_out
out=$PWD/gpurun_out
for gk in "2 0" "8 3" "8 0"; do python tools/stripe_profile.py $gk 20 2>/dev/null | tee -a $out/stripe_stages_$tag.jsonl; done
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_stripe_2_0_$tag.csv python tools/stripe_profile.py 2 0 6 > /dev/null 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_stripe_8_3_$tag.csv python tools/stripe_profile.py 8 3 6 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ("launches_stripe_2_0_r2t.csv", "launches_stripe_8_3_r2t.csv"):
    lines=[l for l in open("gpurun_out/"+f) if not l.startswith("==")]
    rows=[]
    for row in csv.DictReader(lines):
        try: rows.append((row["Kernel Name"], float(row["Metric Value"])))
        except Exception: pass
    idx=[i for i,(k,v) in enumerate(rows) if "stripe_cull" in k]
    seg=rows[idx[-2]:idx[-1]]
    agg=collections.OrderedDict()
    for k,v in seg:
        k=k.split("(")[0][:44]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v/1000
    print(f, "kernels", len(seg), "sum us", round(sum(v for _,v in seg)/1000,1))
    for k,(n,v) in agg.items(): print(f"   {k:44s} x{n:2d} {v:8.1f}")
PY
