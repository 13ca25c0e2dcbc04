for ks in 1 2 4 8; do
VMMT_ROWLIN_KS=$ks ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__cluster_size --clock-control none -k regex:rowlin -c 100 --csv --log-file gpurun_out/rowlin_ncu_$ks.csv python tools/rowlin_bench.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/rowlin_ncu_$ks.csv")))
hdr=None; agg={}
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); agg.setdefault(d["ID"],{})[d["Metric Name"]]=d["Metric Value"]
seen={}
for k,v in agg.items():
    sig=(v.get("launch__grid_size"),v.get("launch__cluster_size"))
    seen.setdefault(sig,[]).append(float(v["gpu__time_duration.sum"])/1e3)
print("KS=$ks", {k: round(min(v),1) for k,v in seen.items()})
PY
done
