mkdir -p gpurun_out/r04k
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r04k/prep.csv python tools/prep_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r04k/prep.csv")) if len(r)>5]
hdr=None; v=[]
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if "strip_prep" in d["Kernel Name"]: v.append(float(d["Metric Value"].replace(",",""))/1e3)
print("strip_prep durations (us), 6 per phase setting [after phase 1 | after barrier+scan | whole]:")
for i in range(0,len(v),6): print("  ", ["%.1f"%x for x in v[i:i+6]])
PY
