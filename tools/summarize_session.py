"""Turn the raw artefacts of one tools/gpu_session.sh run (gpurun_out/<tag>/) into the committed summaries under
profiles/: launch list per kernel, selected ncu metrics per kernel, traffic.json, the bench line(s).
    python tools/summarize_session.py <tag>
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
src = "gpurun_out/%s/" % tag
KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'sm__cycles_active.min', 'sm__cycles_active.max',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'short_scoreboard', 'wait', 'barrier', 'mio_throttle', 'lg_throttle', 'no_instruction',
         'math_pipe_throttle', 'not_selected', 'branch_resolving')]

# 1. launch list
rows = list(csv.reader(open(src + "launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]; iN, iV, iM = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > iV and r[iM] == "gpu__time_duration.sum":
        agg.setdefault(r[iN].split("(")[0][:60], []).append(float(r[iV].replace(",", "")) / 1e3)
with open("profiles/%s_launches_summary.csv" % tag, "w") as f:
    f.write("kernel,launches,mean_us,min_us,max_us\n")
    for k, v in agg.items():
        f.write("%s,%d,%.2f,%.2f,%.2f\n" % (k, len(v), sum(v) / len(v), min(v), max(v)))
shutil.copy(src + "launches.csv", "profiles/%s_launches.csv" % tag)

# 2. ncu metrics per kernel
out, traffic = [], {}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for rep in ("prof.ncu-rep", "prof_nms.ncu-rep"):
    if not os.path.exists(src + rep):
        continue
    txt = subprocess.run(["ncu", "-i", src + rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); hdr, un = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    if not out:
        out.append([hdr[i] for i in idx]); out.append([un[i] for i in idx])
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        out.append([name if hdr[i] == "Kernel Name" else r[i] for i in idx])
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * mult[un[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * mult[un[hdr.index("dram__bytes_write.sum")]]
        traffic[name] = int(rd + wr)
with open("profiles/%s_ncu_kernels.csv" % tag, "w") as f:
    csv.writer(f).writerows(out)


def find(sub):
    hit = [v for k, v in traffic.items() if sub in k]
    return hit[0] if hit else None


# merged into the committed file: a session that did not capture a kernel keeps the previous session's number for it
t = json.load(open("profiles/traffic.json")) if os.path.exists("profiles/traffic.json") else {}
new = {"roi_align_fwd": find("roi_align_strip_fwd") or find("roi_align_stream_fwd") or find("roi_align_tiled_fwd"),
       "roi_align_fwd_prep": find("strip_prep") or find("roi_align_tiled_prep"),
       "roi_align_bwd_main": find("roi_align_bwd_rows<"), "roi_align_bwd_tables": find("roi_align_bwd_rows_tables"),
       "nms_mask": find("nms_mask"), "nms_scan": find("nms_scan")}
for k, v in new.items():
    if v is not None:
        t[k] = v
        t["source_" + k] = "profiles/%s_ncu_kernels.csv" % tag
if t.get("roi_align_bwd_main") is not None and t.get("roi_align_bwd_tables") is not None:
    t["roi_align_bwd"] = t["roi_align_bwd_main"] + t["roi_align_bwd_tables"]
t["source"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full --clock-control none (source_* names the capture "
               "of each kernel); roi_align_fwd = main forward kernel, roi_align_bwd = tables/transpose kernel + main gather kernel; output "
               "bytes still L2-resident at kernel end are not counted by dram__bytes_write")
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
for f in os.listdir(src):
    if f.startswith("bench") and f.endswith(".json"):
        shutil.copy(src + f, "profiles/%s_%s" % (tag, f))
print(open("profiles/%s_launches_summary.csv" % tag).read())
print(json.dumps(t, indent=1))
