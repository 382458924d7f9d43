"""raw `ncu --metrics gpu__time_duration.sum --csv` log -> compact launch list (idx,kernel,ms,grid,block) + per-kernel summary.
Usage: python tools/ncu_launch_list.py gpurun_out/launches_raw.csv profiles/r1_launches_step_v8.csv profiles/r1_launches_step_v8_summary.txt"""
import collections
import csv
import re
import sys

raw, out_csv, out_sum = sys.argv[1:4]
lines = [l for l in open(raw, errors="replace") if l.startswith('"')]
rows = list(csv.reader(lines))
H = rows[0]
ix = {h: i for i, h in enumerate(H)}
recs = []
for r in rows[1:]:
    if len(r) != len(H) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = val * {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}[unit]
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])[:80]
    recs.append((name, ms, r[ix["Grid Size"]], r[ix["Block Size"]]))
with open(out_csv, "w") as f:
    f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/ncu_step.py; one eager training step "
            "(STC-UNet bf16 N=16 512x512; cold cache, serialised launches)\n")
    w = csv.writer(f)
    w.writerow(["idx", "kernel", "ms", "grid", "block"])
    for i, (n, ms, g, b) in enumerate(recs):
        w.writerow([i, n, "%.6f" % ms, g, b])
agg = collections.OrderedDict()
for n, ms, _, _ in recs:
    d = agg.setdefault(n, [0, 0.0])
    d[0] += 1; d[1] += ms
total = sum(v[1] for v in agg.values())
with open(out_sum, "w") as f:
    f.write(f"# one STC-UNet bf16 N=16 512x512 training step under ncu (serialised, cold cache): {len(recs)} launches, {total:.2f} ms\n")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%9.3f ms %5.1f%% %5d  %s\n" % (ms, 100 * ms / total, c, n))
print(open(out_sum).read()[:1500])
