"""raw `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of ONE eager training step
(tools/ncu_step.py) -> per-kernel JSON: launches, device time, DRAM bytes read / written and the achieved DRAM GB/s against the measured
copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  ncu serialises the launches with a cold L2, so these are per-kernel DRAM figures, not
in-step timings (those are the CUDA-event table of tools/step_profile.py, merged in with --events).
Usage: python tools/ncu_bw_summary.py gpurun_out/r2_bw_raw.csv profiles/r2_ncu_bw_kernels.json [--events gpurun_out/r2_step_profile.txt]"""
import collections, csv, json, os, re, sys

raw, out = sys.argv[1:3]
events = sys.argv[sys.argv.index("--events") + 1] if "--events" in sys.argv else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6546.2
mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(mp):
    peak = json.load(open(mp)).get("hbm_gbs", peak)
lines = [l for l in open(raw, errors="replace") if l.startswith('"')]
rows = list(csv.reader(lines))
H = rows[0]
ix = {h: i for i, h in enumerate(H)}
UNIT_T = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
per_id = collections.OrderedDict()
for r in rows[1:]:
    if len(r) != len(H):
        continue
    d = per_id.setdefault(r[ix["ID"]], dict(name=re.sub(r"\(.*$", "", r[ix["Kernel Name"]])[:90]))
    m, v, u = r[ix["Metric Name"]], float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    if m == "gpu__time_duration.sum":
        d["ms"] = v * UNIT_T[u]
    elif m == "dram__bytes_read.sum":
        d["rd"] = v * UNIT_B[u]
    elif m == "dram__bytes_write.sum":
        d["wr"] = v * UNIT_B[u]
agg = collections.OrderedDict()
for d in per_id.values():
    if "ms" not in d:
        continue
    a = agg.setdefault(d["name"], dict(launches=0, ms=0.0, dram_read_bytes=0.0, dram_write_bytes=0.0, max_launch=None))
    a["launches"] += 1; a["ms"] += d["ms"]; a["dram_read_bytes"] += d.get("rd", 0.0); a["dram_write_bytes"] += d.get("wr", 0.0)
    if a["max_launch"] is None or d["ms"] > a["max_launch"]["ms"]:
        a["max_launch"] = dict(ms=d["ms"], dram_read_bytes=d.get("rd", 0.0), dram_write_bytes=d.get("wr", 0.0))
total = sum(a["ms"] for a in agg.values())
kern = []
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = (a["dram_read_bytes"] + a["dram_write_bytes"]) / 1e9 / (a["ms"] * 1e-3) if a["ms"] > 0 else 0.0
    ml = a["max_launch"]
    ml["dram_gbs"] = (ml["dram_read_bytes"] + ml["dram_write_bytes"]) / 1e9 / (ml["ms"] * 1e-3) if ml["ms"] > 0 else 0.0
    kern.append(dict(kernel=n, tensor_core="umma" in n, launches=a["launches"], ms=round(a["ms"], 4), share=round(a["ms"] / total, 4),
                     dram_read_mb=round(a["dram_read_bytes"] / 1e6, 2), dram_write_mb=round(a["dram_write_bytes"] / 1e6, 2), dram_gbs=round(gbs, 1),
                     dram_frac_of_measured_peak=round(gbs / peak, 3),
                     heaviest_launch=dict(ms=round(ml["ms"], 4), dram_mb=round((ml["dram_read_bytes"] + ml["dram_write_bytes"]) / 1e6, 2),
                                          dram_gbs=round(ml["dram_gbs"], 1), frac=round(ml["dram_gbs"] / peak, 3))))
res = dict(command="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                   "python tools/ncu_step.py  (one eager STC-UNet training step, bf16, N=16, 512x512; launches serialised, cold L2)",
           hbm_peak_gbs=peak, peak_source="MEASURED_PEAKS.json hbm_gbs (copy, read+write)", launches=sum(a["launches"] for a in agg.values()), total_ms=round(total, 3),
           non_tensor_ms=round(sum(k["ms"] for k in kern if not k["tensor_core"]), 3), kernels=kern)
if events and os.path.exists(events):
    ev = []
    for l in open(events):
        m = re.match(r"^(stc_\w+)\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)\s*$", l)
        if m:
            ev.append(dict(entry_point=m.group(1), calls=int(m.group(2)), ms=float(m.group(3)), algorithmic_gb=float(m.group(4)),
                           achieved_gbs=float(m.group(5)), frac_of_measured_peak=round(float(m.group(5)) / peak, 3)))
    res["in_step_cuda_events"] = dict(what="tools/step_profile.py: CUDA events around every C-ABI call of one in-stream step (warm clocks / L2); algorithmic bytes = "
                                           "bytes of the call's tensor operands, each counted once", entries=ev)
json.dump(res, open(out, "w"), indent=1)
print(f"{len(kern)} kernels, {res['launches']} launches, {total:.2f} ms, non-tensor {res['non_tensor_ms']:.2f} ms")
for k in kern[:25]:
    print("%8.3f ms %5d  %7.1f GB/s (%.2f)  %s" % (k["ms"], k["launches"], k["dram_gbs"], k["dram_frac_of_measured_peak"], k["kernel"]))
