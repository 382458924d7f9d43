"""`ncu -i X.ncu-rep --page raw --csv` -> compact JSON of the metrics DESIGN.md quotes (per launch): duration, DRAM bytes, tensor-pipe
utilisation, shared-memory->tensor traffic, L2 throughput.  Usage: python tools/ncu_top_to_json.py raw.csv out.json label1,label2,..."""
import csv, json, sys
raw, out = sys.argv[1:3]
labels = sys.argv[3].split("|") if len(sys.argv) > 3 else []
rows = list(csv.reader(l for l in open(raw, errors="replace") if l.startswith('"')))
H, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__registers_per_thread", "smsp__cycles_active.avg"]
ix = {h: i for i, h in enumerate(H)}
cols = [w for w in want if w in ix] + [h for h in H if "pipe_tensor" in h and h not in want]
launches = []
for k, r in enumerate(data):
    d = dict(launch=labels[k // 2] if k // 2 < len(labels) else "", kernel=r[ix["Kernel Name"]][:60])
    for c in cols:
        d[c] = r[ix[c]]
    launches.append(d)
json.dump(dict(command="ncu --set full --clock-control none --import-source on python tools/top_kernels.py (2 launches per shape)",
               units={c: units[ix[c]] for c in cols}, launches=launches), open(out, "w"), indent=1)
for l in launches:
    print(l["launch"][:44].ljust(46), l["kernel"][:34].ljust(36), l.get("gpu__time_duration.sum"),
          l.get("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"), l.get("dram__bytes_read.sum"), l.get("dram__bytes_write.sum"))
