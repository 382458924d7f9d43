#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_tests.txt 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2h_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print({k.split("::")[-1]:(round(v["ms"],2),round(v["tflops"])) for k,v in d["roofline"]["all_tcgen05_kernels"]["per_kernel"].items()})
PY
