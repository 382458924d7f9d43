#!/bin/bash
cd "$(dirname "$0")/.." 2>/dev/null
mkdir -p gpurun_out
P=gpurun_out/r2q
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 5" > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "n8 exit $?"
python -c "
import json
d=json.loads(open('${P}_bench_n8.json').read().strip().splitlines()[-1]); print('n8', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['config'].get('exchange'), d.get('clocks'))"
