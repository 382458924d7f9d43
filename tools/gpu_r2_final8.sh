#!/bin/bash
# round-2 multi-GPU evidence on one 8-GPU box: the default line at N = 8 and 4, config 5 (UNet++) at N = 8, config 4 (slide) at N = 8
# (the 2-rank parity tests, the peer-collective check and N = 2: tools/gpu_r2_final2.sh on a 2-GPU box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2p
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 5" > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "n8 exit $?"
timeout 600 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 20 --warmup 5" > ${P}_bench_n4.json 2> ${P}_bench_n4.err; echo "n4 exit $?"
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --model unetpp --steps 10 --warmup 3 --no-gpu-eager" > ${P}_bench_unetpp_n8.json 2> ${P}_bench_unetpp_n8.err; echo "unetpp n8 exit $?"
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --workload slide --steps 3 --warmup 2" > ${P}_bench_slide_n8.json 2> ${P}_bench_slide_n8.err; echo "slide n8 exit $?"
for f in n8 n4 unetpp_n8 slide_n8; do python -c "
import json,sys
d=json.loads(open('${P}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['config'].get('exchange'), d.get('clocks'))" 2>&1 | tail -1; done
