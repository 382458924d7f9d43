#!/bin/bash
# round-2 multi-GPU evidence on one 8-GPU box: the default line at N = 8 / 4 / 2, config 5 (UNet++) at N = 8, config 4 (slide) at N = 8,
# the 2-rank parity tests and the peer-collective check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2p
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 5" > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "n8 exit $?"
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --model unetpp --steps 10 --warmup 3" > ${P}_bench_unetpp_n8.json 2> ${P}_bench_unetpp_n8.err; echo "unetpp n8 exit $?"
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --workload slide --steps 3 --warmup 2" > ${P}_bench_slide_n8.json 2> ${P}_bench_slide_n8.err; echo "slide n8 exit $?"
timeout 600 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 20 --warmup 5" > ${P}_bench_n4.json 2> ${P}_bench_n4.err; echo "n4 exit $?"
timeout 600 bash -c "$(declare -f run); run 2 bench.py --gpus 2 --steps 20 --warmup 5" > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "n2 exit $?"
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > ${P}_pytest_dist.log 2>&1; echo "dist pytest exit $?"; tail -n 3 ${P}_pytest_dist.log
timeout 300 bash -c "$(declare -f run); run 2 tools/peer_check.py" > ${P}_peer_check_n2.txt 2>&1; echo "peer_check exit $?"; tail -n 4 ${P}_peer_check_n2.txt
for f in n8 unetpp_n8 slide_n8 n4 n2; do python -c "
import json,sys
d=json.load(open('${P}_bench_$f.json')); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['config'].get('exchange'))" 2>&1 | tail -1; done
