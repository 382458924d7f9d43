#!/bin/bash
# round-2 2-GPU evidence: the 2-rank parity tests, the peer-collective check and the default line at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2p
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > ${P}_pytest_dist.log 2>&1; echo "dist pytest exit $?"; tail -n 3 ${P}_pytest_dist.log
timeout 300 bash -c "$(declare -f run); run 2 tools/peer_check.py" > ${P}_peer_check_n2.txt 2>&1; echo "peer_check exit $?"; tail -n 4 ${P}_peer_check_n2.txt
timeout 600 bash -c "$(declare -f run); run 2 bench.py --gpus 2 --steps 20 --warmup 5" > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "n2 exit $?"
python -c "
import json
d=json.loads(open('${P}_bench_n2.json').read().strip().splitlines()[-1]); print('n2', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['config'].get('exchange'), d.get('clocks'))"
