#!/bin/bash
# round-2 GPU call A: full GPU test suite, default bench line, in-step event profile, ncu launch list with DRAM bytes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench exit $?"; head -c 1500 gpurun_out/r2a_bench_n1.json; tail -n 5 gpurun_out/r2a_bench_n1.err
timeout 600 python tools/step_profile.py > gpurun_out/r2a_step_profile.txt 2>&1; echo "step_profile exit $?"; head -n 40 gpurun_out/r2a_step_profile.txt
timeout 900 python tools/ncu_step.py > gpurun_out/r2a_ncu_plain.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2a_bw_raw.csv python tools/ncu_step.py > gpurun_out/r2a_ncu.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/r2a_ncu.log
