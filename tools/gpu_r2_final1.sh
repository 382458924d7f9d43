#!/bin/bash
# round-2 final single-GPU evidence: GPU test suite, the default bench line, configs 4 and 5 through bench.py, in-step event profile,
# ncu launch list with DRAM bytes of one eager step, ncu --set full of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2p
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > ${P}_smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > ${P}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 ${P}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "bench ref exit $?"
timeout 900 python bench.py --workload slide --steps 3 --warmup 2 > ${P}_bench_slide.json 2> ${P}_bench_slide.err; echo "slide exit $?"
timeout 900 python bench.py --model unetpp --steps 10 --warmup 3 > ${P}_bench_unetpp_n1.json 2> ${P}_bench_unetpp_n1.err; echo "unetpp exit $?"
timeout 900 python bench.py --model unet --steps 10 --warmup 3 --no-gpu-eager > ${P}_bench_unet_n1.json 2> ${P}_bench_unet_n1.err; echo "unet exit $?"
timeout 900 python bench.py --model unet_b --steps 10 --warmup 3 --no-gpu-eager > ${P}_bench_unet_b_n1.json 2> ${P}_bench_unet_b_n1.err; echo "unet_b exit $?"
timeout 600 python tools/step_profile.py --dense > ${P}_step_profile.txt 2>&1; echo "step_profile exit $?"
timeout 600 python tools/conv_sweep.py > ${P}_conv_sweep.txt 2>&1; echo "sweep exit $?"
timeout 900 python tools/ncu_step.py > ${P}_ncu_plain.log 2>&1 && \
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file ${P}_bw_raw.csv python tools/ncu_step.py > ${P}_ncu.log 2>&1; echo "ncu launch list exit $?"
timeout 600 python tools/top_kernels.py > ${P}_top_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:umma -o ${P}_top python tools/top_kernels.py > ${P}_ncu_top.log 2>&1; echo "ncu full exit $?"
ncu -i ${P}_top.ncu-rep --page raw --csv > ${P}_top_raw.csv 2>/dev/null; echo "export exit $?"
rm -f ${P}_top.ncu-rep   # the report itself exceeds what gpurun copies back (64 MiB); the raw CSV is what profiles/ keeps
du -sh gpurun_out
