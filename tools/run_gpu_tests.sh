#!/bin/bash
# Runs the GPU test groups in separate processes (a faulting tcgen05 kernel poisons its CUDA context) and
# keeps the logs under gpurun_out/.  Usage (on the GPU box): bash tools/run_gpu_tests.sh [group ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -m gpu -q --timeout 900 -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name"; timeout 1500 $PY "$@" > gpurun_out/test_$name.log 2>&1; echo "exit $? ($name)"; tail -n 25 gpurun_out/test_$name.log; }
groups=${@:-"ops umma_first umma_fprop umma_halo umma_grad umma_attn umma_misc model_fp32 model_bf16 model_misc"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
for g in $groups; do
case $g in
  ops) run ops tests/test_ops_gpu.py ;;
  umma_first) run umma_first "tests/test_umma_gpu.py::test_conv_fprop_tcgen05[shape0]" "tests/test_umma_gpu.py::test_conv_fprop_tcgen05[shape1]" ;;
  umma_fprop) run umma_fprop tests/test_umma_gpu.py -k "test_conv_fprop_tcgen05" ;;
  umma_grad) run umma_grad tests/test_umma_gpu.py -k "test_conv_dgrad_wgrad_tcgen05 or test_wgrad_halo_tcgen05" ;;
  umma_attn) run umma_attn tests/test_umma_gpu.py -k "test_attention_tcgen05" ;;
  umma_halo) run umma_halo tests/test_umma_gpu.py -k "test_conv_halo_tcgen05" ;;
  umma_misc) run umma_misc tests/test_umma_gpu.py -k "test_linear_tokens_tcgen05 or test_tcgen05_matches_simt_large" ;;
  model_fp32) run model_fp32 tests/test_model_gpu.py -k "fp32" ;;
  model_bf16) run model_bf16 tests/test_model_gpu.py -k "bf16" ;;
  model_misc) run model_misc tests/test_model_gpu.py -k "not fp32 and not bf16" ;;
esac
done
