#!/bin/bash
# parity of the halo convs (pair variant on) + role profile with and without CTA pairs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py -x -q -m gpu > gpurun_out/r2g_umma_tests.txt 2>&1; echo "umma tests rc=$?"
tail -3 gpurun_out/r2g_umma_tests.txt
timeout 300 python tools/convh_prof.py > gpurun_out/r2g_convh_prof_cta2.txt 2>&1; echo "prof rc=$?"
STC_CONVH_CTA2=0 timeout 300 python tools/convh_prof.py > gpurun_out/r2g_convh_prof_cta1.txt 2>&1
