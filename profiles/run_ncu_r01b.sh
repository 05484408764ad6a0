#!/bin/bash
# Round-1 (second pass, persistent GEMM): launch list of the un-captured re-run of the timed steps
# (NVTX range plb_eager; the timed region itself replays a CUDA graph) + full captures.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-merge --no-cpu-baseline"
$CMD > gpurun_out/plain_b.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_eager/" --metrics gpu__time_duration.sum --clock-control none -c 5000 \
    --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_launches_b.log 2>&1
$CMD > gpurun_out/plain_b2.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_eager/" --set full --clock-control none --import-source on \
    -k regex:gemm3xtf32_v2 -s 150 -c 2 -o gpurun_out/gemm_v2_c2048_r01 $CMD > gpurun_out/ncu_gemm_b1.log 2>&1
ncu --nvtx --nvtx-include "plb_eager/" --set full --clock-control none --import-source on \
    -k regex:gemm3xtf32_v2 -s 60 -c 2 -o gpurun_out/gemm_v2_mid_r01 $CMD > gpurun_out/ncu_gemm_b2.log 2>&1
ncu --nvtx --nvtx-include "plb_eager/" --set full --clock-control none --import-source on \
    -k regex:cross_finalize -s 150 -c 2 -o gpurun_out/finalize_r01 $CMD > gpurun_out/ncu_fin_b.log 2>&1
tail -1 gpurun_out/plain_b.log | cut -c1-300
