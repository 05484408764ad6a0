#!/bin/bash
# Round-1 profiling recipe (run under gpurun): plain run first, then the launch list of the
# timed region (NVTX range plb_timed) and full captures of the dominant kernel.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-merge --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_timed/" --metrics gpu__time_duration.sum --clock-control none -c 5000 \
    --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_timed/" --set full --clock-control none --import-source on \
    -k regex:gemm3xtf32 -s 150 -c 3 -o gpurun_out/gemm_c2048_r01 $CMD > gpurun_out/ncu_gemm1.log 2>&1
ncu --nvtx --nvtx-include "plb_timed/" --set full --clock-control none --import-source on \
    -k regex:gemm3xtf32 -s 40 -c 2 -o gpurun_out/gemm_mid_r01 $CMD > gpurun_out/ncu_gemm2.log 2>&1
ncu --nvtx --nvtx-include "plb_timed/" --set full --clock-control none --import-source on \
    -k regex:pack_split -s 0 -c 2 -o gpurun_out/pack_r01 $CMD > gpurun_out/ncu_pack.log 2>&1
tail -3 gpurun_out/plain.log | cut -c1-400
