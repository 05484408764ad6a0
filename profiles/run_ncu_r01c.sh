#!/bin/bash
# Round-1 third pass: kernels rewritten after the earlier captures + the LAP kernel.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-merge --no-cpu-baseline"
$CMD > gpurun_out/plain_c.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_eager/" --set full --clock-control none --import-source on \
    -k regex:cross_finalize -s 120 -c 2 -o gpurun_out/finalize_v2_r01 $CMD > gpurun_out/ncu_c1.log 2>&1
ncu --nvtx --nvtx-include "plb_eager/" --metrics gpu__time_duration.sum --clock-control none -c 5000 \
    --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_c0.log 2>&1
CMD2="python profiles/pleas_step_driver.py"
$CMD2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_pleas/" --set full --clock-control none --import-source on \
    -k regex:pack_im2col -s 20 -c 2 -o gpurun_out/pack_im2col_r01 $CMD2 > gpurun_out/ncu_c2.log 2>&1
ncu --nvtx --nvtx-include "plb_pleas/" --metrics gpu__time_duration.sum --clock-control none -c 4000 \
    --csv --log-file gpurun_out/launches_pleas_r01c.csv $CMD2 > gpurun_out/ncu_c3.log 2>&1
CMD3="python profiles/lap_single.py"
$CMD3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lap_kernel -s 1 -c 1 \
    -o gpurun_out/lap_n2048_r01 $CMD3 > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/plain_c3.log
