#!/bin/bash
# launch list of one PLeaS normal-equation step (NVTX range plb_pleas) + full capture of pack_im2col
set -x
CMD="python profiles/pleas_step_driver.py"
$CMD > gpurun_out/pleas_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_pleas/" --metrics gpu__time_duration.sum --clock-control none -c 4000 \
    --csv --log-file gpurun_out/launches_pleas_r01.csv $CMD > gpurun_out/ncu_pleas.log 2>&1
tail -2 gpurun_out/pleas_plain.log
