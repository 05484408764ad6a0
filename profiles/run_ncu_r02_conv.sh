#!/bin/bash
# Round 2, convolution kernel: one `--set full` capture of conv3xtf32_kernel on two layer classes (a 1x1 layer of
# layer1 and the 3x3 layer of layer3, both models in one launch), and the launch list of two calibration steps
# with the kernel in place (kernel shares of the step).  Each command runs plain first.
set -x
for i in 0 2; do
  python profiles/experiments/conv_probe.py $i > gpurun_out/conv_probe_plain$i.log 2>&1 &&
  ncu --set full --import-source on --clock-control none -k regex:conv3xtf32 -c 1 -f \
      -o gpurun_out/r02_conv_shape$i python profiles/experiments/conv_probe.py $i > gpurun_out/ncu_conv$i.log 2>&1
done
CMD="python bench.py --steps 2 --warmup 3 --no-merge --no-cpu-baseline --no-extra-rooflines"
$CMD > gpurun_out/plain_r02c.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_eager/" --metrics gpu__time_duration.sum --clock-control none -c 5000 \
    --csv --log-file gpurun_out/launches_r02_conv.csv $CMD > gpurun_out/ncu_r02c.log 2>&1
tail -2 gpurun_out/ncu_r02c.log
