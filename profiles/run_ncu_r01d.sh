#!/bin/bash
# Round-1 final pass: launch list of the timed steps after the fused narrow-tap kernel, paired packs and
# LAP v2 (kernel shares of a calibration step), plus the whole-path bench that the numbers in RESULTS.md cite.
set -x
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
CMD="python bench.py --steps 2 --warmup 3 --no-merge --no-cpu-baseline"
$CMD > gpurun_out/plain_d.log 2>&1 &&
ncu --nvtx --nvtx-include "plb_eager/" --metrics gpu__time_duration.sum --clock-control none -c 5000 \
    --csv --log-file gpurun_out/launches_r01d.csv $CMD > gpurun_out/ncu_d0.log 2>&1
tail -1 gpurun_out/ncu_d0.log
