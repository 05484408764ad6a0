#!/bin/bash
# Round 2, last session: lap_kernel_v3.  One `--set full` capture of the n=2048 solve (structured -cdist costs,
# second launch), the bench records of the final state, and the config-3 benchmark.  Plain runs first.
set -x
python profiles/lap_single.py > gpurun_out/lap_single_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:lap_kernel_v3 --launch-skip 1 -c 1 -f \
    -o gpurun_out/r02c_lap_v3_n2048 python profiles/lap_single.py > gpurun_out/ncu_lap_v3.log 2>&1
python bench.py > gpurun_out/bench_r02c_full.json 2> gpurun_out/bench_r02c_full.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02c_20steps.json 2> gpurun_out/bench_r02c_20steps.err
python benchmarks/weight_matching_rn50.py > gpurun_out/weight_matching_rn50_r02c.jsonl 2> gpurun_out/wm_r02c.err
tail -3 gpurun_out/ncu_lap_v3.log; tail -c 300 gpurun_out/bench_r02c_full.err; head -c 600 gpurun_out/weight_matching_rn50_r02c.jsonl
