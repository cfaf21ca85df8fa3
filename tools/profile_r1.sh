#!/bin/bash
# Round-1 profiling pass (run under gpurun): launch list of the bench command
# and one `ncu --set full` capture per dominant kernel.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r1_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdpp_jit_rollout -s 3 -c 1 \
    -o gpurun_out/r1_discrete_rollout python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdpp_jit_continuous -s 3 -c 1 \
    -o gpurun_out/r1_continuous_rollout python tools/time_continuous.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_discrete -s 3 -c 1 \
    -o gpurun_out/r1_render_discrete python tools/time_images.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:discrete_rollout_kernel -s 2 -c 1 \
    -o gpurun_out/r1_discrete_rollout_aot python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-jit > /dev/null 2>&1
ls -la gpurun_out
