#!/bin/bash
# Round-2 profiling pass (run under gpurun): launch list of the bench command
# and one `ncu --set full` capture per dominant kernel.  Summarise afterwards
# with `python profiles/summarize_ncu.py gpurun_out/r2_*.ncu-rep`.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 \
    --sustain 0 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_under_ncu.log 2>&1
for n in fp64 fast boxmuller; do
  $NCU -k regex:mdpp_jit_rollout -s 3 -c 1 -o gpurun_out/r2_discrete_rollout_$n \
      python tools/time_one.py $n 65536 1000 2 > /dev/null 2>&1
done
$NCU -k regex:mdpp_jit_rollout -s 3 -c 1 -o gpurun_out/r2_discrete_rollout_hetero_fp64 \
    python tools/time_hetero.py fp64 > /dev/null 2>&1
$NCU -k regex:mdpp_jit_rollout -s 3 -c 1 -o gpurun_out/r2_discrete_rollout_hetero_fast \
    python tools/time_hetero.py fast > /dev/null 2>&1
$NCU -k regex:mdpp_jit_continuous -s 3 -c 1 \
    -o gpurun_out/r2_continuous_rollout python tools/time_continuous.py > /dev/null 2>&1
$NCU -k regex:mdpp_jit_rollout -s 12 -c 1 -o gpurun_out/r2_discrete_step_T1 \
    python tools/time_step_kernel.py fp64 > /dev/null 2>&1
cat > /tmp/r2_render.py <<'PY'
import sys, warnings, torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorRLToyEnv
base = dict(seed=0, state_space_type="discrete", action_space_type="discrete", state_space_size=8, action_space_size=8, sequence_length=1, delay=0, reward_density=0.25, terminal_state_density=0.25, image_representations=True, image_width=100, image_height=100)
cfg = dict(base, image_transforms="shift", image_sh_quant=4) if sys.argv[1] == "shift" else dict(base, image_transforms="shift,scale,rotate", image_scale_range=(0.5, 1.5), image_ro_quant=1, image_sh_quant=4)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    env = VectorRLToyEnv(16384, autoreset=True, horizon=100, **cfg)
st = torch.randint(0, 8, (16384,), device="cuda")
for k in range(6): env.render_observation(st, step_index=k)
torch.cuda.synchronize()
PY
$NCU -k regex:render_discrete -s 4 -c 1 -o gpurun_out/r2_render_discrete python /tmp/r2_render.py shift > /dev/null 2>&1
$NCU -k regex:render_discrete -s 4 -c 1 -o gpurun_out/r2_render_discrete_rotate python /tmp/r2_render.py rot > /dev/null 2>&1
ls -la gpurun_out | grep r2_
