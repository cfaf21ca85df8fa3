#!/bin/bash
# one `ncu --set full` capture of the rotated renderer (run under gpurun)
mkdir -p gpurun_out
cat > /tmp/r_render.py <<'PY'
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
NAME=${1:-r2b_render_discrete_rotate}
ncu --set full --clock-control none --import-source on -k regex:render_discrete -s 4 -c 1 \
    -o gpurun_out/$NAME python /tmp/r_render.py rot > /dev/null 2>&1
ls -la gpurun_out | grep $NAME
