cat > /tmp/tail_run.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from mdp_playground_b200 import VectorGymEnvTail
Nt, side, pad = 8192, 84, 20
tail = VectorGymEnvTail(Nt, seed=0, n_actions=18, state_space_type="discrete", delay=2, transition_noise=0.1, reward_noise=0.5, reward_scale=2.0, image_transforms="shift", image_side=side, image_padding=pad, image_sh_quant=2)
a = torch.randint(0, 18, (Nt,), dtype=torch.int32, device="cuda")
frames = torch.randint(0, 256, (Nt, side, side, 3), dtype=torch.uint8, device="cuda")
r = torch.rand(Nt, dtype=torch.float64, device="cuda"); d = torch.zeros(Nt, dtype=torch.uint8, device="cuda")
for _ in range(4):
    tail.actions(a); tail.post(frames, r, d)
torch.cuda.synchronize()
Nm = 1 << 20
t2 = VectorGymEnvTail(Nm, seed=0, obs_dim=17, state_space_type="continuous", delay=1, transition_noise=0.05, reward_noise=0.1)
obs = torch.rand((Nm, 17), device="cuda"); r = torch.rand(Nm, dtype=torch.float64, device="cuda"); d = torch.zeros(Nm, dtype=torch.uint8, device="cuda")
for _ in range(4): t2.post(obs, r, d)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:tail_ --csv python /tmp/tail_run.py 2>/dev/null | grep -v "^==" | tail -60 | awk -F'","' '{print $5, $(NF-2), $(NF)}' | tail -50
