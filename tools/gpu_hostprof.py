"""Where does a training step spend its time? host-issue time (no syncs) vs device time per section."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import FLAGS, PENDULUM, synth_batch
from causaldiffae_b200 import script_util as su, dist_util, logger
from causaldiffae_b200.train_util import TrainLoop
import causaldiffae_b200.nn as cnn

torch.cuda.set_device(0)
dist_util.setup_dist(); logger.configure(dir="/tmp/cdae_hp", format_strs=[])
cnn.RNG_MODE = "device"
B = 64
model, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **FLAGS}, A=PENDULUM)
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for n, p in model.named_parameters():
        if float(p.abs().sum()) == 0.0 and p.dim() > 1:
            p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)
model.cuda()
loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                 log_interval=10**9, save_interval=10**9, resume_checkpoint="", rep_cond=True, n_vars=4, causal_modeling=True)
loop.log_quartiles = False
x, cond = synth_batch(B, 0, device=torch.device("cuda"))
for _ in range(4):
    loop.run_step(x, dict(cond))
torch.cuda.synchronize()

# (1) pure host issue time vs total
t0 = time.perf_counter()
for _ in range(10):
    loop.run_step(x, dict(cond))
t_issue = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / 10
print(json.dumps({"host_issue_ms": 1e3 * t_issue, "total_ms": 1e3 * t_total}))

# (2) per-section with syncs
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        loop.run_step(x, dict(cond))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
