"""Per-launch timing of the tensor-core ops of one training step (eager plan, CUDA events): which shapes are slow?"""
import os, sys, collections, json
os.environ["CDAE_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import FLAGS, PENDULUM, synth_batch
from causaldiffae_b200 import script_util as su, dist_util, logger, ops
from causaldiffae_b200.train_util import TrainLoop
import causaldiffae_b200.nn as cnn

torch.cuda.set_device(0)
dist_util.setup_dist(); logger.configure(dir="/tmp/cdae_op", format_strs=[])
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
for _ in range(3):
    loop.run_step(x, dict(cond))
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for rep in range(3):
    ops.PROFILE = []
    loop.run_step(x, dict(cond))
    torch.cuda.synchronize()
    for kind, info, flops, e0, e1 in ops.PROFILE:
        a = agg[(kind, info)]
        a[0] += e0.elapsed_time(e1); a[1] += flops; a[2] += 1
ops.PROFILE = None
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
tot = sum(v[0] for _, v in rows) / 3
print(f"tensor-core op time per step: {tot:.2f} ms")
for (kind, info), (ms, fl, n) in rows[:45]:
    print(f"{ms/3:7.3f} ms/step  n={n//3:3d}  {fl/ms/1e9:7.1f} TF/s  {kind:6s} {info}")
