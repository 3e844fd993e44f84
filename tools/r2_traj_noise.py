"""How far do two runs of THIS implementation drift apart on the cfg2 masking run (same init, data, t, noise, xi, masks)?
The fused step accumulates weight gradients with fp32 atomics, so it is not bit-stable from run to run; the loss-parity gate
against the oracle (tests/test_train_parity_gpu.py) cannot be tighter than this self-distance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_train_parity_gpu as T
from causaldiffae_b200 import script_util as su, dist_util, logger
from causaldiffae_b200.train_util import TrainLoop
from oracle import schedules
dev = torch.device("cuda:0")
full = {**su.model_and_diffusion_defaults(), **T.FLAGS2}
dist_util.setup_dist()
logger.configure(dir="/tmp/cdae_noise", format_strs=[])
B, STEPS = 16, 500
runs = []
for r in range(3):
    torch.manual_seed(0)
    model, diff = su.create_model_and_diffusion(**full, A=T.CIRCUIT)
    model.to(dev)
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=4,
                     causal_modeling=True, in_channels=3, masking=True)
    gen = torch.Generator().manual_seed(321)
    mine = []
    for step in range(STEPS):
        x, c = T.structured_batch64(B, gen)
        np.random.seed(2000 + step)
        t, w = schedules.uniform_sample_t(1000, B)
        noise = torch.randn(x.shape, generator=gen).to(dev)
        torch.manual_seed(7000 + step)
        loop.schedule_sampler = T._Fixed(t, w)
        loop.noise_override = noise
        loop.run_step(x.to(dev), {"c": c.to(dev)})
        loop.step += 1
        diff.kl_weight = loop.linear_kl_weight_scheduler(loop.step, 50000, 0.0, 1.0)
        mine.append(float(loop.last_loss))
    runs.append(np.array(mine))
for a in range(3):
    for b in range(a + 1, 3):
        r100 = np.abs(runs[a].reshape(-1, 100).mean(1) - runs[b].reshape(-1, 100).mean(1)) / runs[b].reshape(-1, 100).mean(1)
        r50 = np.abs(runs[a].reshape(-1, 50).mean(1) - runs[b].reshape(-1, 50).mean(1)) / runs[b].reshape(-1, 50).mean(1)
        print(f"run {a} vs run {b}: 100-step windows {np.round(r100, 4)} max {r100.max():.4f} mean {r100.mean():.4f} | 50-step max {r50.max():.4f}"
              f" | whole-run mean loss rel diff {abs(runs[a].mean() - runs[b].mean()) / runs[b].mean():.4f}")
