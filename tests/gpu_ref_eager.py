"""Reference algorithm (oracle restatement) executed by eager PyTorch ON THE SAME B200: the 'existing Blackwell
library kernels' bar of SURVEY.md 2.2 (cuDNN/cuBLAS via ATen).  Test/measurement tooling only (imports oracle/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import model as om, diffusion as od, schedules
from bench import FLAGS, PENDULUM, synth_batch


def run(B, mode, steps=5):
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    dev = torch.device("cuda")
    cfg = om.config_from_flags(**FLAGS, A=PENDULUM)
    sd = {k: v.to(dev) for k, v in om.seeded_state_dict(cfg, seed=0).items()}
    diff = od.Diffusion(steps=1000)
    tr = od.RefTrainer(sd, cfg, diff, lr=1e-4)
    x, cond = synth_batch(B, 0, device=dev)
    np.random.seed(0); torch.manual_seed(0)

    def step():
        t, w = schedules.uniform_sample_t(1000, B)
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16" else torch.autocast("cuda", enabled=False)
        with ctx:
            # RefTrainer.run_step does forward/backward/AdamW/EMA exactly like the reference TrainLoop
            tr.run_step(x, torch.from_numpy(t).to(dev), torch.randn_like(x), torch.from_numpy(w).to(dev), c=cond["c"],
                        xi=torch.randn(B, 512, device=dev))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dict(mode=mode, batch=B, ms_per_step=1000 * dt, img_per_s=B / dt)


if __name__ == "__main__":
    out = [run(64, m) for m in ("fp32", "tf32", "bf16")]
    print(json.dumps({"eager_torch_reference_on_b200": out}))
