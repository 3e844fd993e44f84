"""Worker of tests/test_dist_gpu.py (one process per GPU under torchrun, NCCL): every rank runs ONE
`TrainLoop.forward_backward` on its own shard and checks that the exchanged gradient (what the fused optimizer consumes:
reduced arena x grad_scale) equals the MEAN over ranks of the per-rank oracle gradients (ref train_util.py:107-126 DDP)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PENDULUM = [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]
FLAGS = dict(image_size=64, num_channels=128, num_res_blocks=2, num_heads=4, attention_resolutions="16,8",
             class_cond=False, rep_cond=True, n_vars=4, causal_modeling=True, in_channels=3, learn_sigma=False,
             rescale_timesteps=False, rescale_learned_sigmas=False, diffusion_steps=1000)


def main():
    import torch.distributed as dist
    from causaldiffae_b200 import script_util as su, dist_util, logger
    from causaldiffae_b200.train_util import TrainLoop
    from oracle import model as om, diffusion as od, schedules
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dist_util.setup_dist()
    logger.configure(dir=f"/tmp/cdae_distgrad_{rank}", format_strs=[])
    full = {**su.model_and_diffusion_defaults(), **FLAGS}
    model, diff = su.create_model_and_diffusion(**full, A=PENDULUM)
    cfg = om.config_from_flags(**full, A=PENDULUM)
    sd = om.seeded_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    model.to(dev)
    B = 4
    bf16_wire = os.environ.get("CDAE_TEST_BF16_WIRE", "0") == "1"
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=4,
                     causal_modeling=True, in_channels=3)
    loop.grad_wire_dtype = torch.bfloat16 if bf16_wire else torch.float32
    diff.kl_weight = 0.3
    names = om.trainable_names(cfg)

    def shard(r):
        g = torch.Generator().manual_seed(100 + r)
        return torch.rand(B, 3, 64, 64, generator=g), torch.rand(B, 4, generator=g), torch.randn(B, 3, 64, 64, generator=g)

    # oracle gradients of EVERY rank's shard (computed locally: no communication needed for the expectation)
    odiff = od.Diffusion(steps=1000)
    odiff.kl_weight = 0.3
    osd = {k: v.to(dev) for k, v in sd.items()}
    mean_ref = {n: torch.zeros_like(osd[n]) for n in names}
    for r in range(world):
        x, c, noise = shard(r)
        for n in names:
            osd[n].grad = None
            osd[n].requires_grad_(True)
        np.random.seed(50 + r)
        t, w = schedules.uniform_sample_t(1000, B)
        torch.manual_seed(900 + r)
        terms = od.training_losses(odiff, osd, cfg, x.to(dev), torch.from_numpy(t).to(dev), noise.to(dev), c=c.to(dev))
        (terms["loss"] * torch.from_numpy(w).to(dev)).mean().backward()
        for n in names:
            mean_ref[n] += osd[n].grad / world
    # this rank's step through the public TrainLoop path (samples t with np.random, draws noise with torch)
    x, c, noise = shard(rank)
    loop.noise_override = noise.to(dev)            # the shard's noise; everything else is TrainLoop's own (fused) path
    # four times on unchanged weights and identical draws: eager, eager, graph capture + replay, replay - the last one goes
    # through the two-graph form whose early gradient range is all-reduced beside the rest of the backward
    for _ in range(4):
        np.random.seed(50 + rank)
        torch.manual_seed(900 + rank)
        loop.forward_backward(x, {"c": c})
    fs = loop._fused[B]
    assert fs.runs == 4 and (fs.graph_a is not None) == (os.environ.get("CDAE_OVERLAP_ALLREDUCE", "1") != "0")
    named = dict(model.named_parameters())
    if getattr(loop, "_reduced_grads", None) is not None:
        flat = loop._reduced_grads.float() * loop._grad_scale
        got = {n: v for n, v in loop.engine.export_state(flat).items()}
    else:
        got = {n: named[n].grad.float() * loop._grad_scale for n in names}
    gsq_ref = sum(float((mean_ref[n] ** 2).sum()) for n in names)
    tot = float(np.sqrt(sum(float(((got[n] - mean_ref[n]) ** 2).sum()) for n in names) / gsq_ref))
    errs = {n: float((got[n] - mean_ref[n]).norm() / (mean_ref[n].norm() + 1e-12)) for n in names
            if float(mean_ref[n].norm()) >= 1e-6 * np.sqrt(gsq_ref)}
    worst = max(errs, key=errs.get)
    # every rank must hold the same reduced gradient, bit for bit
    chk = torch.stack([sum(v.double().sum() for v in got.values())])
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(bool(torch.equal(gathered[0], g)) for g in gathered)
    res = dict(rank=rank, world=world, overlapped=fs.graph_a is not None, whole_rel_l2=tot, worst=worst, worst_err=errs[worst], identical_across_ranks=same,
               backend=dist.get_backend(), wire="bf16" if bf16_wire else "fp32")
    print("DISTGRAD " + json.dumps(res), flush=True)
    ok = tot < 3e-2 and errs[worst] < 8e-2 and same
    dist.barrier()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
