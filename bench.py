#!/usr/bin/env python
"""bench.py — CausalDiffAE denoising hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (libcdae via the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): Pendulum-shaped synthetic 3x64x64, 4-variable causal DAG, CausalDiffAE training
(num_channels 128, 2 res blocks, attention at 16x16 / 8x8, rep_cond + causal_modeling), per-GPU batch 64, bf16 tensor-core
compute with fp32 master weights; one "step" = TrainLoop.run_step (forward + backward + fused AdamW/EMA) on one batch.
N > 1: one process per GPU under torchrun, NCCL gradient all-reduce, weak scaling (per-GPU batch fixed).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PENDULUM = [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]
FLAGS = dict(image_size=64, num_channels=128, num_res_blocks=2, num_heads=4, attention_resolutions="16,8",
             class_cond=False, rep_cond=True, n_vars=4, causal_modeling=True, in_channels=3, learn_sigma=False,
             rescale_timesteps=False, rescale_learned_sigmas=False, diffusion_steps=1000, noise_schedule="linear")
FWD_GFLOP_PER_IMG = 60.62          # SURVEY.md 8d (2*MAC, forward, cfg2/3); train = 3x forward
WORKLOAD = "pendulum64-train: 3x64x64, 4-var DAG, nc128 x2 res blocks, attn@16,8, per-GPU batch %d"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def synth_batch(B, seed, device=None, pinned=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, 64, 64, generator=g)
    c = torch.rand(B, 4, generator=g)
    if pinned:
        x, c = x.pin_memory(), c.pin_memory()
    if device is not None:
        x, c = x.to(device), c.to(device)
    return x, {"c": c}


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
CFG1 = dict(image_size=32, num_channels=64, num_res_blocks=2, class_cond=True, rep_cond=True, n_vars=2, causal_modeling=True,
            in_channels=1, learn_sigma=False, rescale_timesteps=False, rescale_learned_sigmas=False, diffusion_steps=1000,
            noise_schedule="linear")


def _dezero(model, seed=1):
    """the reference init zeroes every out conv (eps == 0): de-zero like the CUDA arm so the math is generic"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if float(p.abs().sum()) == 0.0 and p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)


def _reference_trainloop(flags, A, B, n_vars, in_channels):
    """the UNMODIFIED reference (oracle/refshim.py imports it from /root/reference or oracle/_ref; two stand-in modules for
    blobfile / mpi4py and the documented patches: encoder depth, injectable DAG) behind its own TrainLoop"""
    from oracle import refshim
    ns = refshim.load()
    torch.manual_seed(0)
    model, diff = refshim.build({**ns.su.model_and_diffusion_defaults(), **flags}, rep_dim=512, A=A)
    _dezero(model)
    model.train()
    if not torch.distributed.is_initialized():
        # a private single-process gloo group (under torchrun MASTER_PORT belongs to the launcher's store)
        import socket
        sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
        torch.distributed.init_process_group(backend="gloo", rank=0, world_size=1, init_method=f"tcp://127.0.0.1:{port}")
    ns.logger.configure(dir=os.path.join("/tmp", f"cdae_ref_{os.getpid()}"), format_strs=[])
    tl = ns.train.TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                            log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=n_vars,
                            causal_modeling=True, in_channels=in_channels)
    return ns, model, diff, tl


def _reference_cfg1_workload(cores):
    """BASELINE.json configs[0] / SURVEY 8(d): MorphoMNIST-shaped 1x32x32, 2-variable graph, 64 ch x 2 res blocks, B = 16:
    100 reference `TrainLoop.run_step` + one 10-step DDIM counterfactual through the reference's own API, on the CPU"""
    from oracle import refshim
    B = 16
    ns, model, diff, tl = _reference_trainloop(CFG1, None, B, 2, 1)
    g = torch.Generator().manual_seed(0)
    x, c, y = torch.rand(B, 1, 32, 32, generator=g), torch.rand(B, 2, generator=g), torch.randint(0, 10, (B,), generator=g)
    np.random.seed(0); torch.manual_seed(0)
    tl.run_step(x, {"c": c, "y": y})
    t0 = time.perf_counter()
    for _ in range(100):
        tl.run_step(x, {"c": c, "y": y})
    dt_train = time.perf_counter() - t0
    _, d10 = refshim.build({**ns.su.model_and_diffusion_defaults(), **CFG1, "timestep_respacing": "ddim10"}, rep_dim=512)
    model.eval()
    t0 = time.perf_counter()
    with torch.no_grad():
        mu, var = model.rep_emb.encode(x)
        var = torch.ones(var.shape) * 0.001
        mu[:, :256] = 0.2
        At = torch.tensor([[0, 1], [0, 0]], dtype=torch.float32)
        z_post = model.causal_mask.nonlinearity_add_back_noise(mu, model.causal_mask.causal_masking(mu, At))
        z = ns.nn.reparameterize(z_post, var)
        x_T = d10.q_sample(x, torch.tensor([d10.num_timesteps - 1] * B), noise=torch.randn_like(x))
        img = d10.ddim_sample_loop(model, tuple(x.shape), noise=x_T, clip_denoised=True, model_kwargs=dict(z=z, y=y))
    dt_ddim = time.perf_counter() - t0
    assert img.shape == x.shape
    return {"workload": "cfg1 morphomnist32: 100 reference TrainLoop.run_step + DDIM-10 counterfactual, batch 16, fp32",
            "train_img_per_s": B * 100 / dt_train, "train_s": dt_train, "ddim10_img_per_s": B / dt_ddim, "ddim10_s": dt_ddim,
            "cores": cores}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores - the unmodified reference
    package through its own TrainLoop when it is available (oracle/_ref on the GPU box), else the oracle port.  Same config /
    metric as the CUDA arm; each step is a bounded sample of the workload (batch --ref-batch of the per-GPU batch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refshim
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = args.ref_batch
    x, cond = synth_batch(Bs, 0)
    np.random.seed(0); torch.manual_seed(0)
    extra = {}
    if refshim.available():
        kind = "reference"
        ns, model, diff, tl = _reference_trainloop(FLAGS, PENDULUM, Bs, 4, 3)

        def step():
            tl.run_step(x, dict(cond))
    else:
        kind = "port"
        from oracle import model as om, diffusion as od, schedules
        cfg = om.config_from_flags(**FLAGS, A=PENDULUM)
        sd = om.seeded_state_dict(cfg, seed=0)
        tr = od.RefTrainer(sd, cfg, od.Diffusion(steps=1000), lr=1e-4, ema_rate=0.9999)

        def step():
            t, w = schedules.uniform_sample_t(1000, Bs)
            tr.run_step(x, torch.from_numpy(t), torch.randn_like(x), torch.from_numpy(w), c=cond["c"])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    if kind == "reference" and not args.no_cfg1:
        try:
            extra["cfg1_workload"] = _reference_cfg1_workload(cores)
        except Exception as ex:
            extra["cfg1_workload"] = {"error": repr(ex)}
    sample = (f"{args.steps} TrainLoop.run_step of the {'unmodified reference' if kind == 'reference' else 'oracle port'} at batch "
              f"{Bs} (of the per-GPU batch {args.batch}), fp32, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "train_img_per_s", "value": val, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.batch, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, **extra,
    }))


def cpu_baseline_sample(batch):
    """the reference arm (the unmodified reference on the host cores, or the oracle port without it) on a bounded sample,
    in a child process that cannot see the GPU: the reference's TrainLoop wraps the model in a CUDA DistributedDataParallel
    whenever torch.cuda.is_available() (ref train_util.py:107-118), and this baseline is its CPU path"""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "8", "--warmup", "1",
                        "--no-cfg1", "--batch", str(batch)], capture_output=True, text=True, env=env, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        raise RuntimeError("reference arm failed: " + r.stderr[-400:])
    return json.loads(line[-1])["cpu_baseline"]


def cuda_cfg1_workload(dev):
    """BASELINE.json configs[0] on this build (the reference arm times the same workload on the CPU, `cfg1_workload` there):
    MorphoMNIST-shaped 1x32x32, 2-variable graph, 64 ch x 2 res blocks, batch 16: 100 `TrainLoop.run_step` from host batches with a
    loss read back per step, then one 10-step DDIM counterfactual - wall clock, like the reference arm."""
    from causaldiffae_b200 import script_util as su
    from causaldiffae_b200.train_util import TrainLoop
    from causaldiffae_b200.sampling import counterfactual
    B = 16
    torch.manual_seed(0)
    model, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **CFG1})
    _dezero(model)
    model.to(dev)
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=2,
                     causal_modeling=True, in_channels=1)
    g = torch.Generator().manual_seed(0)
    x, c, y = torch.rand(B, 1, 32, 32, generator=g).pin_memory(), torch.rand(B, 2, generator=g).pin_memory(), \
        torch.randint(0, 10, (B,), generator=g).pin_memory()
    np.random.seed(0)
    for _ in range(4):
        loop.run_step(x, {"c": c, "y": y})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        loop.run_step(x, {"c": c, "y": y})
        float(loop.last_loss)
    torch.cuda.synchronize()
    dt_train = time.perf_counter() - t0
    _, d10 = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **CFG1, "timestep_respacing": "ddim10"})
    model.eval()
    xd, yd = x.to(dev), y.to(dev)
    counterfactual(model, d10, xd, do_var=0, do_value=0.2, y=yd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    img = counterfactual(model, d10, xd, do_var=0, do_value=-0.1, y=yd)
    img.cpu()
    dt_ddim = time.perf_counter() - t0
    return {"workload": "cfg1 morphomnist32: 100 TrainLoop.run_step (host batches, loss read per step) + DDIM-10 counterfactual, batch 16",
            "train_img_per_s": B * 100 / dt_train, "train_s": dt_train, "ddim10_img_per_s": B / dt_ddim, "ddim10_s": dt_ddim}


def gpu_reference_sample(B, steps=3):
    """the same algorithm (the oracle restatement of the reference, same weights / batch) executed by EAGER PyTorch on this
    GPU - cuDNN / cuBLAS through ATen, bf16 autocast: the 'existing Blackwell library kernels' bar this build is measured
    against on equal hardware (SURVEY 8d 'GPU reference').  Reported next to the headline, never part of it."""
    from oracle import model as om, diffusion as od, schedules
    dev = torch.device("cuda")
    cfg = om.config_from_flags(**FLAGS, A=PENDULUM)
    sd = {k: v.to(dev) for k, v in om.seeded_state_dict(cfg, seed=0).items()}
    tr = od.RefTrainer(sd, cfg, od.Diffusion(steps=1000), lr=1e-4)
    x, cond = synth_batch(B, 0, device=dev)
    np.random.seed(0); torch.manual_seed(0)

    def step():
        t, w = schedules.uniform_sample_t(1000, B)
        with torch.autocast("cuda", dtype=torch.bfloat16):
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
    del tr, sd
    torch.cuda.empty_cache()
    return {"value": B / dt, "unit": "img/s", "ms_per_step": 1000 * dt, "batch": B,
            "what": "oracle restatement of the reference, eager PyTorch (cuDNN/cuBLAS) bf16 autocast, same B200, same batch"}


# ---------------------------------------------------------------------------------------------- dominant-kernel roofline
def conv_roofline(peaks, B):
    """Live CUDA-event timing of the dominant kernel class: the 3x3 implicit-GEMM conv (tcgen05), at the layer shape
    that carries the most FLOPs in cfg2 (res.out 128->128 @ 64x64 and 256->256 @ 32x32, 10 % each).
    Inputs are rotated over buffers totalling more than L2 (126 MB) so that every launch streams from HBM."""
    from causaldiffae_b200 import ops
    dev = torch.device("cuda")
    out = {}
    for name, (C, S) in {"conv3x3_128c_64px": (128, 64), "conv3x3_256c_32px": (256, 32)}.items():
        nbuf = max(2, int(200e6 // (B * S * S * C * 2)) + 1)
        xs = [torch.randn(B, S, S, C, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
        ys = [torch.empty(B, S, S, C, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        w = (torch.randn(C, 9 * C, device=dev) * 0.02).to(torch.bfloat16)
        bias = torch.zeros(C, device=dev)
        segs, _ = ops.conv_segments([C], 3)
        descs = [ops.make_igemm_desc([x], segs, w, y, C, bias=bias) for x, y in zip(xs, ys)]
        for d in descs:
            ops.igemm(d)
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for d in descs:
                ops.igemm(d)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * nbuf)
        flops = 2.0 * B * S * S * C * 9 * C
        out[name] = dict(ms=ms, tflops=flops / ms / 1e9, flops=flops)
        try:    # the same launches as the forward pass issues them: with the GroupNorm statistics epilogue
            stats = torch.zeros(B, C, 2, device=dev)
            sdescs = [ops.make_igemm_desc([x], segs, w, y, C, bias=bias, stats=stats) for x, y in zip(xs, ys)]
            for d in sdescs:
                ops.igemm(d)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                for d in sdescs:
                    ops.igemm(d)
            e1.record(); torch.cuda.synchronize()
            out[name]["tflops_stats"] = flops / (e0.elapsed_time(e1) / (reps * nbuf)) / 1e9
        except Exception:
            pass
    top = out["conv3x3_128c_64px"]
    # dram__bytes_read + dram__bytes_write of this kernel and shape from the committed `ncu --set full` capture
    # (profiles/r2_ncu_igemm3t_summary.json): the input once (67 MB; the weights stay in L2) + the part of the 67 MB output
    # that was written back before the kernel ended; algorithmic bytes are 134.5 MB, so nothing is re-read
    traffic, tsrc = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_igemm3t_summary.json")))[0]
        traffic, tsrc = (cap["dram_read_MB"] + cap["dram_write_MB"]) * 1e6, "profiles/r2_ncu_igemm3t_summary.json"
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "igemm3t_kernel<2,5,3> (3x3 conv 128->128 @64x64, batch %d)" % B,
            "achieved": top["tflops"], "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": top["tflops"] / peaks["tf_burst"],
            "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": tsrc,
            "peak_source": peaks["src"] + " bf16 burst",
            "flops_per_launch": top["flops"],
            "other_shapes": {k: round(v["tflops"], 1) for k, v in out.items()},
            # forward launches also accumulate the consumer GroupNorm's channel sums in the epilogue (data-gradient launches
            # of the same kernel do not): TFLOP/s of that form
            "with_statistics_epilogue": {k: round(v["tflops_stats"], 1) for k, v in out.items() if "tflops_stats" in v}}


def hbm_kernels(peaks):
    """The fused HBM-bound kernels at HBM-saturating sizes (one per-GPU shard is launch-latency-scale, so the
    roofline fraction is taken at 4096x3x64x64 = 50 M elements / the 93.5 M-parameter arena): live CUDA-event timing,
    buffers rotated so that consecutive launches never hit L2.  achieved = algorithmic bytes / time."""
    from causaldiffae_b200 import ops
    dev = torch.device("cuda")
    Bn, per = 4096, 3 * 64 * 64
    n = Bn * per

    def timeit(fns, reps=4):
        for f in fns:
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for f in fns:
                f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * len(fns))

    bufs = [[torch.randn(Bn, 3, 64, 64, device=dev) for _ in range(4)] for _ in range(2)]
    t = torch.randint(0, 1000, (Bn,), device=dev)
    ta, tb = torch.rand(1000, device=dev), torch.rand(1000, device=dev)
    gs = torch.full((Bn,), 1.0 / Bn, device=dev)
    coef = torch.rand(50, 8, device=dev); coef[:, 5] = 1.0; coef[:, 6] = 0.0; coef[:, 4] = 0.0
    tidx = torch.full((1,), 7, device=dev, dtype=torch.int32)
    out = {}
    ms = timeit([lambda a=a: ops.q_sample(a[0], a[1], t, ta, tb, out=a[2]) for a in bufs])
    out["q_sample"] = (12.0 * n, ms)
    ms = timeit([lambda a=a: ops.mse_loss(a[0], a[1], gs, want_grad=True) for a in bufs])
    out["mse_fwd_bwd"] = (12.0 * n, ms)           # reads pred, target; writes the gradient
    ms = timeit([lambda a=a: ops.ddim_step(a[0], a[1], coef, tidx, out=a[2]) for a in bufs])
    out["ddim_step"] = (12.0 * n, ms)
    ms = timeit([lambda a=a: ops.ddim_step(a[0], a[1], coef, tidx, eps_u=a[3], w=1.5, out=a[2]) for a in bufs])
    out["ddim_step_guided"] = (16.0 * n, ms)
    del bufs
    P = 93_500_000 // 4 * 4
    sets = [[torch.randn(P, device=dev) * 0.01 for _ in range(5)] for _ in range(2)]
    for s_ in sets:
        s_[3].abs_()
    hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-8, 0.0, 0.9999, 1.0, 0.0], device=dev)
    gsq = torch.zeros(1, device=dev)
    astep = torch.zeros(1, device=dev, dtype=torch.int64)
    ms = timeit([lambda a=a: ops.adam_ema(a[0], a[1], a[2], a[3], a[4], hyper, astep, gsq) for a in sets])
    out["adam_ema"] = (36.0 * P, ms)
    del sets
    # GroupNorm32 + FiLM + SiLU on the heaviest cfg2 layer shape (concat 128+128 channels at 64x64, the per-GPU batch of
    # the workload: 268 MB per tensor): streaming forward (4 B/elem) and the resident cluster backward (6 B/elem)
    Bg, HW, C0, C1 = 64, 4096, 128, 128
    Cg = C0 + C1
    ng = Bg * HW * Cg
    gam, bet = torch.randn(Cg, device=dev), torch.randn(Cg, device=dev)
    film = torch.randn(Bg, 2 * Cg, device=dev) * 0.1
    gsets = []
    for _ in range(2):
        x0 = torch.randn(Bg, HW, 1, C0, device=dev).to(torch.bfloat16)
        x1 = torch.randn(Bg, HW, 1, C1, device=dev).to(torch.bfloat16)
        st0 = torch.stack([x0.float().sum(dim=(1, 2)), (x0.float() ** 2).sum(dim=(1, 2))], dim=-1).contiguous()
        st1 = torch.stack([x1.float().sum(dim=(1, 2)), (x1.float() ** 2).sum(dim=(1, 2))], dim=-1).contiguous()
        gsets.append(dict(x0=x0, x1=x1, st0=st0, st1=st1, y=torch.empty(Bg, HW, 1, Cg, device=dev, dtype=torch.bfloat16),
                          dy=torch.randn(Bg, HW, 1, Cg, device=dev).to(torch.bfloat16), dx0=torch.empty_like(x0),
                          dx1=torch.empty_like(x1), mean=torch.empty(Bg, 32, device=dev), rstd=torch.empty(Bg, 32, device=dev),
                          ))
    ms = timeit([lambda a=a: ops.gn_apply_fwd(a["x0"], a["st0"], gam, bet, x1=a["x1"], stats1=a["st1"], film=film, silu=True,
                                              out=a["y"], mean=a["mean"], rstd=a["rstd"]) for a in gsets])
    out["groupnorm_fwd_stream"] = (4.0 * ng, ms)
    dg, db, dfl = torch.zeros(Cg, device=dev), torch.zeros(Cg, device=dev), torch.zeros_like(film)

    def gbwd(a):
        ops.gn_bwd(a["dy"], a["x0"], gam, bet, a["mean"], a["rstd"], x1=a["x1"], film=film, silu=True, dx0=a["dx0"],
                   dx1=a["dx1"], dgamma=dg, dbeta=db, dfilm=dfl)
    ms = timeit([lambda a=a: gbwd(a) for a in gsets])
    out["groupnorm_bwd_resident"] = (6.0 * ng, ms)     # fallback kernel (reduces and applies in one launch)
    # the backward the training step runs: du and {sum du, sum du*x} come out of the data-gradient conv's epilogue
    # (cdae_igemm_desc.gnb_*), the norm's own backward is this one streaming pass: read du, x; write dx = 6 B / element
    ws = torch.randn(Bg, Cg, 2, device=dev)

    def gbwd2(a):
        ops.gn_bwd_apply(a["dy"], a["x0"], gam, bet, a["mean"], a["rstd"], ws, x1=a["x1"], film=film, dx0=a["dx0"],
                         dx1=a["dx1"], dgamma=dg, dbeta=db, dfilm=dfl)
    ms = timeit([lambda a=a: gbwd2(a) for a in gsets])
    out["groupnorm_bwd"] = (6.0 * ng, ms)
    return {k: {"GB/s": round(b / ms / 1e6, 1), "frac": round(b / ms / 1e6 / peaks["hbm"], 3), "ms": round(ms, 4)}
            for k, (b, ms) in out.items()} | {"peak": peaks["hbm"], "peak_source": peaks["src"] + " hbm copy",
                                              "sizes": "4096x3x64x64 fp32 tensors; 93.5 M-element arenas; GroupNorm: 64x(128+128)x64x64 bf16"}


# ---------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch.distributed as dist
    from causaldiffae_b200 import script_util as su, dist_util, logger, engine as eng_mod
    from causaldiffae_b200.train_util import TrainLoop
    import causaldiffae_b200.nn as cnn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.setup_dist()
    logger.configure(dir=os.path.join("/tmp", f"cdae_bench_{os.getpid()}"), format_strs=[])
    cnn.RNG_MODE = "device"        # throughput runs draw xi / masks on the device generator (no per-step H2D)
    peaks = load_peaks()
    B = args.batch
    torch.manual_seed(0)
    model, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **FLAGS}, A=PENDULUM)
    # de-zero the zero_module tensors (reference init makes eps == 0) so that the timed math is generic
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if float(p.abs().sum()) == 0.0 and p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * fan_in ** -0.5)
    model.to(dev)
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=4,
                     causal_modeling=True, in_channels=3)
    np.random.seed(1234 + rank)
    dev_batches = [synth_batch(B, 100 + rank * 10 + i, device=dev) for i in range(2)]
    host_batches = [synth_batch(B, 200 + rank * 10 + i, pinned=True) for i in range(3)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def step_dev(i):
        x, cond = dev_batches[i % len(dev_batches)]
        loop.run_step(x, dict(cond))

    last = {}

    def step_host(i):
        x, cond = host_batches[i % len(host_batches)]
        loop.run_step(x, dict(cond))
        last["loss"] = float(loop.last_loss)       # D2H read of the step's result

    clocks = ClockSampler(local)
    clocks.start()                  # nvidia-smi needs ~1 s before its first row (longer on an 8-GPU box): start it early
    for i in range(max(args.warmup, 3)):
        step_dev(i)
    t_wait = time.time()
    while not clocks.rows and time.time() - t_wait < 8.0:      # keep the GPU under load until the sampler is live
        step_dev(0)
    clocks.rows.clear()
    ms_step = timed(step_dev, args.steps)
    for i in range(3):
        step_host(i)
    ms_e2e = timed(step_host, args.steps)
    if not clocks.rows:             # short runs: sample a little longer under the same load
        t_wait = time.time()
        while not clocks.rows and time.time() - t_wait < 3.0:
            step_dev(0)
    clk = clocks.stop()

    # counted, not derived: kernels issued through the C ABI while the step graph was captured (+ the optimizer's two)
    fs = loop._fused.get(B)
    launches_per_step = (fs.graph_kernels + 2) if fs is not None and getattr(fs, "graph_kernels", None) else None
    value = B * world / (ms_step / 1000)
    e2e = B * world / (ms_e2e / 1000)
    train_flops = 3 * FWD_GFLOP_PER_IMG * 1e9 * B
    res = {
        "metric": "train_img_per_s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD % B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": "per-step working set (activations + 93.5M-param arenas, several GB) exceeds the 126 MB L2; no flush",
                   "graphs": eng_mod.USE_GRAPHS},
        "e2e": {"value": e2e, "unit": "img/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(B * (3 * 64 * 64 + 4) * 4 + B * 12), "d2h_bytes_per_step": 4,
                "loss": last.get("loss")},
        "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else None,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clk,
        "step_roofline": {"bound": "tensor", "achieved": train_flops / ms_step / 1e9, "peak": peaks["tf_sus"],
                          "unit": "TFLOP/s", "frac": train_flops / ms_step / 1e9 / peaks["tf_sus"],
                          "note": "algorithmic 3 x 60.62 GFLOP/img over the whole step vs " + peaks["src"] + " sustained bf16"},
    }
    if rank == 0:
        try:
            res["roofline"] = conv_roofline(peaks, B)
        except Exception as ex:   # never lose the headline number to the auxiliary measurement
            res["roofline"] = {"error": repr(ex)}
        try:
            res["hbm_kernels"] = hbm_kernels(peaks)
        except Exception as ex:
            res["hbm_kernels"] = {"error": repr(ex)}
        if world == 1 and not args.no_cfg1:
            try:
                res["cfg1_workload"] = cuda_cfg1_workload(dev)
            except Exception as ex:
                res["cfg1_workload"] = {"error": repr(ex)}
        if world == 1 and not args.no_gpu_ref:
            try:
                res["gpu_reference"] = gpu_reference_sample(B)
            except Exception as ex:
                res["gpu_reference"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu:
            try:
                res["cpu_baseline"] = cpu_baseline_sample(B)
            except Exception as ex:
                res["cpu_baseline"] = {"error": repr(ex)}
        if not args.no_ddim:
            try:
                res["ddim"] = ddim_bench(model, world, dev, args)
            except Exception as ex:
                res["ddim"] = {"error": repr(ex)}
    elif not args.no_ddim:
        try:
            ddim_bench(model, world, dev, args)
        except Exception:
            pass
    if rank == 0:
        print(json.dumps(res), flush=True)
    barrier()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


def ddim_bench(model, world, dev, args):
    """Secondary workload (BASELINE configs[3]): counterfactual encode -> do() -> DDIM decode, batch of interventions
    sharded over ranks with no communication until a final gather."""
    import torch.distributed as dist
    from causaldiffae_b200 import script_util as su
    from causaldiffae_b200.sampling import counterfactual
    steps = args.ddim_steps
    _, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), **FLAGS,
                                               "timestep_respacing": f"ddim{steps}"}, A=PENDULUM)
    Bd = args.ddim_batch
    g = torch.Generator().manual_seed(7)
    x = torch.rand(Bd, 3, 64, 64, generator=g).to(dev)
    model.eval()
    out = counterfactual(model, diff, x, do_var=0, do_value=0.2)        # warm-up (plan build + graph capture)
    out = counterfactual(model, diff, x, do_var=0, do_value=0.2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = counterfactual(model, diff, x, do_var=0, do_value=-0.35)
    if world > 1:
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    peaks = load_peaks()
    tf = steps * FWD_GFLOP_PER_IMG * 1e9 * Bd / float(ms) / 1e9
    res = {"metric": f"ddim{steps}_counterfactual_img_per_s", "value": Bd * world / (float(ms) / 1000), "unit": "img/s",
           "batch_per_gpu": Bd, "ms": float(ms), "tflops_per_gpu": tf, "frac_of_sustained_peak": tf / peaks["tf_sus"]}
    if world == 1:
        # BASELINE configs[4]: classifier-free guided sampling (w = 1.5): two UNet forwards per DDIM step + the fused combine
        try:
            counterfactual(model, diff, x, do_var=0, do_value=0.2, w=1.5)
            torch.cuda.synchronize()
            e0.record()
            counterfactual(model, diff, x, do_var=0, do_value=-0.35, w=1.5)
            e1.record()
            torch.cuda.synchronize()
            gms = e0.elapsed_time(e1)
            gtf = 2 * steps * FWD_GFLOP_PER_IMG * 1e9 * Bd / gms / 1e9
            res["guided_w1.5"] = {"value": Bd / (gms / 1000), "unit": "img/s", "ms": gms, "tflops_per_gpu": gtf,
                                  "frac_of_sustained_peak": gtf / peaks["tf_sus"]}
        except Exception as ex:
            res["guided_w1.5"] = {"error": repr(ex)}
    model.train()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--ref-batch", type=int, default=4, help="bounded per-step sample of the reference arm")
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--ddim-batch", type=int, default=512, help="interventions per GPU (BASELINE configs[3]: 4096 over 8 GPUs)")
    ap.add_argument("--no-ddim", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip the eager-PyTorch-on-this-GPU reference leg")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the cfg1 (100 steps + DDIM-10, batch 16) workload")
    args = ap.parse_args()
    if args.impl == "reference":
        # the reference's CPU path: hide the GPUs before anything initialises CUDA (its TrainLoop would otherwise build a CUDA
        # DistributedDataParallel around the CPU model).  Every step is a bounded sample (batch --ref-batch): K and W are
        # honoured as given
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
