"""The fused training step (train_util.FusedStep: one CUDA graph of hand-written kernels, no autograd, no ATen compute) against
(a) the generic autograd path of the same package on the same draws and (b) the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PENDULUM = [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]
CFGS = {
    "cfg1-classcond": (dict(image_size=32, num_channels=64, num_res_blocks=2, class_cond=True, rep_cond=True, n_vars=2,
                            causal_modeling=True, in_channels=1), None),
    "cfg2s-masking": (dict(image_size=64, num_channels=64, num_res_blocks=1, class_cond=False, rep_cond=True, n_vars=4,
                           causal_modeling=True, in_channels=3, masking=True), PENDULUM),
    "norep": (dict(image_size=32, num_channels=64, num_res_blocks=1, class_cond=True, rep_cond=False, n_vars=4,
                   causal_modeling=False, in_channels=3), None),
}
COMMON = dict(learn_sigma=False, rescale_timesteps=False, rescale_learned_sigmas=False, diffusion_steps=1000)


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def make_loop(flags, A, sd, fused, lr=1e-4):
    from causaldiffae_b200 import script_util as su, dist_util, logger
    from causaldiffae_b200.train_util import TrainLoop
    full = {**su.model_and_diffusion_defaults(), **COMMON, **flags}
    model, diff = su.create_model_and_diffusion(**full, A=A)
    model.load_state_dict(sd, strict=True)
    model.cuda()
    dist_util.setup_dist()
    logger.configure(dir="/tmp/cdae_fused_test", format_strs=[])
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=8, microbatch=-1, lr=lr, ema_rate="0.9999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=flags["rep_cond"],
                     n_vars=flags["n_vars"], causal_modeling=flags["causal_modeling"], in_channels=flags["in_channels"],
                     masking=flags.get("masking", False))
    loop.use_fused = loop.use_fused and fused
    return loop, model, diff


@pytest.mark.parametrize("name", list(CFGS))
def test_fused_step_vs_autograd_path_and_oracle(name):
    from causaldiffae_b200 import script_util as su, logger
    from oracle import model as om, diffusion as od, schedules
    flags, A = CFGS[name]
    full = {**su.model_and_diffusion_defaults(), **COMMON, **flags}
    cfg = om.config_from_flags(**full, A=A)
    sd = om.seeded_state_dict(cfg, seed=0)
    B, S, C = 8, flags["image_size"], flags["in_channels"]      # 8: a batch whose keep mask is all zero is 0/0 (as in the reference)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, C, S, S, generator=g)
    cond = {}
    if flags["rep_cond"]:
        cond["c"] = torch.rand(B, flags["n_vars"], generator=g)
    if flags["class_cond"]:
        cond["y"] = torch.randint(0, 10, (B,), generator=g)
    results = {}
    for fused in (False, True):
        loop, model, diff = make_loop(flags, A, sd, fused)
        assert loop.use_fused == fused
        diff.kl_weight = 0.3
        losses, grads = [], None
        for step in range(5):          # fused: eager, eager, capture, replay, replay
            np.random.seed(10 + step); torch.manual_seed(20 + step)
            if step == 0:
                loop.forward_backward(x.cuda(), {k: v.cuda() for k, v in cond.items()})
                grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
                loop.optimize_normal()
            else:
                loop.run_step(x.cuda(), {k: v.cuda() for k, v in cond.items()})
            losses.append(float(loop.last_loss))
        kv = logger.getkvs()
        results[fused] = dict(losses=losses, grads=grads, arena=loop.engine.arena.clone(), kv=kv,
                              bn=[b.clone() for b in model.buffers()])
    a, f = results[False], results[True]
    print(name, "losses autograd", np.round(a["losses"], 5), "fused", np.round(f["losses"], 5))
    np.testing.assert_allclose(f["losses"], a["losses"], rtol=2e-2)
    gtot = np.sqrt(sum(float((v.float() ** 2).sum()) for v in a["grads"].values()))
    diff_tot = np.sqrt(sum(float(((f["grads"][n] - a["grads"][n]).float() ** 2).sum()) for n in a["grads"]))
    assert diff_tot / gtot < 2e-2, diff_tot / gtot            # two bf16 runs with fp32 atomics in different orders
    assert relerr(f["arena"], a["arena"]) < 1e-3
    for ba, bf in zip(a["bn"], f["bn"]):
        assert relerr(bf, ba) < 1e-4
    for key in ("loss", "mse", "grad_norm") + (("kld_rep",) if flags["rep_cond"] else ()):
        np.testing.assert_allclose(f["kv"][key], a["kv"][key], rtol=2e-2, err_msg=key)
    assert any(k.startswith("loss_q") for k in f["kv"]) and set(k for k in a["kv"] if k.startswith("mse_q")) == \
        set(k for k in f["kv"] if k.startswith("mse_q"))
    # (b) the oracle on the draws of step 0
    np.random.seed(10); torch.manual_seed(20)
    t, w = schedules.uniform_sample_t(1000, B)
    noise = torch.randn(x.shape, device="cuda").cpu()
    odiff = od.Diffusion(steps=1000)
    odiff.kl_weight = 0.3
    names = om.trainable_names(cfg)
    for n in names:
        sd[n].requires_grad_(True)
    ref = od.training_losses(odiff, sd, cfg, x, torch.from_numpy(t), noise, y=cond.get("y"), c=cond.get("c"),
                             rep_cond=flags["rep_cond"])
    lref = (ref["loss"] * torch.from_numpy(w)).mean()
    lref.backward()
    np.testing.assert_allclose(f["losses"][0], float(lref), rtol=2e-2)
    gref = np.sqrt(sum(float((sd[n].grad ** 2).sum()) for n in names))
    worst = 0.0
    for n in names:
        if float(sd[n].grad.norm()) < 1e-6 * gref:
            continue
        worst = max(worst, relerr(f["grads"][n], sd[n].grad))
    print(name, "fused step vs oracle: worst per-tensor gradient rel L2", worst)
    assert worst < 4e-2, worst


def test_two_graph_step_with_early_gradient_exchange_equals_the_single_graph():
    """multi-GPU form of the fused step on ONE rank (use_ddp forced: the all-reduces are no-ops): part "a" | early-range
    exchange on the communication stream beside part "b" | tail exchange | optimizer - against the single-graph step"""
    from oracle import model as om
    from causaldiffae_b200 import script_util as su
    flags, A = CFGS["cfg2s-masking"]
    full = {**su.model_and_diffusion_defaults(), **COMMON, **flags}
    cfg = om.config_from_flags(**full, A=A)
    sd = om.seeded_state_dict(cfg, seed=0)
    B, S, C = 8, flags["image_size"], flags["in_channels"]
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, C, S, S, generator=g)
    cond = {"c": torch.rand(B, flags["n_vars"], generator=g)}
    res = {}
    for ddp in (False, True):
        for wire in ((torch.bfloat16, torch.float32) if ddp else (torch.float32,)):
            loop, model, diff = make_loop(flags, A, sd, True)
            loop.use_ddp, loop.grad_wire_dtype = ddp, wire
            diff.kl_weight = 0.3
            losses = []
            for step in range(6):          # eager, eager, capture + replay, replay ...
                np.random.seed(10 + step); torch.manual_seed(20 + step)
                loop.run_step(x.cuda(), {k: v.cuda() for k, v in cond.items()})
                losses.append(float(loop.last_loss))
            fs = loop._fused[B]
            assert (fs.graph_a is not None) == ddp and (fs.graph is not None) == (not ddp)
            assert 0 < loop.engine.early_end < loop.engine.n_params
            res[(ddp, wire)] = (losses, loop.engine.arena.clone(), loop.ema_params[0][0].clone())
    base = res[(False, torch.float32)]
    for key in ((True, torch.float32), (True, torch.bfloat16)):
        got = res[key]
        np.testing.assert_allclose(got[0], base[0], rtol=5e-3)
        tol = 2e-3 if key[1] == torch.bfloat16 else 5e-4      # bf16 wire: the gradient is rounded once more before AdamW
        assert relerr(got[1], base[1]) < tol and relerr(got[2], base[2]) < tol, (key, relerr(got[1], base[1]))
