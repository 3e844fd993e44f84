"""State hazards of the kernel engine (bf16 operand copies, shared plan buffers, flat gradient arena) on the GPU:
every one of these used to be silent (advisor findings, round 1)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(image_size=32, num_channels=64, num_res_blocks=1, class_cond=False, rep_cond=True, n_vars=4,
           causal_modeling=True, in_channels=3, learn_sigma=False, rescale_timesteps=False,
           rescale_learned_sigmas=False, diffusion_steps=1000)


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build(seed=0):
    from causaldiffae_b200 import script_util as su
    from oracle import model as om
    full = {**su.model_and_diffusion_defaults(), **CFG}
    model, diff = su.create_model_and_diffusion(**full)
    cfg = om.config_from_flags(**full)
    sd = om.seeded_state_dict(cfg, seed=seed)
    model.load_state_dict(sd, strict=True)
    model.cuda()
    return model, diff, cfg, sd


def inputs(B=3, seed=4):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 3, 32, 32, generator=g), torch.randint(0, 1000, (B,), generator=g),
            torch.randn(B, 512, generator=g))


def test_weights_written_after_first_forward_are_repacked():
    """forward -> load_state_dict / in-place parameter writes -> forward must run on the NEW weights (the bf16 operand
    copies are refreshed), checked against the oracle on the new weights"""
    from oracle import model as om
    model, diff, cfg, sd = build(seed=0)
    x, t, z = inputs()
    model.eval()
    with torch.no_grad():
        for _ in range(3):                                    # eager, capture, replay on the first weights
            first = model(x.cuda(), t.cuda(), z=z.cuda())[0].clone()
        assert relerr(first, om.unet_forward(sd, cfg, x, t, z=z, training=False)[0]) < 3e-2
        sd2 = om.seeded_state_dict(cfg, seed=7)
        model.load_state_dict(sd2, strict=True)               # after the engine has packed and captured its graphs
        ref2 = om.unet_forward(sd2, cfg, x, t, z=z, training=False)[0]
        got2 = model(x.cuda(), t.cuda(), z=z.cuda())[0].clone()
        assert relerr(got2, ref2) < 3e-2, relerr(got2, ref2)
        assert relerr(got2, first) > 0.5
        # in-place writes through a Parameter (zero_module / a stock optimizer / master_params_to_model_params)
        from causaldiffae_b200.nn import zero_module
        zero_module(model.out[2])
        assert float(model(x.cuda(), t.cuda(), z=z.cuda())[0].abs().max()) == 0.0
        with torch.no_grad():
            model.out[2].weight.copy_(sd2["out.2.weight"]); model.out[2].bias.copy_(sd2["out.2.bias"])
        assert relerr(model(x.cuda(), t.cuda(), z=z.cuda())[0], ref2) < 3e-2


def test_two_grad_enabled_forwards_before_backward():
    """two model calls inside one loss: the second forward must not overwrite the activations the first backward needs"""
    model, diff, cfg, sd = build()
    x, t, z = inputs()
    x2 = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    model.train()
    eng = model.engine
    named = dict(model.named_parameters())

    def grads(fn):
        eng.grad_arena.zero_()
        fn()
        return {n: p.grad.detach().clone() for n, p in named.items()}

    def one(xx, scale):
        eps = model(xx.cuda(), t.cuda(), z=z.cuda())[0]
        (eps.square().mean() * scale).backward()

    def both():
        a = model(x.cuda(), t.cuda(), z=z.cuda())[0]
        b = model(x2.cuda(), t.cuda(), z=z.cuda())[0]         # same batch size, first backward still pending
        (a.square().mean() + 2.0 * b.square().mean()).backward()

    for _ in range(2):
        one(x, 1.0)                                           # warm the plan / graphs
    ga = grads(lambda: one(x, 1.0))
    gb = grads(lambda: one(x2, 2.0))
    gab = grads(both)
    tot = np.sqrt(sum(float(((gab[n] - ga[n] - gb[n]) ** 2).sum()) for n in named))
    ref = np.sqrt(sum(float(((ga[n] + gb[n]) ** 2).sum()) for n in named))
    assert tot / ref < 2e-2, tot / ref                        # fp32 atomics reorder: equal to the bf16 noise floor
    # more forwards than plan instances awaiting backward: the stale backward raises instead of using wrong activations
    from causaldiffae_b200._lib import CdaeError
    outs = [model(x.cuda(), t.cuda(), z=z.cuda())[0] for _ in range(eng.MAX_PENDING + 1)]
    with pytest.raises((CdaeError, RuntimeError)):
        outs[0].sum().backward()
    outs[-1].sum().backward()


def test_zero_grad_keeps_the_flat_gradient_arena():
    model, diff, cfg, sd = build()
    x, t, z = inputs()
    model.train()
    eng = model.engine
    eps = model(x.cuda(), t.cuda(), z=z.cuda())[0]
    eps.square().mean().backward()
    assert float(eng.grad_arena.abs().sum()) > 0
    model.zero_grad()                                          # torch's default would be set_to_none=True
    assert float(eng.grad_arena.abs().sum()) == 0.0
    for p in model.parameters():
        assert p.grad is not None and eng.owns_grad(p.grad)
    # a DAG adjacency that requires grad takes the autograd path (the fused backward has no dL/dA)
    A = torch.tensor(model.A, dtype=torch.float32, device="cuda", requires_grad=True)
    u = torch.randn(3, 512, device="cuda")
    model.causal_mask(u, A).sum().backward()
    assert A.grad is not None and float(A.grad.abs().sum()) > 0


def test_optimize_fp16_skips_the_step_on_non_finite_gradients():
    """ref train_util.py:276-290: NaN/Inf gradients -> no parameter / moment / EMA update, lg_loss_scale -= 1"""
    from causaldiffae_b200 import dist_util, logger
    from causaldiffae_b200.train_util import TrainLoop
    model, diff, cfg, sd = build()
    dist_util.setup_dist()
    logger.configure(dir="/tmp/cdae_fp16_guard", format_strs=[])
    B = 4
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=B, microbatch=-1, lr=1e-3, ema_rate="0.99,0.999",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", use_fp16=True, rep_cond=True,
                     n_vars=4, causal_modeling=True, in_channels=3)
    g = torch.Generator().manual_seed(0)
    x, c = torch.rand(B, 3, 32, 32, generator=g), torch.rand(B, 4, generator=g)
    np.random.seed(0)
    loop.run_step(x, {"c": c})
    assert loop.opt.step_count == 1
    np.testing.assert_allclose(float(loop.lg_loss_scale), 20.0 + 1e-3, rtol=1e-6)
    snap = [loop.engine.arena.clone(), loop.opt.exp_avg.clone(), loop.opt.exp_avg_sq.clone(), loop.ema_params[0][0].clone(),
            loop.ema_params[1][0].clone()]
    xb = x.clone(); xb[1, 0, 3, 3] = float("nan")
    loop.run_step(xb, {"c": c})
    now = [loop.engine.arena, loop.opt.exp_avg, loop.opt.exp_avg_sq, loop.ema_params[0][0], loop.ema_params[1][0]]
    for a, b in zip(snap, now):
        assert torch.equal(a, b)
    assert loop.opt.step_count == 1
    np.testing.assert_allclose(float(loop.lg_loss_scale), 19.0 + 1e-3, rtol=1e-6)
    loop.run_step(x, {"c": c})                                  # and training goes on
    assert loop.opt.step_count == 2 and not torch.equal(snap[0], loop.engine.arena)
    assert bool(torch.isfinite(loop.engine.arena).all())


def test_optimizer_state_dict_is_keyed_by_parameter_and_rejects_flat_arenas():
    """the moments are saved per parameter with the model's shapes (the file does not depend on the arena order, which puts
    the representation path last); a flat-arena file of an older layout is refused instead of being mis-mapped"""
    from causaldiffae_b200 import dist_util, logger
    from causaldiffae_b200.train_util import TrainLoop
    model, diff, cfg, sd = build()
    dist_util.setup_dist()
    logger.configure(dir="/tmp/cdae_optstate", format_strs=[])
    loop = TrainLoop(model=model, diffusion=diff, data=None, batch_size=4, microbatch=-1, lr=1e-3, ema_rate="0.99",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint="", rep_cond=True, n_vars=4,
                     causal_modeling=True, in_channels=3)
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        loop.run_step(torch.rand(4, 3, 32, 32, generator=g).cuda(), {"c": torch.rand(4, 4, generator=g).cuda()})
    eng = loop.engine
    assert 0 < eng.early_end < eng.n_params
    late = [n for n, p in model.named_parameters() if eng.param_offsets[id(p)] >= eng.early_end]
    assert late and all(n.split(".")[0] in ("time_embed", "rep_emb", "up_emb", "causal_mask") for n in late)
    osd = loop.opt.state_dict()
    names = dict(model.named_parameters())
    assert set(osd["exp_avg"]) == set(names) and all(osd["exp_avg"][n].shape == names[n].shape for n in names)
    m0, v0, step0 = loop.opt.exp_avg.clone(), loop.opt.exp_avg_sq.clone(), loop.opt.step_count
    assert float(m0.abs().sum()) > 0
    loop.opt.exp_avg.zero_(); loop.opt.exp_avg_sq.zero_(); loop.opt.step_dev.zero_()
    loop.opt.load_state_dict(osd)
    assert torch.equal(loop.opt.exp_avg, m0) and torch.equal(loop.opt.exp_avg_sq, v0) and loop.opt.step_count == step0
    with pytest.raises(ValueError):
        loop.opt.load_state_dict(dict(step=3, param_groups=osd["param_groups"], exp_avg=m0, exp_avg_sq=v0))
