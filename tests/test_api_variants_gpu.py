"""Reference paths outside the shipped launch lines (live code in the reference, ref unet.py:157,190-197):
additive timestep conditioning (`use_scale_shift_norm=False`: h = GN(conv1(.) + emb_out)) and `dropout > 0` between SiLU and
the ResBlock's second convolution.  Both against the oracle; dropout with the masks of the implementation under test injected
into the oracle (torch's own dropout stream cannot be reproduced by another generator)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BASE = dict(image_size=32, num_channels=64, num_res_blocks=1, class_cond=False, rep_cond=True, n_vars=4, causal_modeling=True,
            in_channels=3, learn_sigma=False, rescale_timesteps=False, rescale_learned_sigmas=False, diffusion_steps=1000)


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build(flags, seed=0):
    from causaldiffae_b200 import script_util as su
    from oracle import model as om, diffusion as od
    full = {**su.model_and_diffusion_defaults(), **flags}
    model, diff = su.create_model_and_diffusion(**full)
    cfg = om.config_from_flags(**full)
    sd = om.seeded_state_dict(cfg, seed=seed)
    model.load_state_dict(sd, strict=True)
    model.cuda()
    return model, diff, cfg, sd, od.Diffusion(steps=1000)


def inputs(B=4, seed=5):
    g = torch.Generator().manual_seed(seed)
    return dict(x0=torch.rand(B, 3, 32, 32, generator=g), noise=torch.randn(B, 3, 32, 32, generator=g),
                t=torch.randint(0, 1000, (B,), generator=g), c=torch.rand(B, 4, generator=g),
                z=torch.randn(B, 512, generator=g), w=torch.rand(B, generator=g) + 0.5)


def test_additive_timestep_conditioning_vs_oracle():
    from oracle import model as om, diffusion as od
    model, diff, cfg, sd, odiff = build({**BASE, "use_scale_shift_norm": False})
    assert model.input_blocks[1][0].emb_layers[1].weight.shape[0] == 64          # Linear(emb, Cout), not 2*Cout
    inp = inputs()
    model.eval()
    with torch.no_grad():
        x_t = odiff.q_sample(inp["x0"], inp["t"], inp["noise"])
        ref = om.unet_forward(sd, cfg, x_t, inp["t"], z=inp["z"], training=False)[0]
        for rep in range(3):
            got = model(x_t.cuda(), inp["t"].cuda(), z=inp["z"].cuda())[0]
            assert relerr(got, ref) < 3e-2, (rep, relerr(got, ref))
    diff.kl_weight = odiff.kl_weight = 0.3
    names = om.trainable_names(cfg)
    for n in names:
        sd[n].requires_grad_(True)
    torch.manual_seed(21)
    r = od.training_losses(odiff, sd, cfg, inp["x0"], inp["t"], inp["noise"], c=inp["c"])
    (r["loss"] * inp["w"]).mean().backward()
    model.train()
    named = dict(model.named_parameters())
    gref = np.sqrt(sum(float((sd[n].grad ** 2).sum()) for n in names))
    for rep in range(3):
        model.engine.grad_arena.zero_()
        torch.manual_seed(21)
        terms = diff.training_losses(model, inp["x0"].cuda(), inp["t"].cuda(), model_kwargs=dict(c=inp["c"].cuda()),
                                     noise=inp["noise"].cuda(), rep_cond=True, causal_modeling=True)
        (terms["loss"] * inp["w"].cuda()).mean().backward()
        assert relerr(terms["mse"], r["mse"]) < 2e-2
        worst = max(relerr(named[n].grad, sd[n].grad) for n in names if float(sd[n].grad.norm()) >= 1e-6 * gref)
        assert worst < 5e-2, (rep, worst)
    # the gradient of the additive embedding path in particular (per-image column sums of d conv1-output)
    for n in names:
        if ".emb_layers.1." in n:
            assert relerr(named[n].grad, sd[n].grad) < 3e-2, n


def test_dropout_train_masks_and_eval_identity():
    from causaldiffae_b200 import ops
    from causaldiffae_b200.engine import run_layer_train
    from oracle import model as om
    p = 0.3
    model, diff, cfg, sd, odiff = build({**BASE, "dropout": p})
    model0, *_ = build({**BASE, "dropout": 0.0})
    inp = inputs()
    # eval mode: nn.Dropout is the identity
    model.eval(); model0.eval()
    with torch.no_grad():
        a = model(inp["x0"].cuda(), inp["t"].cuda(), z=inp["z"].cuda())[0]
        b = model0(inp["x0"].cuda(), inp["t"].cuda(), z=inp["z"].cuda())[0]
    assert relerr(a, b) < 2e-2
    # one ResBlock, teacher-forced, training mode: regenerate the masks the kernel drew and hand them to the oracle
    model.train()
    rb, prefix = model.input_blocks[1][0], "input_blocks.1.0."
    g = torch.Generator().manual_seed(2)
    B = 3
    x = torch.randn(B, 64, 32, 32, generator=g).cuda()
    emb = torch.randn(B, 256, generator=g).cuda()
    dout = torch.randn(B, 64, 32, 32, generator=g).cuda()
    model.engine.grad_arena.zero_()
    out, dxs, dfilm = run_layer_train(rb, x, emb=emb, dout=dout, dropout_p=p)
    pl = run_layer_train.last_plan
    ones = torch.ones(B, 32, 32, 64, device="cuda", dtype=torch.bfloat16)
    ops.dropout_(ones, pl.drop_state, 0, pl.drop_p)
    mask = ones.float().permute(0, 3, 1, 2).contiguous()                 # 0 or 1/(1-p)
    keep = float((mask > 0).float().mean())
    assert abs(keep - (1 - p)) < 0.01 and abs(float(mask.max()) - 1 / (1 - p)) < 1e-2
    sdg = {k: v.cuda() for k, v in sd.items()}
    pn = [n for n in sdg if n.startswith(prefix)]
    leaves = {n: sdg[n].clone().requires_grad_(True) for n in pn}
    xo = x.clone().requires_grad_(True)
    ref = om.resblock({**sdg, **leaves}, prefix, xo, emb, drop=mask)
    ref.backward(dout)
    assert relerr(out, ref) < 1e-2 and relerr(dxs[0], xo.grad) < 1e-2, (relerr(out, ref), relerr(dxs[0], xo.grad))
    named = dict(model.named_parameters())
    for n in pn:
        if ".emb_layers." not in n:
            assert relerr(named[n].grad, leaves[n].grad) < 1.5e-2, (n, relerr(named[n].grad, leaves[n].grad))
    # a second forward of the same plan draws new masks; a full training step stays finite
    pl._run_fwd_eager()
    ones2 = torch.ones_like(ones)
    ops.dropout_(ones2, pl.drop_state, 0, pl.drop_p)
    assert float((ones2.float() != ones.float()).float().mean()) > 0.2
    model.engine.grad_arena.zero_()
    terms = diff.training_losses(model, inp["x0"].cuda(), inp["t"].cuda(), model_kwargs=dict(c=inp["c"].cuda()),
                                 noise=inp["noise"].cuda(), rep_cond=True, causal_modeling=True)
    terms["loss"].mean().backward()
    assert bool(torch.isfinite(model.engine.grad_arena).all()) and float(model.engine.grad_arena.abs().sum()) > 0
