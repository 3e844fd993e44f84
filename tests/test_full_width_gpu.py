"""Parity at the BENCHED configuration (BASELINE.json configs[1]/[2]: 3x64x64, nc128 x 2 res blocks, attention at 16x16 /
8x8, 4 heads, 93.5 M parameters) against the oracle run eagerly in fp32 (TF32 off) ON THE GPU - the CPU oracle is too slow
at this width, the restated algorithm is device-agnostic (oracle/diffusion.py header).  Pendulum and Circuit DAGs.

  * eps given z, end to end                                   <= 3e-2  (PyTorch's own bf16 autocast of the oracle is
    measured next to it as the yardstick and printed)
  * teacher-forced forward AND backward of one layer per distinct (kind, Cin, Cout, H) of SURVEY App. A.1 - concat inputs
    1024 / 896 / 768 / 640 / 512 / 384 / 256, 1x1 skips, up / down convs, both attention shapes - through the same
    planned kernels (statistics epilogue, streaming GroupNorm, two-source K loop) the full torso launches   <= 1e-2
  * training_losses: mse / kld_rep and the gradient of every parameter tensor
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

PENDULUM = [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]
CIRCUIT = [[0, 1, 1, 1], [0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 0, 0]]
CFG2_FULL = dict(image_size=64, num_channels=128, num_res_blocks=2, num_heads=4, attention_resolutions="16,8",
                 class_cond=False, rep_cond=True, n_vars=4, causal_modeling=True, in_channels=3, learn_sigma=False,
                 rescale_timesteps=False, rescale_learned_sigmas=False, diffusion_steps=1000)


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def relerr(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build(flags, A, seed=0):
    from causaldiffae_b200 import script_util as su
    from oracle import model as om, diffusion as od
    full = {**su.model_and_diffusion_defaults(), **flags}
    model, diff = su.create_model_and_diffusion(**full, A=A)
    cfg = om.config_from_flags(**full, A=A)
    sd = om.seeded_state_dict(cfg, seed=seed)
    model.load_state_dict(sd, strict=True)
    model.cuda()
    sd = {k: v.cuda() for k, v in sd.items()}
    odiff = od.Diffusion(steps=full["diffusion_steps"], timestep_respacing=full["timestep_respacing"])
    return model, diff, cfg, sd, odiff


def inputs(B, seed=5):
    g = torch.Generator().manual_seed(seed)
    d = dict(x0=torch.rand(B, 3, 64, 64, generator=g), noise=torch.randn(B, 3, 64, 64, generator=g),
             t=torch.randint(0, 1000, (B,), generator=g), c=torch.rand(B, 4, generator=g),
             z=torch.randn(B, 512, generator=g), w=torch.rand(B, generator=g) + 0.5)
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("A", [PENDULUM, CIRCUIT], ids=["pendulum", "circuit"])
def test_full_width_eps_vs_oracle(A):
    from oracle import model as om
    model, diff, cfg, sd, odiff = build(CFG2_FULL, A)
    inp = inputs(4)
    model.eval()
    with torch.no_grad():
        x_t = odiff.q_sample(inp["x0"], inp["t"], inp["noise"])
        ref = om.unet_forward(sd, cfg, x_t, inp["t"], z=inp["z"], training=False)[0]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = om.unet_forward(sd, cfg, x_t, inp["t"], z=inp["z"], training=False)[0]
        yard = relerr(ac, ref)
        for rep in range(3):   # eager, graph capture, graph replay
            got = model(x_t, inp["t"], z=inp["z"])[0]
            err = relerr(got, ref)
            print(f"full-width eps rel L2 vs fp32 oracle: {err:.3e} (torch bf16 autocast of the oracle: {yard:.3e})")
            assert err < 2e-2 and err < 1.5 * yard, (rep, err, yard)       # measured 9.0e-3 vs 1.37e-2 for torch autocast


def layer_cases(model):
    """one (prefix, module, kind, [(channels, resolution) per input]) per distinct shape of the torso"""
    from causaldiffae_b200.unet import ResBlock, AttentionBlock, Upsample, Downsample
    cases, seen = [], set()

    def add(prefix, mod, kind, ins, cout):
        key = (kind, tuple(ins), cout)
        if key not in seen:
            seen.add(key)
            cases.append((prefix, mod, kind, ins))

    def walk(blocks, name, ch, res, skips=None, pop=False):
        for i, blk in blocks:
            skip = skips.pop() if pop else None
            for j, mod in enumerate(blk):
                prefix = f"{name}.{i}.{j}." if name != "middle_block" else f"middle_block.{j}."
                if isinstance(mod, ResBlock):
                    ins = [(ch, res)] + ([skip] if skip is not None else [])
                    skip = None
                    add(prefix, mod, "res", ins, mod.out_channels)
                    ch = mod.out_channels
                elif isinstance(mod, AttentionBlock):
                    add(prefix, mod, "attn", [(ch, res)], ch)
                elif isinstance(mod, Downsample):
                    add(prefix, mod, "down", [(ch, res)], ch)
                    res //= 2
                elif isinstance(mod, Upsample):
                    add(prefix, mod, "up", [(ch, res)], ch)
                    res *= 2
            if skips is not None and not pop:
                skips.append((ch, res))
        return ch, res

    mc, S = model.model_channels, model.image_size
    skips = [(mc, S)]
    ch, res = walk(list(enumerate(model.input_blocks))[1:], "input_blocks", mc, S, skips)
    ch, res = walk([(0, model.middle_block)], "middle_block", ch, res)
    walk(list(enumerate(model.output_blocks)), "output_blocks", ch, res, skips, pop=True)
    return cases


@pytest.mark.parametrize("fused_gn_bwd", [False, True], ids=["gn-bwd-resident", "gn-bwd-from-dgrad-epilogue"])
def test_full_width_per_layer_forward_and_backward(fused_gn_bwd, monkeypatch):
    """north_star: per-layer relative L2 <= 1e-2 in bf16 - forward output, input gradient(s), FiLM gradient and every
    parameter gradient of the layer, teacher-forced (both sides get the same fp32 input / output gradient).  Both GroupNorm
    backward variants: the resident reduce-and-apply kernel (default) and the statistics from the data-gradient epilogue."""
    from causaldiffae_b200 import engine as eng_mod
    from causaldiffae_b200.engine import run_layer_train
    from oracle import model as om
    monkeypatch.setattr(eng_mod, "FUSED_GN_BWD", fused_gn_bwd)
    model, diff, cfg, sd, odiff = build(CFG2_FULL, PENDULUM)
    model.train()
    eng = model.engine
    named = dict(model.named_parameters())
    g = torch.Generator().manual_seed(11)
    B = 2
    emb = torch.randn(B, 512, generator=g).cuda()
    cases = layer_cases(model)
    kinds = {}
    worst = dict(out=0.0, dx=0.0, dparam=0.0, dfilm=0.0)
    report = []
    for prefix, mod, kind, ins in cases:
        kinds[kind] = kinds.get(kind, 0) + 1
        res = ins[0][1]
        xs = [torch.randn(B, c, r, r, generator=g).cuda() for c, r in ins]
        # oracle: fp32 autograd on the concatenated input
        xo = torch.cat(xs, dim=1).clone().requires_grad_(True)
        pnames = [n for n in named if n.startswith(prefix)]
        leaves = {n: sd[n].clone().requires_grad_(True) for n in pnames}
        sdl = {**sd, **leaves}
        if kind == "res":
            ref = om.resblock(sdl, prefix, xo, emb)
        elif kind == "attn":
            ref = om.attention_block(sdl, prefix, xo, mod.num_heads)
        elif kind == "down":
            ref = om.downsample(sdl, prefix, xo)
        else:
            ref = om.upsample(sdl, prefix, xo)
        dout = torch.randn(ref.shape, generator=g).cuda()
        ref.backward(dout)
        # CUDA path
        eng.grad_arena.zero_()
        out, dxs, dfilm = run_layer_train(mod, xs if len(xs) > 1 else xs[0], emb=emb, dout=dout)
        e_out = relerr(out, ref)
        dx_ref = torch.split(xo.grad, [c for c, _ in ins], dim=1)
        e_dx = max(relerr(a, b) for a, b in zip(dxs, dx_ref))
        e_par, worst_name = 0.0, ""
        for n in pnames:
            if ".emb_layers." in n:
                continue
            e = relerr(named[n].grad, leaves[n].grad)
            if e > e_par:
                e_par, worst_name = e, n[len(prefix):]
        e_film = 0.0
        if kind == "res":
            # d(emb_layers output): compare through the two gradients it feeds (bias: column sums, weight: outer product)
            foff, width = eng.film_off[id(mod)], 2 * mod.out_channels
            de = dfilm[:, foff:foff + width]
            e_film = max(relerr(de.sum(0), leaves[prefix + "emb_layers.1.bias"].grad),
                         relerr(de.t() @ F.silu(emb), leaves[prefix + "emb_layers.1.weight"].grad))
        report.append((prefix, kind, ins, e_out, e_dx, e_par, worst_name, e_film))
        worst = dict(out=max(worst["out"], e_out), dx=max(worst["dx"], e_dx), dparam=max(worst["dparam"], e_par),
                     dfilm=max(worst["dfilm"], e_film))
    for r in report:
        print("%-24s %-5s %-28s out %.2e  dx %.2e  dparam %.2e (%s)  dfilm %.2e" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]))
    print("layers:", kinds, "worst:", worst)
    assert kinds.get("res", 0) >= 15 and kinds.get("attn", 0) >= 2 and kinds.get("up", 0) == 3 and kinds.get("down", 0) == 3
    bad = [r for r in report if max(r[3], r[4], r[5], r[7]) > 1e-2]
    assert not bad, bad


@pytest.mark.parametrize("A,masking", [(PENDULUM, False), (CIRCUIT, True)], ids=["pendulum", "circuit-masking"])
def test_full_width_training_losses_and_gradients(A, masking):
    from oracle import model as om, diffusion as od
    flags = {**CFG2_FULL, "masking": masking}
    model, diff, cfg, sd, odiff = build(flags, A)
    inp = inputs(4)
    diff.kl_weight = odiff.kl_weight = 0.3
    names = om.trainable_names(cfg)
    for n in names:
        sd[n].requires_grad_(True)
    torch.manual_seed(21)
    ref = od.training_losses(odiff, sd, cfg, inp["x0"], inp["t"], inp["noise"], c=inp["c"])
    (ref["loss"] * inp["w"]).mean().backward()
    gref = {n: sd[n].grad.clone() for n in names}
    # yardstick: the same oracle under torch's bf16 autocast (what "bf16" costs an eager PyTorch user)
    for n in names:
        sd[n].grad = None
    torch.manual_seed(21)
    for k in list(sd):          # BatchNorm running statistics were advanced by the first pass: irrelevant in training mode
        pass
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ac = od.training_losses(odiff, sd, cfg, inp["x0"], inp["t"], inp["noise"], c=inp["c"])
        (ac["loss"] * inp["w"]).mean().backward()
    gsq_ref = sum(float((g ** 2).sum()) for g in gref.values())
    big = [n for n in names if float(gref[n].norm()) >= 1e-6 * np.sqrt(gsq_ref)]
    yard = {n: relerr(sd[n].grad, gref[n]) for n in big}
    model.train()
    eng = model.engine
    named = dict(model.named_parameters())
    for rep in range(3):           # eager, capture, replay
        eng.grad_arena.zero_()
        torch.manual_seed(21)
        terms = diff.training_losses(model, inp["x0"], inp["t"], model_kwargs=dict(c=inp["c"]), noise=inp["noise"],
                                     rep_cond=True, causal_modeling=True)
        (terms["loss"] * inp["w"]).mean().backward()
        assert relerr(terms["mse"], ref["mse"]) < 2e-2, rep
        assert relerr(terms["kld_rep"], ref["kld_rep"]) < 1e-3, rep
        gsq = sum(float((p.grad.float() ** 2).sum()) for p in named.values())
        assert abs(np.sqrt(gsq) / np.sqrt(gsq_ref) - 1) < 2e-2, (rep, gsq, gsq_ref)
        errs = {n: relerr(named[n].grad, gref[n]) for n in big}
        worst = max(errs, key=errs.get)
        wy = max(yard, key=yard.get)
        tot = float(np.sqrt(sum(float(((named[n].grad - gref[n]) ** 2).sum()) for n in names) / gsq_ref))
        print(f"rep {rep}: whole-gradient rel L2 {tot:.3e}; worst tensor {worst} {errs[worst]:.3e} "
              f"(median {np.median(list(errs.values())):.3e}); torch-autocast yardstick worst {wy} {yard[wy]:.3e} "
              f"(median {np.median(list(yard.values())):.3e})")
        # measured on B200 (profiles/r2_tests_gpu_106pass.log): whole gradient 1.2e-4, worst tensor 1.1e-2, median 7e-3 -
        # below what torch's own bf16 autocast of the reference reaches (worst 1.5e-1, median 1.2e-2)
        assert tot < 1e-3, (rep, tot)
        assert errs[worst] < 2.5e-2, (rep, worst, errs[worst])
