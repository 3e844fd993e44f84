"""Kernel-level parity of the fp32 representation path (csrc/rep.cu through the C ABI) against plain fp32 torch / the
oracle: the SGEMM with its operand views, the conv encoder forward + backward (batch-statistics BatchNorm), the embedding
trunk + FiLM projections, the latent kernels (reparameterisation, keep mask, closed-form KL), the loss assembly and the
counter-based generator.  Tolerance: north_star's fp32 bound, relative L2 <= 1e-4 (accumulation order differs)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_sgemm_operand_views():
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(0)
    dev = torch.device("cuda")
    for (M, N, K) in [(64, 5632, 512), (5632, 512, 64), (64, 512, 5632), (7, 33, 19), (1, 512, 64), (130, 70, 300)]:
        x, W, b = torch.randn(M, K, generator=g).to(dev), torch.randn(N, K, generator=g).to(dev), torch.randn(N, generator=g).to(dev)
        ref = F.linear(x.double(), W.double(), b.double())
        out = torch.empty(M, N, device=dev)
        ops.linear_fwd(x, W, b, out)
        assert relerr(out, ref) < 2e-6, (M, N, K)
        ops.linear_fwd(x, W, b, out, silu_in=True)
        assert relerr(out, F.linear(F.silu(x.double()), W.double(), b.double())) < 2e-6
        ops.linear_fwd(x, W, None, out, accumulate=True)                     # += through atomics
        assert relerr(out, F.linear(F.silu(x.double()), W.double(), b.double()) + x.double() @ W.double().t()) < 2e-6
        ops.linear_fwd(x, W, b, out, act_out=1)
        assert relerr(out, F.softplus(ref) + 1e-8) < 2e-6
        # backward of y = silu(x) W^T + b
        dy = torch.randn(M, N, generator=g).to(dev)
        dW, db, dx = torch.zeros_like(W), torch.zeros_like(b), torch.full_like(x, 7.0)
        ops.linear_bwd(x, W, dy, dW, db, dx=dx, silu_in=True)
        ops.silu_bwd_(dx, x)
        xr = x.double().clone().requires_grad_(True)
        Wr, br = W.double().clone().requires_grad_(True), b.double().clone().requires_grad_(True)
        (F.linear(F.silu(xr), Wr, br) * dy.double()).sum().backward()
        assert relerr(dW, Wr.grad) < 5e-6 and relerr(db, br.grad) < 5e-6 and relerr(dx, xr.grad) < 5e-6, (M, N, K)


@pytest.mark.parametrize("S,Cin,nv,B", [(64, 3, 4, 8), (32, 1, 2, 5), (96, 4, 4, 3), (28, 1, 2, 4)])
def test_encoder_forward_backward_vs_oracle(S, Cin, nv, B):
    """GaussianConvEncoder.encode (ref nn.py:93-110), training mode: mu / var, every parameter gradient, the BatchNorm
    running buffers; then eval mode on the advanced buffers"""
    from causaldiffae_b200.nn import GaussianConvEncoder, encoder_hidden_dims
    from oracle import model as om
    dims = encoder_hidden_dims(S, nv)
    cfg = om.UNetConfig(image_size=S, in_channels=Cin, model_channels=32, num_res_blocks=1, rep_dim=512, n_vars=nv,
                        encoder_dims=list(dims))
    full = om.seeded_state_dict(cfg, seed=3)
    sd = {k: v.cuda() for k, v in full.items() if k.startswith("rep_emb.")}
    enc = GaussianConvEncoder(in_channels=Cin, latent_dim=512, hidden_dims=dims, num_vars=nv)
    enc.load_state_dict({k[len("rep_emb."):]: v for k, v in sd.items()}, strict=True)
    enc.cuda().train()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, Cin, S, S, generator=g).cuda()
    gm, gv = torch.randn(B, 512, generator=g).cuda(), torch.randn(B, 512, generator=g).cuda()
    names = [k for k in sd if "running" not in k and "num_batches" not in k]
    for n in names:
        sd[n].requires_grad_(True)
    mu_r, var_r = om.encoder_encode(sd, cfg, x, training=True)
    ((mu_r * gm).sum() + (var_r * gv).sum()).backward()
    mu, var = enc.encode(x)
    ((mu * gm).sum() + (var * gv).sum()).backward()
    assert relerr(mu, mu_r) < 1e-4 and relerr(var, var_r) < 1e-4, (relerr(mu, mu_r), relerr(var, var_r))
    named = dict(enc.named_parameters())
    gtot = math.sqrt(sum(float((sd[n].grad.double() ** 2).sum()) for n in names))
    worst = 0.0
    for n in names:
        ref = sd[n].grad
        got = named[n[len("rep_emb."):]].grad
        got = torch.zeros_like(ref) if got is None else got
        if float(ref.norm()) < 1e-6 * gtot:          # conv biases in front of a batch-statistics BatchNorm: exactly zero here
            assert float(got.norm()) <= 1e-5 * gtot, n
            continue
        worst = max(worst, relerr(got, ref))
        assert relerr(got, ref) < 2e-4, (n, relerr(got, ref))
    print(f"encoder {S}px: mu {relerr(mu, mu_r):.2e} var {relerr(var, var_r):.2e} worst parameter gradient {worst:.2e}")
    for k in range(len(dims)):
        q = f"rep_emb.encoder.{k}.1."
        bn = enc.encoder[k][1]
        assert relerr(bn.running_mean, sd[q + "running_mean"]) < 1e-4 and relerr(bn.running_var, sd[q + "running_var"]) < 1e-4
        assert int(bn.num_batches_tracked) == int(sd[q + "num_batches_tracked"]) == 1
    enc.eval()
    with torch.no_grad():
        mu_e, var_e = enc.encode(x)
        mu_er, var_er = om.encoder_encode(sd, cfg, x, training=False)
    assert relerr(mu_e, mu_er) < 1e-4 and relerr(var_e, var_er) < 1e-4


@pytest.mark.parametrize("class_cond,context_cond", [(False, False), (True, True)])
def test_trunk_and_film_vs_oracle(class_cond, context_cond):
    """timestep_embedding -> time_embed (+ label_emb, c_emb) + up_emb(z) -> all emb_layers (ref unet.py:545-554,616,148-154)"""
    from causaldiffae_b200 import script_util as su
    from causaldiffae_b200.rep import _FilmFn, anchor
    from oracle import model as om
    flags = dict(image_size=32, num_channels=64, num_res_blocks=1, class_cond=class_cond, context_cond=context_cond,
                 rep_cond=True, n_vars=4, causal_modeling=True, in_channels=3, learn_sigma=False)
    full = {**su.model_and_diffusion_defaults(), **flags}
    model, _ = su.create_model_and_diffusion(**full)
    cfg = om.config_from_flags(**full)
    sd = om.seeded_state_dict(cfg, seed=2)
    model.load_state_dict(sd, strict=True)
    model.cuda()
    sd = {k: v.cuda() for k, v in sd.items()}
    eng = model.engine
    g = torch.Generator().manual_seed(8)
    B = 6
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    y = torch.randint(0, 10, (B,), generator=g).cuda() if class_cond else None
    c = torch.rand(B, 4, generator=g).cuda() if context_cond else None
    z = torch.randn(B, 512, generator=g).cuda()
    gf = torch.randn(B, eng.film_width, generator=g).cuda()
    names = [n for n in sd if n.split(".")[0] in ("time_embed", "label_emb", "c_emb", "up_emb") or ".emb_layers." in n]
    for n in names:
        sd[n].requires_grad_(True)
    zr = z.clone().requires_grad_(True)
    emb = om.embedding_trunk(sd, cfg, t, y, c) + F.linear(zr, sd["up_emb.weight"], sd["up_emb.bias"])
    rbs = eng.resblocks()
    prefixes = [n[:-len("emb_layers.1.weight")] for n in sd if n.endswith("emb_layers.1.weight")]
    film_ref = torch.cat([F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"]) for p in prefixes], dim=1)
    (film_ref * gf).sum().backward()
    eng.grad_arena.zero_()
    zz = z.clone().requires_grad_(True)
    film = _FilmFn.apply(model, anchor(z.device), t, y, c, zz, None, 0.0, True)
    (film * gf).sum().backward()
    # the trunk (time_embed, label_emb, c_emb, up_emb) is fp32; the FiLM projection itself (a [B,256] x [256, sum 2C] GEMM and
    # its two backward GEMMs) runs on the tensor cores with bf16 operands and fp32 accumulation, like the torso it feeds
    assert relerr(film, film_ref) < 5e-3, relerr(film, film_ref)
    assert relerr(zz.grad, zr.grad) < 1e-2
    named = dict(model.named_parameters())
    for n in names:
        assert relerr(named[n].grad, sd[n].grad) < 1e-2, (n, relerr(named[n].grad, sd[n].grad))
    assert len(rbs) == len(prefixes)
    # respaced / rescaled model timesteps (ref respace.py:119-124) folded into the embedding kernel
    tmap = torch.arange(0, 1000, 20, device="cuda")
    ts = torch.randint(0, 50, (B,), generator=g).cuda()
    with torch.no_grad():
        f1 = _FilmFn.apply(model, anchor(z.device), ts, y, c, z, tmap, 0.25, False)
        f2 = _FilmFn.apply(model, anchor(z.device), tmap[ts].float() * 0.25, y, c, z, None, 0.0, False)
    assert torch.equal(f1, f2)


@pytest.mark.parametrize("masked,causal", [(False, True), (True, True), (True, False)])
def test_latent_and_step_loss_vs_torch(masked, causal):
    from causaldiffae_b200 import ops
    from oracle import diffusion as od
    g = torch.Generator().manual_seed(4)
    B, D, n = 9, 512, 4
    dev = torch.device("cuda")
    mu, zp, xi = (torch.randn(B, D, generator=g).to(dev) for _ in range(3))
    var = (torch.rand(B, D, generator=g) + 0.05).to(dev)
    c, w = torch.rand(B, n, generator=g).to(dev), (torch.rand(B, generator=g) + 0.5).to(dev)
    keep = (torch.rand(B, generator=g) < 0.6).float().to(dev) if masked else None
    mse = torch.rand(B, generator=g).to(dev)
    gz = torch.randn(B, D, generator=g).to(dev)
    klw = 0.37
    # torch reference: the oracle's representation_loss on the same quantities
    mur, varr, zpr = (t.clone().requires_grad_(True) for t in (mu, var, zp))
    kp = keep[:, None] if masked else 1.0
    z_ref = (zpr + (varr * 0.001) ** 0.5 * xi) * kp
    zpm_ref = zpr * kp
    kld_ref = od.representation_loss(mur, varr, zpm_ref, causal, keep, c)
    loss_ref = mse + klw * kld_ref
    total_ref = (loss_ref * w).mean()
    (total_ref + (z_ref * gz).sum()).backward()
    # kernels
    z, zpm, kld = torch.empty_like(mu), torch.empty_like(mu), torch.empty(B, device=dev)
    ops.latent_fwd(mu, var, zp, xi, keep, c, z, zpm, kld, n, causal, 0.001)
    loss, gscale, dkld, total = [torch.empty(B, device=dev) for _ in range(3)] + [torch.zeros(1, device=dev)]
    logs = torch.zeros(20, device=dev)
    t = torch.randint(0, 1000, (B,), generator=g).to(dev)
    ops.step_loss(mse, kld, keep, w, torch.tensor([klw], device=dev), t, 1000, loss, gscale, dkld, total, logs)
    assert relerr(z, z_ref) < 1e-6 and relerr(zpm, zpm_ref) < 1e-6
    assert relerr(loss, loss_ref.expand(B) if loss_ref.dim() == 0 else loss_ref) < 1e-5
    assert relerr(total, total_ref.reshape(1)) < 1e-5
    assert relerr(gscale, w / B) < 1e-6
    dzp, dmu, dvar = torch.empty_like(mu), torch.empty_like(mu), torch.empty_like(mu)
    ops.latent_bwd(mu, var, zp, xi, keep, c, gz, dkld, None, None, None, dzp, dmu, dvar, n, causal, 0.001)
    assert relerr(dzp, zpr.grad) < 1e-4 and relerr(dmu, mur.grad) < 1e-4 and relerr(dvar, varr.grad) < 1e-4
    # logger sums: weighted means and per-quartile sums (ref train_util.py:401-407)
    q = (4 * t // 1000).clamp(0, 3)
    lw = (loss_ref * w).detach()
    np.testing.assert_allclose(float(logs[0]), float(lw.sum() if lw.dim() else lw * B), rtol=1e-4)
    for k in range(4):
        np.testing.assert_allclose(float(logs[8 + k]), float((mse * w)[q == k].sum()), rtol=1e-4, atol=1e-6)
        assert int(logs[16 + k]) == int((q == k).sum())


def test_counter_based_generator():
    from causaldiffae_b200 import ops
    dev = torch.device("cuda")
    state = torch.tensor([1234, 0], device=dev, dtype=torch.int64)
    a = ops.randn_(torch.empty(1 << 20, device=dev), state)
    assert int(state[1]) == (1 << 18)
    b = ops.randn_(torch.empty(1 << 20, device=dev), state)        # continues the stream
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3 and abs(float((a * b).mean())) < 5e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05 and bool(torch.isfinite(a).all())
    state2 = torch.tensor([1234, 0], device=dev, dtype=torch.int64)
    a2 = ops.randn_(torch.empty(1 << 20, device=dev), state2)
    assert torch.equal(a, a2)                                      # same seed / offset -> same draws
    k = ops.randn_(torch.empty(100003, device=dev), state, bernoulli=True, keep_prob=0.5)
    assert set(k.unique().tolist()) <= {0.0, 1.0} and abs(float(k.mean()) - 0.5) < 0.01
