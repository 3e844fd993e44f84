"""Host-side logic of causaldiffae_b200 that needs no GPU: integer/float64 work bit-exact against the reference
fixtures, the C-ABI surface, API compatibility, and the loud failure without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_space_timesteps_and_spaced_diffusion_bit_exact(golden):
    from causaldiffae_b200.respace import space_timesteps
    from causaldiffae_b200.script_util import create_gaussian_diffusion
    for steps, spec in cases.RESPACINGS:
        assert sorted(space_timesteps(steps, spec)) == golden[f"space/{steps}/{spec}"].tolist()
        d = create_gaussian_diffusion(steps=steps, timestep_respacing=spec)
        assert d.timestep_map == golden[f"spaced/{steps}/{spec}/timestep_map"].tolist()
        np.testing.assert_array_equal(d.betas, golden[f"spaced/{steps}/{spec}/betas"])
        np.testing.assert_array_equal(d.alphas_cumprod_prev, golden[f"spaced/{steps}/{spec}/alphas_cumprod_prev"])
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError):
        space_timesteps(10, "20")


def test_schedule_tables_bit_exact(golden):
    from causaldiffae_b200 import gaussian_diffusion as gd
    for name, steps in cases.SCHEDULES:
        d = gd.GaussianDiffusion(betas=gd.get_named_beta_schedule(name, steps), model_mean_type=gd.ModelMeanType.EPSILON,
                                 model_var_type=gd.ModelVarType.FIXED_LARGE, loss_type=gd.LossType.MSE)
        for tab in cases.TABLES:
            np.testing.assert_array_equal(getattr(d, tab), golden[f"sched/{name}{steps}/{tab}"])
    with pytest.raises(NotImplementedError):
        gd.get_named_beta_schedule("quadratic", 10)


def test_uniform_sampler_and_embedding_bit_exact(golden):
    from causaldiffae_b200.resample import create_named_schedule_sampler
    from causaldiffae_b200.script_util import create_gaussian_diffusion
    from causaldiffae_b200.nn import timestep_embedding
    for seed, T, B in cases.SAMPLER:
        np.random.seed(seed)
        s = create_named_schedule_sampler("uniform", create_gaussian_diffusion(steps=T))
        t, w = s.sample(B, torch.device("cpu"))
        np.testing.assert_array_equal(t.numpy(), golden[f"sampler/{seed}/{T}/{B}/t"])
        np.testing.assert_array_equal(w.numpy(), golden[f"sampler/{seed}/{T}/{B}/w"])
    with pytest.raises(NotImplementedError):
        create_named_schedule_sampler("nope", None)
    for ts, dim in cases.TEMB:
        np.testing.assert_array_equal(timestep_embedding(torch.tensor(ts), dim).numpy(), golden[f"temb/{dim}"])
    # float timesteps (rescale_timesteps=True path) are accepted
    assert timestep_embedding(torch.tensor([0.5, 2.0]), 8).shape == (2, 8)


def test_kl_weight_schedule_and_topo_order(golden):
    from causaldiffae_b200.train_util import TrainLoop
    from causaldiffae_b200.nn import topo_order
    from causaldiffae_b200.unet import DAGS
    tl = object.__new__(TrainLoop)
    got = [tl.linear_kl_weight_scheduler(s, 50000, 0.0, 1.0) for s in cases.KLW_STEPS]
    np.testing.assert_array_equal(np.array(got), golden["klw"])
    for A in DAGS.values():
        assert topo_order(A) == list(range(len(A)))


def test_ddim_coef_table_matches_oracle_arithmetic():
    from causaldiffae_b200.gaussian_diffusion import ddim_coef_table
    from oracle import diffusion as od
    d = od.Diffusion(steps=1000, timestep_respacing="ddim50")
    for eta in (0.0, 0.7):
        tab = ddim_coef_table(d.tables, eta=eta)
        t = torch.arange(50)
        ab, abp = d.extract("alphas_cumprod", t, 1), d.extract("alphas_cumprod_prev", t, 1)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        # numpy's sqrt is correctly rounded; torch's vectorised CPU sqrt may differ by one ulp
        np.testing.assert_allclose(tab[:, 2], torch.sqrt(abp).numpy(), rtol=2e-7, atol=0)
        np.testing.assert_allclose(tab[:, 3], torch.sqrt(1 - abp - sigma ** 2).numpy(), rtol=2e-7, atol=0)
        np.testing.assert_allclose(tab[1:, 4], sigma[1:].numpy(), rtol=4e-7, atol=0)
        assert tab[0, 4] == 0.0


def test_state_dict_keys_and_shapes_match_reference_format():
    """checkpoint wire format: identical keys/shapes/order as the reference-pinned oracle layout"""
    from causaldiffae_b200 import script_util as su
    from oracle import model as om
    for flags in (dict(image_size=64, rep_cond=True, causal_modeling=True, n_vars=4),
                  dict(image_size=32, num_channels=64, class_cond=True, rep_cond=True, causal_modeling=True, n_vars=2, in_channels=1),
                  dict(image_size=28, num_channels=32, num_res_blocks=1, context_cond=True)):
        full = {**su.model_and_diffusion_defaults(), **flags}
        model, _ = su.create_model_and_diffusion(**full)
        ref = om.param_shapes(om.config_from_flags(**full))
        sd = model.state_dict()
        assert list(sd.keys()) == [n for n, _, _ in ref]
        for n, shape, _ in ref:
            assert tuple(sd[n].shape) == tuple(shape), n
    with pytest.raises(ValueError):
        su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), "image_size": 48})
    assert len(su.model_and_diffusion_defaults()) == 26


def test_c_abi_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    hdr = open(os.path.join(ROOT, "include", "cdae.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(cdae_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "causaldiffae_b200", "libcdae.so"))
    for name in declared:
        assert hasattr(lib, name), f"libcdae.so does not export {name}"
    from causaldiffae_b200 import _lib
    assert set(_lib._SIGS) <= declared | {"cdae_version", "cdae_init"}
    assert _lib.load().cdae_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """the product path must fail loudly without the CUDA device (no silent eager/CPU fallback)"""
    from causaldiffae_b200 import script_util as su, ops
    from causaldiffae_b200._lib import CdaeError
    model, diff = su.create_model_and_diffusion(**{**su.model_and_diffusion_defaults(), "image_size": 32, "num_channels": 64})
    x = torch.rand(2, 3, 32, 32)
    with pytest.raises((CdaeError, AssertionError)):
        diff.q_sample(x, torch.tensor([1, 2]), torch.randn_like(x))
    with pytest.raises(CdaeError):
        model(x, torch.tensor([1, 2]))
    src = "".join(open(os.path.join(ROOT, "causaldiffae_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "causaldiffae_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src, "the product must never import the oracle"


def test_fp16_util_master_param_round_trip():
    """ref fp16_util.py:27-76: one flat fp32 master parameter; gradients flatten in parameter order; zero_grad in place"""
    from causaldiffae_b200 import fp16_util as fu
    g = torch.Generator().manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((3, 4), (5,), (2, 2, 2))]
    master = fu.make_master_params(params)
    assert len(master) == 1 and master[0].requires_grad and master[0].dtype == torch.float32
    assert torch.equal(master[0].detach(), torch.cat([p.detach().reshape(-1) for p in params]))
    for p in params:
        p.grad = torch.randn(p.shape, generator=g)
    fu.model_grads_to_master_grads(params, master)
    assert torch.equal(master[0].grad, torch.cat([p.grad.reshape(-1) for p in params]))
    with torch.no_grad():
        master[0].mul_(2.0)
    fu.master_params_to_model_params(params, master)
    for p, u in zip(params, fu.unflatten_master_params(params, master)):
        assert torch.equal(p.detach(), u) and u.shape == p.shape
    grads = [p.grad for p in params]
    fu.zero_grad(params)
    assert all(p.grad is gr and float(gr.abs().sum()) == 0.0 for p, gr in zip(params, grads))    # zeroed in place, never None
    lin = torch.nn.Linear(2, 2)
    assert fu.convert_module_to_f16(lin) is lin and fu.convert_module_to_f32(lin) is lin and lin.weight.dtype == torch.float32


def test_c_abi_error_convention_without_a_gpu():
    """entry points return 0 / a negative cdae_status (never raise, never exit) and leave a message in
    cdae_last_error(); argument and shape checks run before any CUDA call, zero-sized batches are accepted as no-ops."""
    from causaldiffae_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(4096 + 64)
    p = (ctypes.addressof(buf) + 63) // 64 * 64               # a 64 B aligned, never dereferenced pointer
    OK, ERR_ARG, ERR_SHAPE = 0, -1, -2
    # null pointers
    assert lib.cdae_q_sample(None, p, p, p, p, p, 2, 64, None) == ERR_ARG
    assert b"null" in lib.cdae_last_error().lower()
    assert lib.cdae_gather_images(None, None, p, p, None, 2, 8, 8, 3, 0, 0, None) == ERR_ARG
    # shapes the kernels do not support
    assert lib.cdae_q_sample(p, p, p, p, p, p, 2, 63, None) == ERR_SHAPE                      # per_sample % 4
    assert lib.cdae_gather_images(p, None, p, p, None, 2, 8, 8, 5, 0, 0, None) == ERR_SHAPE   # 5 channels
    assert b"channels" in lib.cdae_last_error()
    assert lib.cdae_gather_images(p, None, p, p, None, 2, 3, 3, 1, 0, 0, None) == ERR_SHAPE   # H*W % 4
    assert lib.cdae_gn_apply_fwd(p, 40, p, None, 0, None, 2, 16, p, p, None, 0, 0, 1, p, p, p, None, None) == ERR_SHAPE   # C % 32
    assert lib.cdae_gn_bwd_apply(p, p, 64, None, 0, 2, 16, p, p, None, 0, 0, p, p, None, None, p, None, 0, None, None, None,
                                 None) == ERR_ARG                                          # statistics workspace missing
    d = _lib.IgemmDesc()
    d.out, d.wgt, d.nsrc, d.nseg, d.N, d.H, d.W = p, p, 1, 0, 1, 8, 8
    assert lib.cdae_igemm(ctypes.byref(d), None) == ERR_ARG                                   # no K segments
    assert b"nseg" in lib.cdae_last_error()
    # zero-sized batches: accepted before any pointer check (empty torch tensors carry null pointers)
    assert lib.cdae_q_sample(None, None, None, None, None, None, 0, 64, None) == OK
    assert lib.cdae_gather_images(None, None, None, None, None, 0, 8, 8, 3, 0, 0, None) == OK
    assert lib.cdae_gn_fwd(None, 64, None, 0, 0, 16, None, None, None, 0, 0, 1, None, None, None, None) == OK
    d.N = 0
    assert lib.cdae_igemm(ctypes.byref(d), None) == OK
