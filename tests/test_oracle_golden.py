"""oracle/ (CPU restatement) against fixtures produced by the real reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import schedules, model as om, diffusion as od
from tests.golden import cases


def test_schedule_tables_bit_exact(golden):
    for name, steps in cases.SCHEDULES:
        tabs = schedules.diffusion_tables(schedules.named_beta_schedule(name, steps))
        for tab in cases.TABLES:
            np.testing.assert_array_equal(tabs[tab], golden[f"sched/{name}{steps}/{tab}"], err_msg=f"{name}{steps}/{tab}")


def test_space_timesteps_and_spaced_tables_bit_exact(golden):
    for steps, spec in cases.RESPACINGS:
        use = schedules.space_timesteps(steps, spec)
        np.testing.assert_array_equal(np.array(sorted(use)), golden[f"space/{steps}/{spec}"])
        d = od.Diffusion(steps=steps, timestep_respacing=spec)
        np.testing.assert_array_equal(np.array(d.timestep_map), golden[f"spaced/{steps}/{spec}/timestep_map"])
        np.testing.assert_array_equal(d.tables["betas"], golden[f"spaced/{steps}/{spec}/betas"])
        np.testing.assert_array_equal(d.tables["alphas_cumprod_prev"], golden[f"spaced/{steps}/{spec}/alphas_cumprod_prev"])


def test_space_timesteps_known_answers():
    # SURVEY.md 8a KATs
    assert sorted(schedules.space_timesteps(1000, "ddim10")) == list(range(0, 1000, 100))
    assert sorted(schedules.space_timesteps(1000, "10")) == [0, 111, 222, 333, 444, 555, 666, 777, 888, 999]
    assert sorted(schedules.space_timesteps(1000, "50"))[:4] == [0, 20, 41, 61]
    with pytest.raises(ValueError):
        schedules.space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError):
        schedules.space_timesteps(10, "20")


def test_uniform_sampler_bit_exact(golden):
    for seed, T, B in cases.SAMPLER:
        np.random.seed(seed)
        t, w = schedules.uniform_sample_t(T, B)
        np.testing.assert_array_equal(t, golden[f"sampler/{seed}/{T}/{B}/t"])
        np.testing.assert_array_equal(w, golden[f"sampler/{seed}/{T}/{B}/w"])
    np.random.seed(123)
    assert schedules.uniform_sample_t(1000, 8)[0].tolist() == [696, 286, 226, 551, 719, 423, 980, 684]


def test_timestep_embedding_and_kl_weight(golden):
    for ts, dim in cases.TEMB:
        np.testing.assert_array_equal(om.timestep_embedding(torch.tensor(ts), dim).numpy(), golden[f"temb/{dim}"])
    np.testing.assert_array_equal(np.array([schedules.kl_weight_schedule(s) for s in cases.KLW_STEPS]), golden["klw"])


def test_topo_order_identity_on_shipped_dags():
    for A in schedules.DAGS.values():
        assert schedules.topo_order(A) == list(range(len(A)))
    assert schedules.topo_order([[0, 0], [1, 0]]) == [1, 0]
    with pytest.raises(ValueError):
        schedules.topo_order([[0, 1], [1, 0]])


def _setup(case):
    cfg = om.config_from_flags(**case["flags"], A=case.get("A"))
    sd = om.seeded_state_dict(cfg, seed=case["wseed"])
    return cfg, sd, cases.make_inputs(case)


@pytest.mark.parametrize("name", list(cases.ALL_CASES))
def test_training_losses_and_grads(golden, golden_shipped, name):
    case = cases.ALL_CASES[name]
    golden = golden_shipped if name in cases.SHIPPED_CASES else golden
    cfg, sd, inp = _setup(case)
    for n in om.trainable_names(cfg):
        sd[n].requires_grad_(True)
    diff = od.Diffusion(steps=case["flags"]["diffusion_steps"])
    diff.kl_weight = case["kl_weight"]
    torch.manual_seed(case["rseed"])
    terms = od.training_losses(diff, sd, cfg, inp["x0"], inp["t"], inp["noise"],
                               y=inp["y"] if cfg.num_classes else None, c=inp["c"])
    (terms["loss"] * inp["w"]).mean().backward()
    for k in ("mse", "kld_rep", "loss"):
        np.testing.assert_allclose(terms[k].detach().numpy(), golden[f"{name}/{k}"], rtol=2e-5, atol=1e-6, err_msg=k)
    gsq = sum(float((sd[n].grad ** 2).sum()) for n in om.trainable_names(cfg))
    np.testing.assert_allclose(gsq, float(golden[f"{name}/grad_sqsum"]), rtol=1e-4)
    for pn in case["grad_probe"]:
        ref = golden[f"{name}/grad/{pn}"]
        got = sd[pn].grad.numpy()
        assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref) + 1e-9, pn


@pytest.mark.parametrize("name", list(cases.ALL_CASES))
def test_eval_paths(golden, golden_shipped, name):
    case = cases.ALL_CASES[name]
    golden = golden_shipped if name in cases.SHIPPED_CASES else golden
    cfg, sd, inp = _setup(case)
    diff = od.Diffusion(steps=case["flags"]["diffusion_steps"])
    with torch.no_grad():
        # the fixture ran one training-mode forward first: BatchNorm running stats were updated by it (Q15)
        torch.manual_seed(case["rseed"])
        od.training_losses(diff, sd, cfg, inp["x0"], inp["t"], inp["noise"], y=inp["y"] if cfg.num_classes else None,
                           c=inp["c"])
        x_t = diff.q_sample(inp["x0"], inp["t"], inp["noise"])
        sub = case.get("sub", 1)                  # large fixtures keep every sub-th pixel
        np.testing.assert_allclose(x_t.numpy()[..., ::sub, ::sub], golden[f"{name}/x_t"], rtol=0, atol=1e-7)
        eps = om.unet_forward(sd, cfg, x_t, diff.model_timesteps(inp["t"]), y=inp["y"] if cfg.num_classes else None,
                              z=inp["z"], training=False)[0]
        ref = golden[f"{name}/eps_given_z"]
        assert np.linalg.norm(eps.numpy()[..., ::sub, ::sub] - ref) <= 2e-5 * np.linalg.norm(ref)
        mu, var = om.encoder_encode(sd, cfg, inp["x0"], training=False)
        np.testing.assert_allclose(mu.numpy(), golden[f"{name}/enc_mu_eval"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(var.numpy(), golden[f"{name}/enc_var_eval"], rtol=1e-4, atol=1e-6)
        zp = om.nonlinearity_add_back_noise(sd, mu, om.causal_masking(mu, cfg.A, cfg.n_vars), cfg.n_vars)
        np.testing.assert_allclose(zp.numpy(), golden[f"{name}/z_post_eval"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name,tag,w", [(n, t, w) for n, c in cases.ALL_CASES.items() for t, w in c["ddim"]])
def test_ddim_counterfactual(golden, golden_shipped, name, tag, w):
    case = cases.ALL_CASES[name]
    golden = golden_shipped if name in cases.SHIPPED_CASES else golden
    cfg, sd, inp = _setup(case)
    d0 = od.Diffusion(steps=case["flags"]["diffusion_steps"])
    with torch.no_grad():
        torch.manual_seed(case["rseed"])
        od.training_losses(d0, sd, cfg, inp["x0"], inp["t"], inp["noise"], y=inp["y"] if cfg.num_classes else None,
                           c=inp["c"])
    diff = od.Diffusion(steps=case["flags"]["diffusion_steps"], timestep_respacing=case["respacing"])
    torch.manual_seed(case["rseed"] + 1)
    xi = torch.randn(inp["x0"].shape[0], 512)
    sample, _, _ = od.counterfactual(diff, sd, cfg, inp["x0"], inp["noise"], xi, do_var=0, do_value=case["do_value"],
                                     on="mu", w=w, y=inp["y"] if cfg.num_classes else None)
    ref = golden[f"{name}/ddim/{tag}"]
    mse = float(((sample.numpy() - ref) ** 2).mean())
    psnr = 10 * np.log10(1.0 / max(mse, 1e-20))
    assert psnr > 80, psnr


def test_three_train_steps_match_reference_trainloop(golden):
    name = "mnist32"
    case = cases.MODEL_CASES[name]
    cfg, sd, inp = _setup(case)
    diff = od.Diffusion(steps=case["flags"]["diffusion_steps"])
    tr = od.RefTrainer(sd, cfg, diff, lr=1e-3, ema_rate=0.99)
    losses = []
    for s in range(case["train_steps"]):
        np.random.seed(case["rseed"] + 10 + s)
        torch.manual_seed(case["rseed"] + 20 + s)
        t, w = schedules.uniform_sample_t(diff.num_timesteps, inp["x0"].shape[0])
        noise = torch.randn_like(inp["x0"])
        out = tr.run_step(inp["x0"], torch.from_numpy(t), noise, torch.from_numpy(w), y=inp["y"], c=inp["c"])
        losses.append(out["loss"])
    np.testing.assert_allclose(losses, golden[f"{name}/train/loss"], rtol=2e-4)
    for pn in case["grad_probe"]:
        ref = golden[f"{name}/train/param/{pn}"]
        assert np.linalg.norm(sd[pn].detach().numpy() - ref) <= 2e-4 * np.linalg.norm(ref) + 1e-7, pn
        refe = golden[f"{name}/train/ema/{pn}"]
        assert np.linalg.norm(tr.ema[pn].numpy() - refe) <= 2e-4 * np.linalg.norm(refe) + 1e-7, pn
