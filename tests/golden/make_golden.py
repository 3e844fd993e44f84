"""Generate tests/golden/*.npz by running the REAL reference (read-only /root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
The reference ships no golden vectors, so these fixtures — outputs of the
reference itself on seeded inputs — are what pins oracle/ (tests/test_oracle_golden.py).
Inputs and weights are NOT stored: they are regenerated from the seeds in `cases.py`.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim, model as om  # noqa: E402
from tests.golden import cases  # noqa: E402


def gen_schedules(ns, out):
    for name, steps in cases.SCHEDULES:
        betas = ns.gd.get_named_beta_schedule(name, steps)
        d = ns.gd.GaussianDiffusion(betas=betas, model_mean_type=ns.gd.ModelMeanType.EPSILON,
                                    model_var_type=ns.gd.ModelVarType.FIXED_LARGE, loss_type=ns.gd.LossType.MSE)
        for tab in cases.TABLES:
            out[f"sched/{name}{steps}/{tab}"] = getattr(d, tab)
    for steps, spec in cases.RESPACINGS:
        use = ns.respace.space_timesteps(steps, spec)
        out[f"space/{steps}/{spec}"] = np.array(sorted(use), dtype=np.int64)
        d = ns.su.create_gaussian_diffusion(steps=steps, timestep_respacing=spec)
        out[f"spaced/{steps}/{spec}/timestep_map"] = np.array(d.timestep_map, dtype=np.int64)
        out[f"spaced/{steps}/{spec}/betas"] = d.betas
        out[f"spaced/{steps}/{spec}/alphas_cumprod_prev"] = d.alphas_cumprod_prev
    for seed, T, B in cases.SAMPLER:
        np.random.seed(seed)
        d = ns.su.create_gaussian_diffusion(steps=T)
        s = ns.resample.create_named_schedule_sampler("uniform", d)
        t, w = s.sample(B, torch.device("cpu"))
        out[f"sampler/{seed}/{T}/{B}/t"] = t.numpy()
        out[f"sampler/{seed}/{T}/{B}/w"] = w.numpy()
    for ts, dim in cases.TEMB:
        out[f"temb/{dim}"] = ns.nn.timestep_embedding(torch.tensor(ts), dim).numpy()
    tl = object.__new__(ns.train.TrainLoop)
    out["klw"] = np.array([tl.linear_kl_weight_scheduler(s, 50000, 0.0, 1.0) for s in cases.KLW_STEPS])


def gen_model_case(ns, name, case, out):
    flags, A = case["flags"], case.get("A")
    model, diff = refshim.build(flags, rep_dim=512, A=A)
    cfg = om.config_from_flags(**flags, A=A)
    model.load_state_dict(om.seeded_state_dict(cfg, seed=case["wseed"]), strict=True)
    model.train()
    inp = cases.make_inputs(case)
    kw = {}
    if flags.get("class_cond"):
        kw["y"] = inp["y"]
    kw["c"] = inp["c"]
    diff.kl_weight = case.get("kl_weight", 0.0)
    torch.manual_seed(case["rseed"])
    terms = diff.training_losses(model, inp["x0"], inp["t"], model_kwargs=dict(kw), noise=inp["noise"],
                                 rep_cond=True, causal_modeling=flags["causal_modeling"])
    loss = (terms["loss"] * inp["w"]).mean()
    loss.backward()
    for k in ("mse", "kld_rep", "loss"):
        out[f"{name}/{k}"] = terms[k].detach().numpy()
    named = dict(model.named_parameters())
    out[f"{name}/grad_sqsum"] = np.array(sum(float((p.grad ** 2).sum()) for p in named.values()))
    for pn in case["grad_probe"]:
        out[f"{name}/grad/{pn}"] = named[pn].grad.numpy().copy()
    # plain forward with injected z (sampling path) and the 5-tuple of the training path
    model.eval()
    with torch.no_grad():
        x_t = diff.q_sample(inp["x0"], inp["t"], inp["noise"])
        sub = case.get("sub", 1)
        out[f"{name}/x_t"] = x_t.numpy()[..., ::sub, ::sub]
        zkw = {k: v for k, v in kw.items() if k == "y"}
        eps_z = model(x_t, torch.tensor(diff.timestep_map)[inp["t"]], z=inp["z"], **zkw)[0]
        out[f"{name}/eps_given_z"] = eps_z.numpy()[..., ::sub, ::sub]
        mu, var = model.rep_emb.encode(inp["x0"])
        out[f"{name}/enc_mu_eval"], out[f"{name}/enc_var_eval"] = mu.numpy(), var.numpy()
        At = torch.tensor(cfg.A, dtype=torch.float32)
        z_pre = model.causal_mask.causal_masking(mu, At)
        out[f"{name}/z_post_eval"] = model.causal_mask.nonlinearity_add_back_noise(mu, z_pre).numpy()
    # DDIM counterfactual (recipe ref scripts/image_causaldae_test.py:405-436), optional guidance
    for tag, w in case["ddim"]:
        m2, d2 = refshim.build({**flags, "timestep_respacing": case["respacing"]}, rep_dim=512, A=A)
        m2.load_state_dict(model.state_dict(), strict=True)
        m2.eval()
        with torch.no_grad():
            mu, var = m2.rep_emb.encode(inp["x0"])
            var = torch.ones(var.shape) * 0.001
            d = 512 // flags["n_vars"]
            mu[:, :d] = case["do_value"]
            z_pre = m2.causal_mask.causal_masking(mu, At)
            z_post = m2.causal_mask.nonlinearity_add_back_noise(mu, z_pre)
            torch.manual_seed(case["rseed"] + 1)
            z = ns.nn.reparameterize(z_post, var)
            t = torch.tensor([d2.num_timesteps - 1] * inp["x0"].shape[0])
            x_T = d2.q_sample(inp["x0"], t, noise=inp["noise"])
            ckw = dict(z=z)
            if flags.get("class_cond"):
                ckw["y"] = inp["y"]
            sample = d2.ddim_sample_loop(m2, tuple(inp["x0"].shape), noise=x_T, clip_denoised=True,
                                         model_kwargs=ckw, w=w)
        out[f"{name}/ddim/{tag}"] = sample.numpy()
    # three optimisation steps through the reference TrainLoop (AdamW lr 1e-3 so that weights move visibly)
    if case.get("train_steps"):
        m3, d3 = refshim.build(flags, rep_dim=512, A=A)
        m3.load_state_dict(om.seeded_state_dict(cfg, seed=case["wseed"]), strict=True)
        m3.train()
        ns.dist.setup_dist()
        ns.logger.configure(dir="/tmp/cdae_golden_log", format_strs=[])
        tl = ns.train.TrainLoop(model=m3, diffusion=d3, data=None, batch_size=inp["x0"].shape[0], microbatch=-1,
                                lr=1e-3, ema_rate="0.99", log_interval=10 ** 9, save_interval=10 ** 9,
                                resume_checkpoint="", rep_cond=True, n_vars=flags["n_vars"],
                                causal_modeling=flags["causal_modeling"], in_channels=flags["in_channels"],
                                masking=flags.get("masking", False))
        losses = []
        for s in range(case["train_steps"]):
            np.random.seed(case["rseed"] + 10 + s)
            torch.manual_seed(case["rseed"] + 20 + s)
            cond = {k: v for k, v in kw.items()}
            tl.run_step(inp["x0"], cond)
            losses.append(ns.logger.getkvs().get("loss", np.nan))
            ns.logger.dumpkvs()
            tl.step += 1
            d3.kl_weight = tl.linear_kl_weight_scheduler(tl.step, 50000, 0.0, 1.0)
        out[f"{name}/train/loss"] = np.array(losses, dtype=np.float64)
        named3 = dict(m3.named_parameters())
        for pn in case["grad_probe"]:
            out[f"{name}/train/param/{pn}"] = named3[pn].detach().numpy().copy()
        names = [n for n, _ in m3.named_parameters()]
        for pn in case["grad_probe"]:
            out[f"{name}/train/ema/{pn}"] = tl.ema_params[0][names.index(pn)].detach().numpy().copy()


def main():
    assert refshim.available(), "reference tree not present"
    ns = refshim.load()
    out = {}
    if "--shipped" in sys.argv:       # the shipped image sizes (28 / 96 / 128 px) -> golden_v2_shipped.npz
        for name, case in cases.SHIPPED_CASES.items():
            gen_model_case(ns, name, case, out)
            print("generated", name)
        path = os.path.join(HERE, "golden_v2_shipped.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) / 1e6, "MB", len(out), "arrays")
        return
    gen_schedules(ns, out)
    for name, case in cases.MODEL_CASES.items():
        gen_model_case(ns, name, case, out)
        print("generated", name)
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB", len(out), "arrays")


if __name__ == "__main__":
    main()
