"""Seeded case definitions shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import torch

SCHEDULES = [("linear", 1000), ("linear", 2000), ("cosine", 100)]
TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
          "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
          "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
RESPACINGS = [(1000, "10"), (1000, "ddim10"), (1000, "ddim50"), (1000, "50"), (1000, "250"), (1000, "ddim250"),
              (1000, "ddim100"), (2000, "250"), (300, "10,15,20"), (100, "ddim5")]
SAMPLER = [(123, 1000, 8), (0, 1000, 64), (7, 250, 16), (5, 2000, 3)]
TEMB = [([0, 1, 999], 8), ([0, 3, 250, 999], 128), ([2, 11], 7)]
KLW_STEPS = [0, 1, 2, 100, 25000, 49999, 50000, 60000]

PENDULUM = [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]

_COMMON = dict(learn_sigma=False, rescale_timesteps=False, rescale_learned_sigmas=False, rep_cond=True,
               causal_modeling=True, diffusion_steps=100)

MODEL_CASES = {
    # MorphoMNIST-shaped: 1x32x32, 2-variable graph, class-conditional (BASELINE config 1, narrowed to 32 ch)
    "mnist32": dict(
        flags=dict(image_size=32, num_channels=32, num_res_blocks=1, class_cond=True, n_vars=2, in_channels=1, **_COMMON),
        A=None, B=3, wseed=0, iseed=11, rseed=77, kl_weight=0.25, respacing="ddim5", do_value=0.2,
        ddim=[("plain", None)], train_steps=3,
        grad_probe=["input_blocks.1.0.in_layers.2.weight", "out.2.weight", "causal_mask.nonlinearities.1.net.0.weight",
                    "rep_emb.fc_var.bias", "middle_block.1.qkv.weight", "label_emb.weight"]),
    # Pendulum-shaped: 3x64x64, 4-variable DAG injected (patch 2), classifier-free masking + guided DDIM (config 2/5)
    "pend64": dict(
        flags=dict(image_size=64, num_channels=32, num_res_blocks=1, class_cond=False, n_vars=4, in_channels=3,
                   masking=True, **_COMMON),
        A=PENDULUM, B=4, wseed=2, iseed=13, rseed=99, kl_weight=0.5, respacing="ddim5", do_value=-0.35,
        ddim=[("plain", None), ("w2", 2.0)], train_steps=0,
        grad_probe=["output_blocks.3.0.skip_connection.weight", "output_blocks.3.2.conv.weight",
                    "input_blocks.2.0.op.weight", "up_emb.weight", "rep_emb.encoder.0.0.weight",
                    "output_blocks.0.0.emb_layers.1.weight"]),
    # Circuit-shaped default DAG, additive (non scale-shift) ResBlock conditioning
    "circ32": dict(
        flags=dict(image_size=32, num_channels=32, num_res_blocks=1, class_cond=False, n_vars=4, in_channels=3,
                   use_scale_shift_norm=False, **_COMMON),
        A=None, B=2, wseed=4, iseed=17, rseed=55, kl_weight=0.0, respacing="ddim5", do_value=0.1,
        ddim=[], train_steps=0,
        grad_probe=["input_blocks.1.0.emb_layers.1.weight", "time_embed.0.weight"]),
}

CIRCUIT = [[0, 1, 1, 1], [0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 0, 0]]

# The image sizes of the reference's shipped launch lines (scripts/{morhomnist,pendulum,circuit}/train_*_causaldae.sh):
# 28 px (three levels, odd 7x7 bottom, attention at 28x28), 96 px (RGBA, no attention level, the reference's own 6-conv
# encoder), 128 px (six levels, attention at 16x16 / 8x8).  Narrow (32 channels, one res block) and small batches so
# that the fixtures stay small (`sub`: image-shaped outputs are stored every sub-th pixel); stored in
# golden_v2_shipped.npz by make_golden.py --shipped.
SHIPPED_CASES = {
    "mnist28": dict(
        flags=dict(image_size=28, num_channels=32, num_res_blocks=1, class_cond=True, n_vars=2, in_channels=1,
                   masking=True, **_COMMON),
        A=None, B=2, wseed=6, iseed=21, rseed=31, kl_weight=0.3, respacing="ddim5", do_value=0.2,
        ddim=[("w15", 1.5)], train_steps=0,
        grad_probe=["out.2.weight", "input_blocks.1.1.qkv.weight", "rep_emb.fc_mu.bias", "label_emb.weight"]),
    "pend96": dict(
        flags=dict(image_size=96, num_channels=32, num_res_blocks=1, class_cond=False, n_vars=4, in_channels=4,
                   masking=True, **_COMMON),
        A=PENDULUM, B=2, sub=4, wseed=7, iseed=22, rseed=32, kl_weight=0.3, respacing="ddim5", do_value=-0.2,
        ddim=[], train_steps=0,
        grad_probe=["out.2.weight", "input_blocks.0.0.weight", "rep_emb.encoder.5.0.weight", "output_blocks.7.0.skip_connection.weight"]),
    "circ128": dict(
        flags=dict(image_size=128, num_channels=32, num_res_blocks=1, class_cond=False, n_vars=4, in_channels=3,
                   **_COMMON),
        A=CIRCUIT, B=2, sub=4, wseed=8, iseed=23, rseed=33, kl_weight=0.3, respacing="ddim5", do_value=0.4,
        ddim=[], train_steps=0,
        grad_probe=["out.2.weight", "middle_block.1.proj_out.weight", "rep_emb.encoder.5.0.weight", "causal_mask.nonlinearities.3.net.2.bias"]),
}
ALL_CASES = {**MODEL_CASES, **SHIPPED_CASES}


def make_inputs(case):
    f = case["flags"]
    g = torch.Generator().manual_seed(case["iseed"])
    B, C, S = case["B"], f["in_channels"], f["image_size"]
    T = f["diffusion_steps"]
    return dict(
        x0=torch.rand(B, C, S, S, generator=g),
        noise=torch.randn(B, C, S, S, generator=g),
        t=torch.randint(0, T, (B,), generator=g),
        y=torch.randint(0, 10, (B,), generator=g),
        c=torch.rand(B, f["n_vars"], generator=g),
        w=torch.rand(B, generator=g) + 0.5,
        z=torch.randn(B, 512, generator=g),
    )
