"""Oracle: noise schedules, diffusion tables, respacing, timestep sampling (numpy float64 / ints).

Test infrastructure only (see oracle/__init__.py).  Every function restates the
reference algorithm and cites the file:line it follows.  Integer outputs here are the
bit-exact bar for the product (`timestep_map`, sampled `t`).
"""
import math

import numpy as np


def named_beta_schedule(name, n):
    """ref gaussian_diffusion.py:21-45 (linear: scaled Ho et al.; cosine: alpha_bar discretisation :48-65)."""
    if name == "linear":
        k = 1000 / n
        return np.linspace(k * 0.0001, k * 0.02, n, dtype=np.float64)
    if name == "cosine":
        def abar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        out = []
        for i in range(n):
            out.append(min(1 - abar((i + 1) / n) / abar(i / n), 0.999))
        return np.array(out)
    raise NotImplementedError(f"unknown beta schedule: {name}")


def diffusion_tables(betas):
    """All float64 per-timestep tables of GaussianDiffusion.__init__ (ref gaussian_diffusion.py:134-179)."""
    betas = np.array(betas, dtype=np.float64)
    assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    ac_next = np.append(ac[1:], 0.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        betas=betas,
        alphas_cumprod=ac,
        alphas_cumprod_prev=ac_prev,
        alphas_cumprod_next=ac_next,
        sqrt_alphas_cumprod=np.sqrt(ac),
        sqrt_one_minus_alphas_cumprod=np.sqrt(1.0 - ac),
        log_one_minus_alphas_cumprod=np.log(1.0 - ac),
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
        sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=post_var,
        posterior_log_variance_clipped=np.log(np.append(post_var[1], post_var[1:])),
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )


def space_timesteps(num_timesteps, section_counts):
    """Retained-step set (ref respace.py:7-61). "ddimN": first integer stride with exactly N steps;
    otherwise per-section fractional stride with Python round() (banker's rounding)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(s) for s in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, cnt in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < cnt:
            raise ValueError(f"cannot divide section of {size} steps into {cnt}")
        frac = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            steps.append(start + round(cur))
            cur += frac
        start += size
    return set(steps)


def respaced_betas(base_betas, use_timesteps):
    """SpacedDiffusion.__init__ (ref respace.py:74-88): new_beta_i = 1 - abar_i / abar_last_kept; timestep_map."""
    use = set(use_timesteps)
    ac = diffusion_tables(base_betas)["alphas_cumprod"]
    last, new_betas, tmap = 1.0, [], []
    for i, a in enumerate(ac):
        if i in use:
            new_betas.append(1 - a / last)
            last = a
            tmap.append(i)
    return np.array(new_betas), tmap


def uniform_sample_t(num_timesteps, batch_size):
    """UniformSampler.sample (ref resample.py:44-69): np.random.choice with p = 1/T on the GLOBAL numpy RNG;
    weights are 1/(T*p[idx]) (== 1.0). Returns (int64 indices, float32 weights)."""
    w = np.ones([num_timesteps])
    p = w / np.sum(w)
    idx = np.random.choice(len(p), size=(batch_size,), p=p)
    weights = 1 / (len(p) * p[idx])
    return idx.astype(np.int64), weights.astype(np.float32)


def kl_weight_schedule(step, total_steps=50000, initial=0.0, final=1.0):
    """TrainLoop.linear_kl_weight_scheduler (ref train_util.py:176-187), called with (step, 50000, 0, 1) at :213."""
    if step >= total_steps:
        return final
    if step <= 0:
        return initial
    if total_steps <= 1:
        return final
    t = step / (total_steps - 1)
    return (1.0 - t) * initial + t * final


def channel_mult_for(image_size):
    """ref script_util.py:140-153."""
    table = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 2, 4, 4), 96: (1, 2, 3, 4), 64: (1, 2, 3, 4),
             32: (1, 2, 2, 2), 28: (1, 2, 2)}
    if image_size not in table:
        raise ValueError(f"unsupported image size: {image_size}")
    return table[image_size]


def attention_ds_for(image_size, attention_resolutions):
    """ref script_util.py:155-157 (integer division image_size // res)."""
    return tuple(image_size // int(r) for r in attention_resolutions.split(","))


def encoder_hidden_dims(image_size, n_vars):
    """Documented oracle patch 1: the shipped encoder (ref nn.py:39-43) is only shape-correct
    for 65..128 px; keep its last L = ceil(log2(S)) - 1 stages so the final map is 2x2 for every supported S."""
    base = [16, 32, 32, 64, 64, 128] if n_vars == 4 else [16, 32, 64, 128]
    L = int(math.ceil(math.log2(image_size))) - 1
    return base[-L:] if L <= len(base) else base


DAGS = {
    # adjacency A[j, i] = 1 iff j is a parent of i (ref unet.py:571-578, image_causaldae_test.py:331,479,773)
    "morphomnist": [[0, 1], [0, 0]],
    "circuit": [[0, 1, 1, 1], [0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 0, 0]],
    "pendulum": [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]],
}


def default_dag(n_vars):
    """ref unet.py:571-575: n_vars==2 -> MorphoMNIST graph, else the Circuit graph."""
    return DAGS["morphomnist"] if n_vars == 2 else DAGS["circuit"]


def topo_order(A):
    """Kahn topological order with smallest-index tie-break. The reference has no sort (Q3); all shipped A are
    strictly upper triangular so this must return the identity permutation on them."""
    A = np.asarray(A)
    n = A.shape[0]
    indeg = (A != 0).sum(axis=0).astype(int).tolist()
    done, order = [False] * n, []
    for _ in range(n):
        nxt = next((i for i in range(n) if not done[i] and indeg[i] == 0), None)
        if nxt is None:
            raise ValueError("adjacency is not a DAG")
        done[nxt] = True
        order.append(nxt)
        for i in range(n):
            if A[nxt, i] != 0:
                indeg[i] -= 1
    return order
