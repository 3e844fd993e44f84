"""Checkpoint wire format + resume (ref train_util.py:128-169,319-398): the files TrainLoop.save() writes
carry the reference's names and state_dict keys, a reference-style checkpoint (plain OIHW fp32 state_dict written by
torch.save) loads, and a resumed loop continues exactly where the saved one stopped (weights, EMA, Adam moments, step
counter) so that its next optimisation step matches the uninterrupted run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FLAGS = dict(image_size=32, num_channels=64, num_res_blocks=1, class_cond=False, rep_cond=True, n_vars=4,
             causal_modeling=True, in_channels=3, learn_sigma=False, rescale_timesteps=False,
             rescale_learned_sigmas=False, diffusion_steps=100)


def _loop(model, diff, resume=""):
    from causaldiffae_b200.train_util import TrainLoop
    return TrainLoop(model=model, diffusion=diff, data=None, batch_size=4, microbatch=-1, lr=1e-3, ema_rate="0.9,0.99",
                     log_interval=10 ** 9, save_interval=10 ** 9, resume_checkpoint=resume, rep_cond=True, n_vars=4,
                     causal_modeling=True, in_channels=3)


def _step(loop, k):
    g = torch.Generator().manual_seed(100 + k)
    x, c = torch.rand(4, 3, 32, 32, generator=g), torch.rand(4, 4, generator=g)
    np.random.seed(200 + k)
    torch.manual_seed(300 + k)
    loop.run_step(x, {"c": c})
    loop.step += 1


def test_save_resume_round_trip(tmp_path, monkeypatch):
    from causaldiffae_b200 import script_util as su, dist_util, logger
    from oracle import model as om
    full = {**su.model_and_diffusion_defaults(), **FLAGS}
    dist_util.setup_dist()
    logger.configure(dir=str(tmp_path), format_strs=[])
    monkeypatch.setenv("DIFFUSION_BLOB_LOGDIR", str(tmp_path))
    sd0 = om.seeded_state_dict(om.config_from_flags(**full), seed=3)           # reference-format state_dict (OIHW fp32)
    # a checkpoint exactly as the reference writes it: torch.save of the state_dict under model%06d.pt
    ref_ckpt = os.path.join(str(tmp_path), "pretrained", "model000000.pt")
    os.makedirs(os.path.dirname(ref_ckpt))
    torch.save(sd0, ref_ckpt)
    model, diff = su.create_model_and_diffusion(**full)
    model.cuda()
    loop = _loop(model, diff, resume=ref_ckpt)
    assert loop.resume_step == 0
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), sd0[k]), f"{k} not restored from the reference-style checkpoint"
    for k in range(3):
        _step(loop, k)
    loop.save()
    files = sorted(os.listdir(str(tmp_path)))
    for name in ("model000003.pt", "ema_0.9_000003.pt", "ema_0.99_000003.pt", "opt000003.pt", "ema_checkpoint.pt"):
        assert name in files, files
    saved = torch.load(os.path.join(str(tmp_path), "model000003.pt"), map_location="cpu")
    assert set(saved) == set(sd0) and all(saved[k].shape == sd0[k].shape and saved[k].dtype == sd0[k].dtype for k in sd0)
    assert all(v.is_contiguous() for v in saved.values())                        # plain OIHW tensors, loadable by the reference
    ema_saved = torch.load(os.path.join(str(tmp_path), "ema_0.99_000003.pt"), map_location="cpu")
    assert set(ema_saved) == set(sd0)
    # continue the uninterrupted run by one more step
    _step(loop, 3)
    cont = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    cont_ema = loop.ema_params[1][0].clone()
    # resume from the files in a fresh model / loop
    model2, diff2 = su.create_model_and_diffusion(**full)
    model2.cuda()
    loop2 = _loop(model2, diff2, resume=os.path.join(str(tmp_path), "model000003.pt"))
    assert loop2.resume_step == 3 and loop2.opt.step_count == loop.opt.step_count - 1
    for k, v in model2.state_dict().items():
        assert torch.equal(v.cpu(), saved[k]), k
    ema2 = loop2.engine.export_state(loop2.ema_params[1][0])
    for k, v in ema2.items():
        assert torch.equal(v.cpu(), ema_saved[k]), f"EMA {k} not restored"
    loop2.step = 0
    diff2.kl_weight = diff.kl_weight
    _step(loop2, 3)
    res = {k: v.detach().float().cpu() for k, v in model2.state_dict().items()}
    # same data, t, noise, xi and optimizer state: only the atomics' summation order differs between the two runs
    num = sum(float((res[k] - cont[k]).double().pow(2).sum()) for k in cont if cont[k].dtype.is_floating_point)
    den = sum(float((cont[k] - saved[k].float()).double().pow(2).sum()) for k in cont if cont[k].dtype.is_floating_point)
    assert den > 0 and (num / den) ** 0.5 < 2e-2, (num, den)                      # the resumed UPDATE equals the continued one
    assert float((loop2.ema_params[1][0] - cont_ema).norm() / cont_ema.norm()) < 1e-5
    assert loop2.step + loop2.resume_step == 4
