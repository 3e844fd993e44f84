"""HBM-resident data path on the GPU: the gather kernel (through the C ABI) is bit-exact against the host restatement of
the reference's per-item arithmetic, and load_data() batches equal the reference loader's golden items."""
import os

import numpy as np
import pytest
import torch

from tests.golden import dataset_fixture as fx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,H,W,C,L,B,mode", [(50, 28, 28, 1, 2, 16, 0), (9, 96, 96, 4, 4, 32, 0), (7, 128, 128, 3, 4, 5, 0),
                                              (6, 64, 64, 3, 0, 6, 1), (300, 6, 10, 2, 1, 257, 0)])
def test_gather_images_bit_exact(n, H, W, C, L, B, mode):
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(n + C)
    im = torch.randint(0, 256, (n, H, W, C), generator=g, dtype=torch.uint8)
    lab = torch.randn(n, L, generator=g) if L else None
    idx = torch.randint(0, n, (B,), generator=g)
    x, c = ops.gather_images(im.cuda(), idx.cuda(), labels=lab.cuda() if L else None, mode=mode)
    u = im[idx].float()
    ref = (u / 255.0 if mode == 0 else u / 127.5 - 1).permute(0, 3, 1, 2)
    assert x.shape == (B, C, H, W) and torch.equal(x.cpu(), ref)
    if L:
        assert torch.equal(c.cpu(), lab[idx])
    else:
        assert c is None


def test_gather_images_rejects_bad_shapes():
    from causaldiffae_b200 import ops
    from causaldiffae_b200._lib import CdaeError
    with pytest.raises(CdaeError):
        ops.gather_images(torch.zeros(2, 3, 3, 1, dtype=torch.uint8, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
    with pytest.raises(CdaeError):
        ops.gather_images(torch.zeros(2, 4, 4, 5, dtype=torch.uint8, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))


def test_load_data_batches_match_reference_items(tmp_path):
    """load_data(...) through the resident loader: every batch row equals the reference loader's item (golden)"""
    from causaldiffae_b200 import image_datasets as ds
    gold = np.load(os.path.join(ROOT, "tests", "golden", "datasets_v1.npz"))
    # CausalCircuit: the reference does not shuffle -> batches are the golden items in order, drop_last
    circ = fx.make_circuit(str(tmp_path / "circuit"))
    it = ds.load_data(data_dir=circ, batch_size=4, image_size=128, split="train")
    gx, gc = gold["circuit/train/0of1/x"], gold["circuit/train/0of1/c"]
    nb = len(gx) // 4
    for k in range(nb + 1):                                 # one batch past the epoch end: the generator restarts
        x, cond = next(it)
        assert x.is_cuda and x.dtype == torch.float32 and x.shape == (4, 3, 128, 128)
        j = (k % nb) * 4
        assert np.array_equal(x.cpu().numpy(), gx[j:j + 4]) and np.array_equal(cond["c"].cpu().numpy(), gc[j:j + 4])
    # MorphoMNIST: shuffled; each batch row must be one golden item with its labels, an epoch has no repeats
    mm = fx.make_morphomnist(str(tmp_path / "morphomnist"))
    it = ds.load_data(data_dir=mm, batch_size=3, image_size=28, class_cond=True)
    gx, gc, gy = (gold[f"morphomnist/train/0of1/{k}"] for k in "xcy")
    seen = []
    for _ in range(len(gx) // 3):
        x, cond = next(it)
        assert x.shape == (3, 1, 28, 28) and cond["y"].dtype == torch.int64
        for r in range(3):
            hit = [i for i in range(len(gx)) if np.array_equal(x[r].cpu().numpy(), gx[i])]
            assert len(hit) == 1
            assert np.array_equal(cond["c"][r].cpu().numpy(), gc[hit[0]]) and int(cond["y"][r]) == int(gy[hit[0]])
            seen.append(hit[0])
    assert len(set(seen)) == len(seen)
    # Pendulum items feed a TrainLoop-shaped consumer: NCHW fp32 in [0,1], c [B,4]
    pend = fx.make_pendulum(str(tmp_path / "pendulum"))
    x, cond = next(ds.load_data(data_dir=pend, batch_size=2, image_size=96))
    assert x.shape == (2, 4, 96, 96) and float(x.min()) >= 0 and float(x.max()) <= 1 and cond["c"].shape == (2, 4)


def test_image_train_recipe_end_to_end(tmp_path, monkeypatch):
    """scripts/image_train.py:20-76 of the reference, against this package: load_data -> create_model_and_diffusion ->
    TrainLoop(...).run_loop() on the Pendulum configuration (96x96 RGBA, 4 causal variables, masking=True)."""
    from causaldiffae_b200 import image_datasets as ds, script_util as su, dist_util, logger
    from causaldiffae_b200.resample import create_named_schedule_sampler
    from causaldiffae_b200.train_util import TrainLoop
    pend = fx.make_pendulum(str(tmp_path / "pendulum"), n_train=9)
    dist_util.setup_dist()
    logger.configure(dir=str(tmp_path / "log"), format_strs=[])
    monkeypatch.setenv("DIFFUSION_BLOB_LOGDIR", str(tmp_path / "log"))
    flags = {**su.model_and_diffusion_defaults(), **dict(image_size=96, in_channels=4, num_channels=64, num_res_blocks=1,
                                                         n_vars=4, rep_cond=True, causal_modeling=True, masking=True,
                                                         learn_sigma=False, class_cond=False)}
    model, diffusion = su.create_model_and_diffusion(**flags)
    model.to(dist_util.dev())
    sampler = create_named_schedule_sampler("uniform", diffusion)
    data = ds.load_data(data_dir=pend, batch_size=4, image_size=96, class_cond=False)
    loop = TrainLoop(model=model, diffusion=diffusion, data=data, batch_size=4, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                     log_interval=2, save_interval=3, resume_checkpoint="", use_fp16=False, fp16_scale_growth=1e-3,
                     schedule_sampler=sampler, weight_decay=0.0, lr_anneal_steps=4, rep_cond=True, n_vars=4,
                     causal_modeling=True, flow_based=False, in_channels=4, masking=True)
    w0 = model.out[2].weight.detach().clone()
    loop.run_loop()
    assert loop.step == 4 and bool(torch.isfinite(loop.last_loss))
    files = sorted(os.listdir(str(tmp_path / "log")))
    assert "model000000.pt" in files and "model000003.pt" in files and "ema_0.9999_000003.pt" in files, files
    sd = torch.load(str(tmp_path / "log" / "model000003.pt"), map_location="cpu")
    assert set(sd) == set(model.state_dict()) and all(bool(torch.isfinite(v.float()).all()) for v in sd.values())
    assert not torch.equal(w0.cpu(), sd["out.2.weight"])          # the zero-initialised out conv has started to move
