"""Data path: this repo's dataset classes against the REAL reference's, item by item and bit for bit.

tests/golden/datasets_v1.npz holds the outputs of improved_diffusion/image_datasets.py (imported in place in the build
container by tests/golden/make_datasets_golden.py) on the seeded on-disk fixtures of tests/golden/dataset_fixture.py;
here the same fixtures are regenerated and read by causaldiffae_b200.image_datasets."""
import os

import numpy as np
import pytest
import torch

from tests.golden import dataset_fixture as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "datasets_v1.npz"))


@pytest.fixture(scope="module")
def roots(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("data"))
    return dict(mm=fx.make_morphomnist(os.path.join(tmp, "morphomnist")), pend=fx.make_pendulum(os.path.join(tmp, "pendulum")),
                circ=fx.make_circuit(os.path.join(tmp, "circuit")), cel=fx.make_celeba(os.path.join(tmp, "celeba")))


def check(ds, gold, tag):
    xs, cs, ys = [], [], []
    for i in range(len(ds)):
        x, d = ds[i]
        xs.append(np.asarray(x, dtype=np.float32))
        if "c" in d:
            cs.append(d["c"])
        if "y" in d:
            ys.append(d["y"])
    x = np.stack(xs)
    assert x.shape == gold[f"{tag}/x"].shape, (tag, x.shape, gold[f"{tag}/x"].shape)
    assert np.array_equal(x, gold[f"{tag}/x"]), f"{tag}: images differ from the reference loader"
    if f"{tag}/c" in gold.files:
        c = np.stack(cs)
        assert c.dtype == np.float32 and np.array_equal(c, gold[f"{tag}/c"]), f"{tag}: labels differ"
    else:
        assert not cs
    if f"{tag}/y" in gold.files:
        y = np.stack(ys)
        assert y.dtype == np.int64 and np.array_equal(y, gold[f"{tag}/y"]), f"{tag}: classes differ"
    # the resident arrays the CUDA gather consumes restate the same items: u8 / 255 (or / 127.5 - 1), NHWC -> NCHW
    images, c, y = ds.host_arrays()
    assert images.dtype == np.uint8 and images.shape[0] == len(ds)
    u = torch.from_numpy(np.ascontiguousarray(images)).float()
    xr = (u / 255.0 if ds.mode == 0 else u / 127.5 - 1).permute(0, 3, 1, 2).numpy()
    assert np.array_equal(xr, gold[f"{tag}/x"])


@pytest.mark.parametrize("shard,ns", [(0, 1), (1, 3)])
def test_morphomnist_items_match_reference(gold, roots, shard, ns):
    from causaldiffae_b200 import image_datasets as ds
    check(ds.MorphoMNISTLike(roots["mm"], columns=["thickness", "intensity"], train=True, shard=shard, num_shards=ns), gold,
          f"morphomnist/train/{shard}of{ns}")


def test_morphomnist_test_and_val_split(gold, roots):
    from causaldiffae_b200 import image_datasets as ds
    check(ds.MorphoMNISTLike(roots["mm"], columns=["thickness", "intensity"], train=False), gold, "morphomnist/test/0of1")
    val = ds.get_dataloader_morphomnist(roots["mm"], 2, "val", 0, 1).dataset
    assert np.array_equal(val.indices, gold["morphomnist/val/indices"])        # random_split(seed 42) subset
    check(val, gold, "morphomnist/val/0of1")


def test_idx_round_trip(tmp_path):
    from causaldiffae_b200 import image_datasets as ds
    a = np.arange(2 * 3 * 5, dtype=np.uint8).reshape(2, 3, 5)
    for name in ("a-idx3-ubyte.gz", "a-idx3-ubyte"):
        ds.save_idx(a, str(tmp_path / name))
        assert np.array_equal(ds.load_idx(str(tmp_path / name)), a)
    with open(tmp_path / "bad", "wb") as f:
        f.write(b"\x01\x02\x03\x04")
    with pytest.raises(ValueError):
        ds.load_idx(str(tmp_path / "bad"))


@pytest.mark.parametrize("split,shard,ns", [("train", 0, 1), ("train", 1, 3), ("test", 0, 1)])
def test_pendulum_items_match_reference(gold, roots, monkeypatch, split, shard, ns):
    from causaldiffae_b200 import image_datasets as ds
    order = [str(s) for s in gold[f"pendulum/listdir_{split}"]]
    real = os.listdir
    # the reference shards in os.listdir order, which depends on the file system: replay the recorded order
    monkeypatch.setattr(ds.os, "listdir", lambda p: order if os.path.basename(p) == split else real(p))
    d = ds.SyntheticLabeled(roots["pend"], split=split, shard=shard, num_shards=ns)
    assert sorted(order) == sorted(real(os.path.join(roots["pend"], split)))
    check(d, gold, f"pendulum/{split}/{shard}of{ns}")
    assert d[0][0].shape == (4, 96, 96)


@pytest.mark.parametrize("split,shard,ns", [("train", 0, 1), ("train", 1, 3), ("test", 0, 1)])
def test_circuit_items_match_reference(gold, roots, split, shard, ns):
    from causaldiffae_b200 import image_datasets as ds
    d = ds.CausalCircuit(roots["circ"], split, shard=shard, num_shards=ns)
    check(d, gold, f"circuit/{split}/{shard}of{ns}")
    assert d[0][0].shape == (3, 128, 128)


def test_image_dataset_matches_reference(gold, roots):
    from causaldiffae_b200 import image_datasets as ds
    files = ds._list_image_files_recursively(roots["cel"])
    assert [os.path.relpath(f, roots["cel"]) for f in files] == [str(s) for s in gold["celeba/files"]]
    names = [os.path.basename(p).split("_")[0] for p in files]
    classes = [{x: i for i, x in enumerate(sorted(set(names)))}[x] for x in names]
    check(ds.ImageDataset(64, files, classes=classes), gold, "celeba/64/0of1")
    check(ds.ImageDataset(64, files, classes=classes, shard=1, num_shards=2), gold, "celeba/64/1of2")


def test_loader_order_sharding_and_errors(roots):
    from causaldiffae_b200 import image_datasets as ds
    d = ds.MorphoMNISTLike(roots["mm"], columns=["thickness", "intensity"], train=True)
    ld = ds.ResidentLoader(d, 4, shuffle=True, generator=torch.Generator().manual_seed(5))
    assert len(ld) == len(d) // 4                                      # drop_last
    o1 = ld.epoch_order()
    assert sorted(o1.tolist()) == list(range(len(d)))                  # a permutation, redrawn every epoch
    assert torch.equal(o1, torch.randperm(len(d), generator=torch.Generator().manual_seed(5)))
    assert ds.ResidentLoader(d, 4, shuffle=False).epoch_order().tolist() == list(range(len(d)))
    # rank-strided shards partition the dataset (ref [shard:][::num_shards])
    parts = [ds.MorphoMNISTLike(roots["mm"], columns=["thickness", "intensity"], train=True, shard=r, num_shards=3) for r in range(3)]
    assert sum(len(p) for p in parts) == len(d)
    full = d.host_arrays()[0]
    for r, p in enumerate(parts):
        assert np.array_equal(p.host_arrays()[0], full[r::3])
    with pytest.raises(ValueError):
        next(ds.load_data(data_dir="", batch_size=2, image_size=28))
    with pytest.raises(ValueError):
        next(ds.load_data(data_dir="/nonexistent/foo", batch_size=2, image_size=28))
    if not torch.cuda.is_available():
        from causaldiffae_b200._lib import CdaeError
        with pytest.raises(CdaeError):                                 # batches are a CUDA gather: no CPU fallback
            next(iter(ld))
