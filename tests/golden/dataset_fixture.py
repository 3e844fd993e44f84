"""Tiny synthetic datasets in the on-disk formats the reference reads (MorphoMNIST idx.gz + csv, Pendulum PNG files,
CausalCircuit npz shards with PNG bytes), generated from fixed seeds.  Shared by make_datasets_golden.py (which runs the
REAL reference loaders on them) and tests/test_datasets_cpu.py / the GPU loader test (which run this repo's loaders)."""
import gzip
import io
import os
import struct

import numpy as np


def _smooth(rng, n, h, w, c):
    """smooth random images (low-frequency cosines): uint8 [n,h,w,c]"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.zeros((n, h, w, c))
    for i in range(n):
        for k in range(c):
            a, b, ph = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6.28, size=2)
            out[i, :, :, k] = 0.5 + 0.25 * np.cos(a * yy * 6.28 / h + ph[0]) + 0.25 * np.cos(b * xx * 6.28 / w + ph[1])
    return np.clip(out * 255, 0, 255).astype(np.uint8)


def _write_idx_gz(path, arr):
    with gzip.open(path, "wb") as f:
        f.write(struct.pack(">HBB", 0, 0x08, arr.ndim))
        f.write(struct.pack(">" + "I" * arr.ndim, *arr.shape))
        f.write(arr.astype(np.uint8).tobytes())


def make_morphomnist(root, n_train=11, n_test=20, seed=1):
    os.makedirs(root, exist_ok=True)
    rng = np.random.RandomState(seed)
    for prefix, n in (("train", n_train), ("t10k", n_test)):
        imgs = _smooth(rng, n, 28, 28, 1)[..., 0]
        labels = rng.randint(0, 10, size=n).astype(np.uint8)
        _write_idx_gz(os.path.join(root, f"{prefix}-images-idx3-ubyte.gz"), imgs)
        _write_idx_gz(os.path.join(root, f"{prefix}-labels-idx1-ubyte.gz"), labels)
        with open(os.path.join(root, f"{prefix}-morpho.csv"), "w") as f:
            f.write("index,area,length,thickness,slant,width,height,intensity\n")
            for i in range(n):
                v = rng.uniform(0, 1, size=7)
                f.write(f"{i},{v[0]*100:.6f},{v[1]*40:.6f},{1.5 + 4 * v[2]:.15f},{v[3]:.6f},{v[4]*20:.6f},{v[5]*20:.6f},"
                        f"{60 + 190 * v[6]:.15f}\n")
    return root


def make_pendulum(root, n_train=7, n_test=3, seed=2):
    from PIL import Image
    rng = np.random.RandomState(seed)
    for split, n in (("train", n_train), ("test", n_test)):
        d = os.path.join(root, split)
        os.makedirs(d, exist_ok=True)
        imgs = _smooth(rng, n, 96, 96, 4)
        for i in range(n):
            lab = [int(rng.randint(-40, 44)), int(rng.randint(60, 148)), int(rng.randint(3, 12)), int(rng.randint(3, 19))]
            Image.fromarray(imgs[i], mode="RGBA").save(os.path.join(d, "a_" + "_".join(map(str, lab)) + ".png"))
    return root


def make_circuit(root, n_per_shard=3, n_test=4, size=160, seed=3):
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    rng = np.random.RandomState(seed)

    def shard(n):
        imgs = _smooth(rng, n, size, size, 3)
        raw = []
        for i in range(n):
            buf = io.BytesIO()
            Image.fromarray(imgs[i], mode="RGB").save(buf, format="PNG")
            raw.append(buf.getvalue())
        blobs = np.array(raw, dtype=np.bytes_).reshape(n, 1)     # fixed-width byte strings: loads without pickle
        lat = rng.uniform(0, 1, size=(n, 1, 4)).astype(np.float32)
        return blobs, lat

    for k in range(5):
        b, l = shard(n_per_shard)
        np.savez(os.path.join(root, f"train-{k}.npz"), imgs=b, original_latents=l)
    b, l = shard(n_test)
    np.savez(os.path.join(root, "test.npz"), imgs=b, original_latents=l)
    return root


def make_celeba(root, n=5, seed=4):
    from PIL import Image
    os.makedirs(os.path.join(root, "sub"), exist_ok=True)
    rng = np.random.RandomState(seed)
    imgs = _smooth(rng, n, 150, 200, 3)
    for i in range(n):
        d = root if i % 2 == 0 else os.path.join(root, "sub")
        Image.fromarray(imgs[i], mode="RGB").save(os.path.join(d, f"{'cat' if i < 2 else 'dog'}_{i}.png"))
    return root
