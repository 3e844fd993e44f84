"""Generate tests/golden/datasets_v1.npz: per-item outputs of the REAL reference dataset classes
(/root/reference/improved_diffusion/image_datasets.py, imported in place in the build container) on the synthetic
on-disk fixtures of dataset_fixture.py.

    python tests/golden/make_datasets_golden.py

Shims (test infrastructure only): blobfile / mpi4py stand-ins (oracle/refshim.py); the reference calls `io.load_idx` on
the stdlib `io` module because its `datasets.morphomnist.io` import is commented out, so the module attribute `io` is
replaced by a namespace with `load_idx` (a plain IDX parser) and `BytesIO`; CausalCircuit hard-codes
`../datasets/causal_circuit/`, so the fixture is generated there relative to a scratch working directory.
"""
import gzip
import io as stdio
import os
import struct
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402
from tests.golden import dataset_fixture as fx  # noqa: E402


def _load_idx(path):
    with gzip.open(path, "rb") as f:
        _, code, ndim = struct.unpack(">HBB", f.read(4))
        shape = struct.unpack(">" + "I" * ndim, f.read(4 * ndim))
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(shape)


def main():
    refshim.load()
    sys.modules["blobfile"].listdir = os.listdir
    sys.modules["blobfile"].isdir = os.path.isdir
    sys.modules["blobfile"].basename = os.path.basename
    import improved_diffusion.image_datasets as ref
    ref.io = types.SimpleNamespace(load_idx=_load_idx, BytesIO=stdio.BytesIO)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        mm = fx.make_morphomnist(os.path.join(tmp, "morphomnist"))
        pend = fx.make_pendulum(os.path.join(tmp, "pendulum"))
        circ = fx.make_circuit(os.path.join(tmp, "datasets", "causal_circuit"))
        cel = fx.make_celeba(os.path.join(tmp, "celeba"))
        os.makedirs(os.path.join(tmp, "scripts"), exist_ok=True)
        os.chdir(os.path.join(tmp, "scripts"))          # '../datasets/causal_circuit' resolves to the fixture
        # os.listdir order is filesystem dependent and the reference shards in that order: record it
        out["pendulum/listdir_train"] = np.array(os.listdir(os.path.join(pend, "train")))
        out["pendulum/listdir_test"] = np.array(os.listdir(os.path.join(pend, "test")))

        def dump(tag, ds):
            xs, cs, ys = [], [], []
            for i in range(len(ds)):
                x, d = ds[i]
                xs.append(np.asarray(x, dtype=np.float32))
                if "c" in d:
                    cs.append(np.asarray(d["c"], dtype=np.float32))
                if "y" in d:
                    ys.append(np.asarray(d["y"], dtype=np.int64))
            out[f"{tag}/x"] = np.stack(xs)
            if cs:
                out[f"{tag}/c"] = np.stack(cs)
            if ys:
                out[f"{tag}/y"] = np.stack(ys)

        for shard, ns in ((0, 1), (1, 3)):
            dump(f"morphomnist/train/{shard}of{ns}", ref.MorphoMNISTLike(mm, columns=["thickness", "intensity"], train=True,
                                                                         shard=shard, num_shards=ns))
            dump(f"pendulum/train/{shard}of{ns}", ref.SyntheticLabeled(pend, split="train", shard=shard, num_shards=ns))
            dump(f"circuit/train/{shard}of{ns}", ref.CausalCircuit(circ, "train", shard=shard, num_shards=ns))
        dump("morphomnist/test/0of1", ref.MorphoMNISTLike(mm, columns=["thickness", "intensity"], train=False))
        dump("pendulum/test/0of1", ref.SyntheticLabeled(pend, split="test"))
        dump("circuit/test/0of1", ref.CausalCircuit(circ, "test"))
        val = ref.get_dataloader_morphomnist(mm, 2, "val", 0, 1).dataset
        out["morphomnist/val/indices"] = np.asarray(val.indices, dtype=np.int64)
        dump("morphomnist/val/0of1", val)
        files = ref._list_image_files_recursively(cel)
        out["celeba/files"] = np.array([os.path.relpath(f, cel) for f in files])
        names = [os.path.basename(p).split("_")[0] for p in files]
        classes = [{x: i for i, x in enumerate(sorted(set(names)))}[x] for x in names]
        dump("celeba/64/0of1", ref.ImageDataset(64, files, classes=classes))
        dump("celeba/64/1of2", ref.ImageDataset(64, files, classes=classes, shard=1, num_shards=2))
        os.chdir(HERE)
    path = os.path.join(HERE, "datasets_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", len(out), "arrays")


if __name__ == "__main__":
    main()
