"""Data path of the reference (improved_diffusion/image_datasets.py) re-designed for a B200: HBM-resident datasets.

The reference decodes one item at a time in a single DataLoader worker (PIL -> ToTensor -> collate) and ships every batch
over PCIe.  Every dataset it supports is tiny next to 180 GB of HBM (MorphoMNIST 47 MB, Pendulum 0.26 GB, CausalCircuit
1.7 GB as uint8), so here a dataset is decoded ONCE on the host, stored as uint8 NHWC in HBM, and a training batch is one
launch of `cdae_gather_images` (csrc/dataset.cu): fp32 NCHW = u8 / 255 - bit-identical to the reference's ToTensor
arithmetic - with the labels gathered alongside.  Same public surface as the reference:

    load_data(data_dir=..., batch_size=..., image_size=..., class_cond=False, split="train", deterministic=False)
        -> generator of (images [B,C,H,W] fp32, {"c": [B,n] fp32, "y": [B] int64})          ref :69-126
    MorphoMNISTLike / SyntheticLabeled (Pendulum) / CausalCircuit / ImageDataset, get_dataloader_{morphomnist,pendulum,circuit}

Item-level semantics kept exactly (checked against the reference classes in tests/test_datasets_cpu.py): rank-strided
sharding `[shard:][::num_shards]` (:145-146,256-263,353-357,459-460), label normalisation ((l - a) / b in fp32,
:372-375), the [3,2,1,0] latent permutation of CausalCircuit (:469-470), Resize(128) of circuit images (:463), the
0.9/0.1 `random_split(seed 42)` validation subset (:320-326), drop_last batches, shuffle everywhere but CausalCircuit.
Deviations forced by reference bugs: the reference calls `io.load_idx` on the stdlib `io` module (its `datasets.morphomnist`
package is not in the repo) - the IDX reader is implemented here; CausalCircuit reads `<root>/train-k.npz` (falling back to
the reference's hard-coded `../datasets/causal_circuit/`).  `__getitem__` is host code (numpy) for API compatibility and
tests; the training path (`ResidentLoader`) is the CUDA gather and raises without a GPU - there is no CPU fallback.
"""
import gzip
import io as _io
import os
import struct

import numpy as np
import torch as th

from . import _lib, ops


# ---------------------------------------------------------------------------------------------- IDX files (MorphoMNIST)
_IDX_DTYPES = {0x08: np.uint8, 0x09: np.int8, 0x0B: ">i2", 0x0C: ">i4", 0x0D: ">f4", 0x0E: ">f8"}


def load_idx(path):
    """IDX (MNIST) reader, optionally gzip-compressed: magic = 0, 0, dtype code, ndim; big-endian uint32 dims; data."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        zero, code, ndim = struct.unpack(">HBB", f.read(4))
        if zero != 0 or code not in _IDX_DTYPES:
            raise ValueError(f"{path}: not an IDX file (magic {zero:#x} {code:#x})")
        shape = struct.unpack(">" + "I" * ndim, f.read(4 * ndim))
        data = np.frombuffer(f.read(), dtype=_IDX_DTYPES[code])
    return data.reshape(shape)


def save_idx(data, path):
    codes = {np.dtype(np.uint8): 0x08, np.dtype(np.int8): 0x09, np.dtype(">i2"): 0x0B, np.dtype(">i4"): 0x0C,
             np.dtype(">f4"): 0x0D, np.dtype(">f8"): 0x0E}
    data = np.asarray(data)
    be = data.dtype.newbyteorder(">") if data.dtype.itemsize > 1 else data.dtype
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(struct.pack(">HBB", 0, codes[np.dtype(be)], data.ndim))
        f.write(struct.pack(">" + "I" * data.ndim, *data.shape))
        f.write(data.astype(be).tobytes())


def _get_paths(root_dir, train):
    prefix = "train" if train else "t10k"
    return (os.path.join(root_dir, prefix + "-images-idx3-ubyte.gz"), os.path.join(root_dir, prefix + "-labels-idx1-ubyte.gz"),
            os.path.join(root_dir, prefix + "-morpho.csv"))


def load_morphomnist_like(root_dir, train=True, columns=None):
    """ref :197-218 -> (images uint8 [n,28,28], labels uint8 [n], metrics DataFrame indexed by 'index')"""
    import pandas as pd
    images_path, labels_path, metrics_path = _get_paths(root_dir, train)
    images, labels = load_idx(images_path), load_idx(labels_path)
    usecols = ["index"] + list(columns) if columns is not None and "index" not in columns else columns
    metrics = pd.read_csv(metrics_path, usecols=usecols, index_col="index")
    return images, labels, metrics


def save_morphomnist_like(images, labels, metrics, root_dir, train):
    """ref :221-238"""
    assert len(images) == len(labels) == len(metrics)
    images_path, labels_path, metrics_path = _get_paths(root_dir, train)
    os.makedirs(root_dir, exist_ok=True)
    save_idx(images, images_path)
    save_idx(labels, labels_path)
    metrics.to_csv(metrics_path, index_label="index")


# ---------------------------------------------------------------------------------------------- datasets
class _Resident:
    """Common base: a dataset is (images uint8 [n,H,W,C], c fp32 [n,L] or None, y int64 [n] or None) on the host;
    `host_arrays()` decodes everything once, `__getitem__` restates the reference's per-item result from them."""
    mode = 0          # 0: u8 / 255 (ToTensor); 1: u8 / 127.5 - 1 (ImageDataset)

    def host_arrays(self):
        raise NotImplementedError

    def __len__(self):
        return self._n

    def __getitem__(self, idx):
        images, c, y = self.host_arrays()
        u = th.from_numpy(np.array(images[idx])).float()
        img = (u / 255.0) if self.mode == 0 else (u / 127.5 - 1)
        out = {}
        if y is not None:
            out["y"] = np.array(y[idx], dtype=np.int64)
        if c is not None:
            out["c"] = np.array(c[idx], dtype=np.float32)
        return img.permute(2, 0, 1), out


class MorphoMNISTLike(_Resident):
    """ref :241-296.  item = (image/255 [1,28,28], {"y": label, "c": [thickness, intensity] (raw csv values)})"""

    def __init__(self, root_dir, train=True, columns=None, shard=0, num_shards=1):
        self.root_dir, self.train = root_dir, train
        images, labels, metrics_df = load_morphomnist_like(root_dir, train, columns)
        self.images = np.ascontiguousarray(images[shard:][::num_shards])
        self.labels = np.ascontiguousarray(labels[shard:][::num_shards])
        if columns is None:
            columns = metrics_df.columns
        self.metrics = {col: np.asarray(metrics_df[col])[shard:][::num_shards] for col in columns}
        self.columns = columns
        self.scale = {"thickness": [3.4, 2.4], "intensity": [161, 94]}
        self.gaussian_scale = {"thickness": [2.5, 0.63], "intensity": [158.0, 48.4]}
        self._n = len(self.images)

    def host_arrays(self):
        c = np.stack([self.metrics["thickness"], self.metrics["intensity"]], axis=1).astype(np.float32)
        return self.images[..., None], c, self.labels.astype(np.int64)


class SyntheticLabeled(_Resident):
    """Pendulum (ref :344-391): <root>/<split>/a_<i>_<j>_<k>_<l>.png, RGBA 96x96; c = (label - scale[:,0]) / scale[:,1]."""

    def __init__(self, root, split="train", shard=0, num_shards=1):
        root = root + "/" + split
        imgs = os.listdir(root)
        self.dataset = split
        self.imgs = [os.path.join(root, k) for k in imgs][shard:][::num_shards]
        self.imglabel = np.asarray([list(map(int, k[:-4].split("_")[1:])) for k in imgs])[shard:][::num_shards]
        self.shard, self.num_shards = shard, num_shards
        self.scale = np.array([[2, 42], [104, 44], [7.5, 4.5], [11, 8]])
        self._n = len(self.imgs)
        self._cache = None

    def host_arrays(self):
        if self._cache is None:
            from PIL import Image
            ims = []
            for pth in self.imgs:
                with Image.open(pth) as im:
                    a = np.asarray(im)
                ims.append(a[..., None] if a.ndim == 2 else a)
            images = np.stack(ims) if ims else np.zeros((0, 96, 96, 4), np.uint8)
            # the reference normalises in fp32: int64 tensor element - python float -> fp32, / python float -> fp32 (:372-375)
            lab = th.from_numpy(np.asarray(self.imglabel).reshape(-1, 4).astype(np.int64))
            sc = th.from_numpy(self.scale.astype(np.float32))
            c = ((lab.float() - sc[:, 0]) / sc[:, 1]).numpy()
            self._cache = (images, c, None)
        return self._cache


class CausalCircuit(_Resident):
    """ref :411-482: npz shards with PNG bytes `imgs[:,0]` and latents `original_latents[:,0,:]`; Resize(128); c = latents[[3,2,1,0]]."""

    def __init__(self, root, dataset="train", shard=0, num_shards=1, resolution=128):
        self.dataset, self.resolution = dataset, resolution
        names = ["test.npz"] if dataset == "test" else [f"train-{k}.npz" for k in range(5)]
        blobs, labels = [], []
        for name in names:
            path = os.path.join(root, name)
            if not os.path.exists(path):
                path = os.path.join("../datasets/causal_circuit", name)      # the reference's hard-coded location
            data = np.load(path)
            lat, temp = data["original_latents"][:, 0, :], data["imgs"][:, 0]
            for i in range(len(temp)):
                blobs.append(temp[i]); labels.append(lat[i])
        self.labels = np.asarray(labels)[shard:][::num_shards]
        self.blobs = blobs[shard:][::num_shards]
        self._n = len(self.blobs)
        self._cache = None

    @staticmethod
    def _resize_shorter(im, size):
        """torchvision.transforms.Resize(int) on a PIL image: shorter edge -> size, bilinear (PIL's resize is area-aware)"""
        from PIL import Image
        w, h = im.size
        short, long = (w, h) if w <= h else (h, w)
        if short == size:
            return im
        new_short, new_long = size, int(size * long / short)
        nw, nh = (new_short, new_long) if w <= h else (new_long, new_short)
        return im.resize((nw, nh), Image.BILINEAR)

    def host_arrays(self):
        if self._cache is None:
            from PIL import Image
            ims = []
            for blob in self.blobs:
                raw = blob.tobytes() if hasattr(blob, "tobytes") and not isinstance(blob, (bytes, bytearray)) else blob
                im = self._resize_shorter(Image.open(_io.BytesIO(raw)), self.resolution)
                a = np.asarray(im)
                ims.append(a[..., None] if a.ndim == 2 else a)
            images = np.stack(ims) if ims else np.zeros((0, self.resolution, self.resolution, 3), np.uint8)
            c = np.asarray(self.labels, dtype=np.float64).reshape(-1, 4)[:, [3, 2, 1, 0]].astype(np.float32)
            self._cache = (images, c, None)
        return self._cache


def _list_image_files_recursively(data_dir):
    """ref :129-138"""
    results = []
    for entry in sorted(os.listdir(data_dir)):
        full_path = os.path.join(data_dir, entry)
        ext = entry.split(".")[-1]
        if "." in entry and ext.lower() in ["jpg", "jpeg", "png", "gif"]:
            results.append(full_path)
        elif os.path.isdir(full_path):
            results.extend(_list_image_files_recursively(full_path))
    return results


class ImageDataset(_Resident):
    """ref :141-183: BOX-halve while >= 2x, BICUBIC to the shorter edge, centre crop, RGB, u8/127.5 - 1; y = class index."""
    mode = 1

    def __init__(self, resolution, image_paths, classes=None, shard=0, num_shards=1):
        self.resolution = resolution
        self.local_images = image_paths[shard:][::num_shards]
        self.local_classes = None if classes is None else classes[shard:][::num_shards]
        self._n = len(self.local_images)
        self._cache = None

    def host_arrays(self):
        if self._cache is None:
            from PIL import Image
            R = self.resolution
            ims = []
            for path in self.local_images:
                with open(path, "rb") as f:
                    im = Image.open(f)
                    im.load()
                while min(*im.size) >= 2 * R:
                    im = im.resize(tuple(x // 2 for x in im.size), resample=Image.BOX)
                scale = R / min(*im.size)
                im = im.resize(tuple(round(x * scale) for x in im.size), resample=Image.BICUBIC)
                arr = np.array(im.convert("RGB"))
                cy, cx = (arr.shape[0] - R) // 2, (arr.shape[1] - R) // 2
                ims.append(arr[cy:cy + R, cx:cx + R])
            images = np.stack(ims) if ims else np.zeros((0, R, R, 3), np.uint8)
            y = None if self.local_classes is None else np.asarray(self.local_classes, dtype=np.int64)
            self._cache = (images, None, y)
        return self._cache


class _Subset(_Resident):
    """index subset of a resident dataset (torch.utils.data.random_split result, ref :320-326)"""

    def __init__(self, base, indices):
        self.base, self.indices = base, np.asarray(indices, dtype=np.int64)
        self.mode = base.mode
        self._n = len(self.indices)

    def host_arrays(self):
        images, c, y = self.base.host_arrays()
        ix = self.indices
        return images[ix], (None if c is None else c[ix]), (None if y is None else y[ix])


# ---------------------------------------------------------------------------------------------- the loader
class ResidentLoader:
    """DataLoader(dataset, batch_size, shuffle, drop_last=True) replacement: the decoded dataset is uploaded to HBM once,
    every batch is one gather kernel launch producing device tensors (no worker processes, no per-batch PCIe copy).
    Iterating yields (images fp32 [B,C,H,W], {"c": fp32 [B,L], "y": int64 [B]}) like the reference's collate."""

    def __init__(self, dataset, batch_size, shuffle=True, drop_last=True, device=None, generator=None):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.generator = generator
        self.device = device
        self._dev = None

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def epoch_order(self):
        """sample order of one epoch (host int64): torch.randperm like RandomSampler, or the identity"""
        n = len(self.dataset)
        return th.randperm(n, generator=self.generator) if self.shuffle else th.arange(n)

    def _upload(self):
        if self._dev is None:
            _lib.lib()                                   # raises without an sm_100 device: no CPU path
            dev = self.device if self.device is not None else th.device("cuda", th.cuda.current_device())
            images, c, y = self.dataset.host_arrays()
            assert images.dtype == np.uint8 and images.ndim == 4
            self._dev = (th.from_numpy(np.ascontiguousarray(images)).to(dev),
                         None if c is None else th.from_numpy(np.ascontiguousarray(c, dtype=np.float32)).to(dev),
                         None if y is None else th.from_numpy(np.ascontiguousarray(y, dtype=np.int64)).to(dev))
        return self._dev

    def batch(self, idx):
        """idx: int64 tensor of dataset rows -> one collated batch on the device"""
        images, c, y = self._upload()
        idx = idx.to(images.device, non_blocking=True).contiguous()
        x, cc = ops.gather_images(images, idx, labels=c, mode=self.dataset.mode)
        out = {}
        if y is not None:
            out["y"] = y.index_select(0, idx)
        if cc is not None:
            out["c"] = cc
        return x, out

    def __iter__(self):
        order = self.epoch_order()
        n, bs = order.numel(), self.batch_size
        stop = n - n % bs if self.drop_last else n
        for i in range(0, stop, bs):
            yield self.batch(order[i:i + bs])


def get_dataloader_morphomnist(path, batch_size, split_set, shard, num_shards):
    """ref :306-341"""
    assert split_set in ["train", "val", "test"]
    dataset = MorphoMNISTLike(root_dir=path, columns=["thickness", "intensity"], train=split_set == "train", shard=shard,
                              num_shards=num_shards)
    if split_set == "val":
        val_ratio = 0.1
        n = len(dataset)
        lengths = [int(n * (1 - val_ratio)), int(n * val_ratio)]
        perm = th.randperm(sum(lengths), generator=th.Generator().manual_seed(42)).tolist()   # = torch random_split
        dataset = _Subset(dataset, perm[lengths[0]:lengths[0] + lengths[1]])
    return ResidentLoader(dataset, batch_size, shuffle=True)


def get_dataloader_pendulum(path, batch_size, split_set, shard, num_shards):
    """ref :394-408"""
    assert split_set in ["train", "val", "test"]
    return ResidentLoader(SyntheticLabeled(path, split=split_set, shard=shard, num_shards=num_shards), batch_size, shuffle=True)


def get_dataloader_circuit(path, batch_size, split_set, shard, num_shards):
    """ref :485-499 (the only loader the reference does not shuffle)"""
    assert split_set in ["train", "val", "test"]
    return ResidentLoader(CausalCircuit(path, split_set, shard=shard, num_shards=num_shards), batch_size, shuffle=False)


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def load_data(*, data_dir, batch_size, image_size, class_cond=False, split="train", deterministic=False):
    """ref :69-126: endless generator over (images, kwargs) batches; the dataset kind is chosen by the directory name."""
    if not data_dir:
        raise ValueError("unspecified data directory")
    rank, world = _rank_world()
    if "celeba" in data_dir:
        all_files = _list_image_files_recursively(data_dir)
        classes = None
        if class_cond:
            class_names = [os.path.basename(path).split("_")[0] for path in all_files]
            sorted_classes = {x: i for i, x in enumerate(sorted(set(class_names)))}
            classes = [sorted_classes[x] for x in class_names]
        loader = ResidentLoader(ImageDataset(image_size, all_files, classes=classes, shard=rank, num_shards=world),
                                batch_size, shuffle=not deterministic)
    elif "morphomnist" in data_dir:
        loader = get_dataloader_morphomnist(data_dir, batch_size, split_set=split, shard=rank, num_shards=world)
    elif "pendulum" in data_dir:
        loader = get_dataloader_pendulum(data_dir, batch_size, split_set=split, shard=rank, num_shards=world)
    elif "circuit" in data_dir:
        loader = get_dataloader_circuit(data_dir, batch_size, split_set=split, shard=rank, num_shards=world)
    else:
        raise ValueError(f"cannot tell the dataset kind from data_dir={data_dir!r} (celeba / morphomnist / pendulum / circuit)")
    while True:
        yield from loader
