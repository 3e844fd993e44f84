"""Import the *real* reference (read-only tree) in the build container to pin the oracle.

Test infrastructure only.  Used by tests/golden/make_golden.py (fixture generation) and by
`bench.py --impl reference` (the reference arm: the unmodified reference timed on the host cores).
Nothing here is copied from the reference: it is imported in place with
  * two stand-in modules for imports the container lacks (blobfile, mpi4py)  [SURVEY.md App. B]
  * the four documented oracle patches of SURVEY.md 8c:
      1. encoder depth follows image size (Q1)            -> model.rep_emb replaced after construction
      2. injectable DAG adjacency (Q2)                    -> `th.tensor` literal intercepted in unet.forward
      3. guidance zeros of width rep_dim, not 64 (Q4)     -> `th.zeros((B, 64))` intercepted in p_mean_variance
      4. seeded de-zeroing of zero_module tensors (Q5)    -> done by loading oracle.model.seeded_state_dict
"""
import os
import sys
import types

def _find_root():
    """the reference tree in the build container, else the unmodified copy oracle/build_ref.py placed under oracle/_ref
    (git-ignored; it travels to the GPU box, where /root/reference does not exist)"""
    cands = [os.environ.get("CDAE_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "improved_diffusion")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "improved_diffusion"))


class _TorchProxy:
    """Delegates to torch; rewrites two literals the reference hard-codes (patches 2 and 3)."""

    def __init__(self, torch_mod):
        self._t = torch_mod
        self.inject_A = None
        self.rep_dim = None

    def __getattr__(self, k):
        return getattr(self._t, k)

    def tensor(self, data, *a, **kw):
        if self.inject_A is not None and isinstance(data, list) and data and isinstance(data[0], list) \
                and len(data) == len(self.inject_A) and len(data) in (2, 4):
            data = self.inject_A
        return self._t.tensor(data, *a, **kw)

    def zeros(self, *size, **kw):
        if self.rep_dim is not None and len(size) == 1 and isinstance(size[0], tuple) and len(size[0]) == 2 \
                and size[0][1] == 64:
            size = ((size[0][0], self.rep_dim),)
        return self._t.zeros(*size, **kw)


_loaded = {}


def load():
    """Returns a namespace with the reference modules (nn, unet, gaussian_diffusion, respace, resample,
    script_util, train_util, dist_util, logger) and the proxies used for patches 2/3."""
    if _loaded:
        return _loaded["ns"]
    import torch
    sys.dont_write_bytecode = True
    if "blobfile" not in sys.modules:
        bf = types.ModuleType("blobfile")
        bf.BlobFile = lambda p, m="rb": open(p, m)
        bf.join, bf.dirname, bf.exists = os.path.join, os.path.dirname, os.path.exists
        sys.modules["blobfile"] = bf
    if "mpi4py" not in sys.modules:
        class _Comm:
            rank, size = 0, 1
            def Get_rank(self): return 0
            def Get_size(self): return 1
            def bcast(self, x, root=0): return x
        mpi = types.ModuleType("mpi4py")
        mpi.MPI = types.SimpleNamespace(COMM_WORLD=_Comm())
        sys.modules["mpi4py"] = mpi
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            tv = types.ModuleType("torchvision"); tvu = types.ModuleType("torchvision.utils")
            tvu.save_image = lambda *a, **k: None
            tv.utils = tvu
            sys.modules["torchvision"], sys.modules["torchvision.utils"] = tv, tvu
    sys.path.insert(0, REF_ROOT)
    import improved_diffusion.nn as rnn
    import improved_diffusion.unet as runet
    import improved_diffusion.gaussian_diffusion as rgd
    import improved_diffusion.respace as rrespace
    import improved_diffusion.resample as rresample
    import improved_diffusion.script_util as rsu
    import improved_diffusion.dist_util as rdist
    import improved_diffusion.logger as rlogger
    import improved_diffusion.train_util as rtrain
    sys.path.pop(0)
    unet_proxy, gd_proxy = _TorchProxy(torch), _TorchProxy(torch)
    runet.th, rgd.th = unet_proxy, gd_proxy
    ns = types.SimpleNamespace(nn=rnn, unet=runet, gd=rgd, respace=rrespace, resample=rresample, su=rsu,
                               dist=rdist, logger=rlogger, train=rtrain, unet_proxy=unet_proxy, gd_proxy=gd_proxy)
    _loaded["ns"] = ns
    return ns


def build(flags, rep_dim=512, A=None):
    """create_model_and_diffusion(**flags) on the real reference + patches 1-3. Returns (model, diffusion)."""
    from . import schedules
    ns = load()
    ns.su.REP_DIM = rep_dim
    full = {**ns.su.model_and_diffusion_defaults(), **flags}
    model, diff = ns.su.create_model_and_diffusion(**full)
    if full["rep_cond"]:
        dims = schedules.encoder_hidden_dims(full["image_size"], full["n_vars"])
        model.rep_emb = ns.nn.GaussianConvEncoder(full["in_channels"], latent_dim=rep_dim, hidden_dims=dims,
                                                  num_vars=full["n_vars"])
    ns.unet_proxy.inject_A = A
    ns.gd_proxy.rep_dim = rep_dim
    return model, diff
