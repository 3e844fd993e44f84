"""Process-group bootstrap (ref improved_diffusion/dist_util.py): same four entry points, but rank/size come from the
torchrun environment instead of MPI, the backend is NCCL over NVLink on GPUs (the reference hard-codes gloo, Q8), and
rank 0 reads checkpoints."""
import io
import os
import socket

import torch as th
import torch.distributed as dist

GPUS_PER_NODE = 8


def setup_dist(backend=None):
    """ref dist_util.py:21-41. Single-process runs create a world of size 1 so that dist.get_world_size() works."""
    if dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if th.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    if "MASTER_PORT" not in os.environ:
        os.environ["MASTER_PORT"] = str(_find_free_port())
    if th.cuda.is_available():
        th.cuda.set_device(dev())
    dist.init_process_group(backend=backend, init_method="env://")


def dev():
    """ref dist_util.py:44-51"""
    if th.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        return th.device(f"cuda:{local % max(1, min(GPUS_PER_NODE, th.cuda.device_count()))}")
    return th.device("cpu")


def load_state_dict(path, **kwargs):
    """ref dist_util.py:54-64: one reader (rank 0), bytes broadcast to the other ranks."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    data = None
    if rank == 0:
        with open(path, "rb") as f:
            data = f.read()
    if world > 1:
        box = [data]
        dist.broadcast_object_list(box, src=0)
        data = box[0]
    return th.load(io.BytesIO(data), **kwargs)


def sync_params(params):
    """ref dist_util.py:67-74 (a no-op there; here rank 0's values are really broadcast)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for p in params:
        with th.no_grad():
            dist.broadcast(p, 0)


def _find_free_port():
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    try:
        s.bind(("", 0))
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        return s.getsockname()[1]
    finally:
        s.close()
