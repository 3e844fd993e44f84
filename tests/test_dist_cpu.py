"""world_size-2 gloo tests (CPU) of the host-side data-parallel plumbing: bootstrap, rank-0 checkpoint broadcast,
parameter sync, gradient averaging convention and batch sharding."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from causaldiffae_b200 import dist_util
    from causaldiffae_b200.sampling_shard import shard_range
    from causaldiffae_b200.train_util import dp_all_reduce_
    dist_util.setup_dist(backend="gloo")
    assert dist.get_world_size() == world and dist.get_rank() == rank
    # checkpoint bytes are read once (rank 0) and broadcast
    path = os.path.join(tmp, "model000007.pt")
    if rank == 0:
        torch.save({"w": torch.arange(5.0)}, path)
    dist.barrier()
    sd = dist_util.load_state_dict(path, map_location="cpu")
    assert torch.equal(sd["w"], torch.arange(5.0))
    # sync_params really broadcasts rank 0's values (the reference's is a no-op)
    p = torch.full((4,), float(rank + 1))
    dist_util.sync_params([p])
    assert torch.equal(p, torch.ones(4))
    # flat gradient all-reduce: SUM on the wire, 1/world folded into the optimizer's grad_scale (DDP mean semantics)
    g = torch.full((8,), float(rank + 1))
    scale = dp_all_reduce_(g)
    assert scale == 1.0 / world and torch.equal(g * scale, torch.full((8,), 1.5))
    # data path: load_data's dataset shard of this rank is the reference's rank-strided slice [rank:][::world]
    from causaldiffae_b200 import image_datasets as ids
    from tests.golden import dataset_fixture as fx
    mm = os.path.join(tmp, "morphomnist")
    if rank == 0:
        fx.make_morphomnist(mm)
    dist.barrier()
    assert ids._rank_world() == (rank, world)
    mine = ids.get_dataloader_morphomnist(mm, 2, "train", *ids._rank_world()).dataset
    full = ids.MorphoMNISTLike(mm, columns=["thickness", "intensity"], train=True)
    assert np.array_equal(mine.host_arrays()[0], full.host_arrays()[0][rank::world])
    assert np.array_equal(mine.host_arrays()[1], full.host_arrays()[1][rank::world])
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(mine)]))
    assert sum(int(c) for c in counts) == len(full)                  # the shards partition the dataset
    # loss-aware timestep resampling: ragged per-rank batches are gathered once, every rank ends with the same history
    from causaldiffae_b200.resample import LossSecondMomentResampler
    from types import SimpleNamespace
    smp = LossSecondMomentResampler(SimpleNamespace(num_timesteps=6), history_per_term=2)
    ts = [torch.tensor([0, 1, 2]), torch.tensor([3, 4, 5, 0, 1])][rank]
    smp.update_with_local_losses(ts, ts.float() + 10.0 * (rank + 1))
    expect = LossSecondMomentResampler(SimpleNamespace(num_timesteps=6), history_per_term=2)
    expect.update_with_all_losses([0, 1, 2, 3, 4, 5, 0, 1], [10.0, 11.0, 12.0, 23.0, 24.0, 25.0, 20.0, 21.0])
    assert np.array_equal(smp._loss_history, expect._loss_history) and np.array_equal(smp._loss_counts, expect._loss_counts)
    # batch sharding of an intervention sweep: disjoint, ordered, covering
    lo, hi = shard_range(4097, rank, world)
    q.put((rank, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
    got = sorted(q.get() for _ in range(2))
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == 4097
    assert abs((got[0][2] - got[0][1]) - (got[1][2] - got[1][1])) <= 1
