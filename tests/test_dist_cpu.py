"""world_size-2 gloo test of the multi-GPU plumbing (sharding, per-rank synthetic inputs, MAX-reduced timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hoig_b200 import dist_utils, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = dist_utils.shard_range(8, rank, world)
    inp = synth.generator_inputs(hi - lo, seed=dist_utils.shard_seed(100, rank), size=16)
    sig = float(inp["bg_inputs"].double().sum())
    ms, ms2 = dist_utils.reduce_max([10.0 + 5.0 * rank, 7.0 - rank], "cpu")
    dist.barrier()
    q.put((rank, lo, hi, sig, ms, ms2, dist_utils.env_rank()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, ms0, m20, e0), (r1, lo1, hi1, s1, ms1, m21, e1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 8)            # contiguous, disjoint, covering
    assert s0 != s1                                        # different synthetic shards per rank
    assert ms0 == ms1 == 15.0 and m20 == m21 == 7.0        # MAX over ranks, same on every rank
    assert e0 == (0, 0, 2) and e1 == (1, 1, 2)
    assert dist_utils.throughput(64, 8, 10, 1000.0) == 5120.0


def test_shard_range_rejects_uneven():
    import pytest
    with pytest.raises(ValueError):
        dist_utils.shard_range(10, 0, 4)
    assert dist_utils.reduce_max([3.0], "cpu") == [3.0]    # not initialised: identity


def _allreduce_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from hoig_b200.training import allreduce_gradients
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((5, 3), (7,), (2, 2, 2), (11,))]
    for i, p in enumerate(params[:3]):          # the last parameter has no gradient (find_unused_parameters case)
        p.grad = torch.full(p.shape, float(rank + 1) * (i + 1))
    nbytes = allreduce_gradients(params, bucket_bytes=40)      # tiny buckets: several flushes
    q.put((rank, nbytes, [None if p.grad is None else p.grad.clone() for p in params]))
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo():
    """The DDP role of the training step (models/trainer.py:237-252): bucketed all-reduce that AVERAGES gradients over ranks."""
    import socket
    import torch
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, grads in res:
        assert nbytes == (15 + 7 + 8) * 4
        for i, g in enumerate(grads[:3]):
            assert torch.allclose(g, torch.full_like(g, 1.5 * (i + 1)))      # mean of (1, 2) * (i + 1)
        assert grads[3] is None
