"""world_size-2 data-parallel test on CPU (gloo): the N>1 host logic -- sharding, the all-reduce of the flat
gradient + tail scalars, 1/N scaling, identical penalty/clamp/Adam on every rank -- checked with the oracle.
Also covers bench.py's rendezvous plumbing (unique-id broadcast, max-over-ranks)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import parity_utils as PU
    import dp_ref
    from oracle import oracle as O
    O.set_num_threads(2)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, C = 4, 1
    base = PU.make_case(B, C, seed=900)          # identical initial parameters on every rank
    case = PU.make_case(B, C, seed=901 + rank)   # rank-distinct shard
    case["PG"], case["PD"] = base["PG"], base["PD"]
    st = PU.fresh_state(case)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    res = dp_ref.rank_step(case, st, B, C, world, allreduce)
    # bench.py plumbing: rank 0's 128-byte id reaches everyone; max over ranks
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, st["PD"].copy(), st["PG"].copy(), res["gradD"].copy(), res["conf"].copy(), ids[0], float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo_replicas_identical_and_equal_serial():
    import torch.multiprocessing as mp
    import parity_utils as PU
    import dp_ref
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r = q.get(timeout=600)
        got[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas stay bit-identical
    np.testing.assert_array_equal(got[0][1], got[1][1])
    np.testing.assert_array_equal(got[0][2], got[1][2])
    assert got[1][5] == bytes(range(128)) and got[0][6] == 2.0 and got[1][6] == 2.0
    assert got[0][4].sum() == 2 * 4  # confusion counts are global (all-reduced)
    # serial emulation of the two ranks (allreduce = explicit sum over both shards)
    B, C = 4, 1
    base = PU.make_case(B, C, seed=900)
    cases = []
    for r in range(world):
        c = PU.make_case(B, C, seed=901 + r)
        c["PG"], c["PD"] = base["PG"], base["PD"]
        cases.append(c)
    # run rank 0's step with an allreduce that adds rank 1's contribution computed on the fly
    import threading
    states = [PU.fresh_state(c) for c in cases]
    bufs, lock, bar = {}, threading.Lock(), threading.Barrier(world)

    def make_allreduce(rank):
        def ar(a):
            with lock:
                bufs[rank] = np.array(a, np.float64)
            bar.wait()
            tot = bufs[0] + bufs[1]
            bar.wait()
            return tot
        return ar

    out = [None, None]

    def run(rank):
        out[rank] = dp_ref.rank_step(cases[rank], states[rank], B, C, world, make_allreduce(rank))

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert PU.relerr(got[0][3], out[0]["gradD"]) < 1e-12
    assert np.abs(got[0][1] - states[0]["PD"]).max() < 1e-12
    assert np.abs(got[0][2] - states[0]["PG"]).max() < 1e-12
    # and the DP gradient is the mean of the per-shard gradients, not the gradient of the concatenated batch
    # (BatchNorm statistics are per replica) -- documented in DESIGN.md section 4
