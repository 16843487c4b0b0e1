"""world_size-2 data-parallel test of the coarse-to-fine loop on CPU (gloo): the host-side semantics the CUDA path
implements for world > 1 (tests/test_gpu_dp.py::test_dp_c2f_two_gpus_average_gradients is the GPU counterpart)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import c2f_utils as CU
    import dp_ref_c2f
    from oracle import oracle as O
    O.set_num_threads(2)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, C = 4, 1
    base = CU.make_case(B, C, seed=950)
    case = CU.make_case(B, C, seed=951 + rank)
    case["PG"], case["PD"] = base["PG"], base["PD"]
    st = CU.fresh_state(case)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    res = dp_ref_c2f.rank_step(case, st, B, C, world, allreduce)
    # the same shard alone (world = 1): its gradient enters the mean
    st1 = CU.fresh_state(case)
    single = dp_ref_c2f.rank_step(case, st1, B, C, 1, lambda a: np.array(a, np.float64), dict(CU.HYPER, D_L1=0.0, D_clamp=0.0, G_clamp=0.0))
    q.put((rank, st["PD"].copy(), st["PG"].copy(), res["gradD"].copy(), res["conf"].copy(), single["gradD"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_c2f_world2_gloo():
    import torch.multiprocessing as mp
    import c2f_utils as CU
    import dp_ref_c2f
    world, port = 2, 29771
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r = q.get(timeout=900)
        got[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got[0][1], got[1][1])  # replicas stay bit-identical
    np.testing.assert_array_equal(got[0][2], got[1][2])
    assert got[0][4].sum() == 2 * 4                      # global confusion counts
    # with D_L1 = 1e-7 and clamp 1 active the reduced gradient is clamp(mean(shard grads) + penalty): check the mean part
    # on the entries the clamp leaves alone
    mean = 0.5 * (got[0][5] + got[1][5])
    base = CU.make_case(4, 1, seed=950)
    pen = 1e-7 * np.sign(base["PD"].astype(np.float64))
    free = np.abs(mean + pen) < 0.99
    assert free.mean() > 0.9
    assert np.abs(got[0][3][free] - (mean + pen)[free]).max() < 1e-12
