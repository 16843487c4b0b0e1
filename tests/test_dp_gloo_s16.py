"""world_size-2 data-parallel test on CPU (gloo) for the --scale 16 loop (fg_s16_train_step with world > 1): sharding,
the all-reduce of the flat gradient + confusion counts, 1/N, identical optimizer steps -- checked with the oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _cases(world, B, C):
    import s16_utils as SU
    base = SU.make_case(B, C, seed=950, init="trained")
    out = []
    for r in range(world):
        c = SU.make_case(B, C, seed=951 + r, init="trained")
        c["PG"], c["PD"] = base["PG"], base["PD"]
        out.append(c)
    return out


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import dp_ref_s16
    from oracle import oracle as O
    O.set_num_threads(2)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, C = 4, 1
    case = _cases(world, B, C)[rank]
    st = dp_ref_s16.fresh_state(case)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    res = dp_ref_s16.rank_step(case, st, B, C, world, allreduce)
    q.put((rank, st["PD"].copy(), st["PG"].copy(), res["gradD"].copy(), res["gradG"].copy(), res["conf"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo_s16_replicas_identical_and_mean_of_shard_gradients():
    import torch.multiprocessing as mp
    import s16_utils as SU
    from oracle import oracle as O
    from oracle import oracle_s16 as OS
    world, port, B, C = 2, 29735, 4, 1
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
    np.testing.assert_array_equal(got[0][1], got[1][1])  # replicas stay bit-identical
    np.testing.assert_array_equal(got[0][2], got[1][2])
    assert got[0][5].sum() == world * B                  # confusion counts are global
    # the all-reduced D gradient is the mean of the two shards' own (pre-penalty) gradients
    cases = _cases(world, B, C)
    raw = []
    for c in cases:
        g, d = OS.f64.G(), OS.f64.D()
        fake = g.forward(c["PG"].astype(np.float64), c["noise_D"], C, SU.bn_init())
        x = np.concatenate([c["real"].astype(np.float64), fake])
        tg = np.concatenate([np.ones(B // 2), np.zeros(B // 2)])
        out = d.forward(c["PD"].astype(np.float64), x, c["masks_D"], True)
        raw.append(d.backward(O.f64.bce_bwd(out, tg))[0])
    mean = (raw[0] + raw[1]) / world
    O.f64.penalty_clamp(cases[0]["PD"].astype(np.float64), mean, SU.HYPER["D_L1"], SU.HYPER["D_L1"], SU.HYPER["D_L2"],
                        SU.HYPER["D_clamp"])
    assert np.abs(got[0][3] - mean).max() < 1e-12 * max(1.0, np.abs(mean).max())
