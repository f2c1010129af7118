"""N > 1 path on CPU: two gloo ranks each own a contiguous block of intersections (kernel-logic
emulation, no step-path communication) and all-reduce the end-of-rollout statistics; the result
must equal one process running the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B_TOTAL, TICKS = 6, 150


def rollout(lo, hi):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity as P
    from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
    tabs = synthetic_arrivals(B_TOTAL, 1000, 30.0, seed=9, rows=24)[lo:hi]
    scene = P.make_scene("emul", hi - lo, vm=6)
    scene.reset(tabs, warmup=True)
    rng = np.random.RandomState(17)
    acts = rng.uniform(-3, 3, size=(TICKS, B_TOTAL, scene.veh_cap)).astype(np.float32)[:, lo:hi]
    for t in range(TICKS):
        scene.step(torch.from_numpy(np.ascontiguousarray(acts[t])))
    return scene.stats_tensor().clone()


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from pve_mcc_for_unsignalized_intersection_b200.distributed import max_over_ranks, reduce_stats, shard_range
    lo, hi = shard_range(B_TOTAL, rank, world)
    total = reduce_stats(rollout(lo, hi))
    slow = max_over_ranks(1.0 + rank, "cpu")
    if rank == 0:
        q.put((total.numpy(), slow))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, 29617, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, slow = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = rollout(0, B_TOTAL).numpy()
    np.testing.assert_array_equal(total[:5], single[:5])          # counts are exact
    np.testing.assert_allclose(total, single, rtol=1e-12)
    assert total[0] > 10000 and slow == 2.0
