"""GPU tier, row N3: the 4-lane intersection (``lane_num=4``) on the CUDA kernel (csrc/scene_step4.cuh, one warp per
intersection) through the C ABI: the reference's own rollouts, free-running batches against the 4-lane oracle, the
pipelined host path, and a full-size batch with size-independent invariants."""
import numpy as np
import pytest
import torch

import parity as P
import test_kernel4_logic_emul as E
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", E.ROLLOUTS4)
def test_golden_rollout4_direct(name):
    scene = E.run_golden("cuda", name)
    assert scene.backend == "cuda-sm_100a" and scene.launch_info["dual"] is False


@pytest.mark.parametrize("name", E.ROLLOUTS4)
def test_golden_rollout4_direct_in_the_64_class(name, monkeypatch):
    """The small capacity class of this path (64 vehicle slots, lists of 64 entries: half the shared memory per warp)
    against the same rollouts of the reference."""
    monkeypatch.setattr(E, "CAPS4", (64, 64))
    scene = E.run_golden("cuda", name)
    assert scene.veh_cap == 64 and scene.backend == "cuda-sm_100a"


def test_free_running_lane4_matches_oracle_in_the_64_class(monkeypatch):
    monkeypatch.setattr(E, "CAPS4", (64, 64))
    assert E.free_run4("cuda", 6, 1400, 300, seed=5) > 9000


def test_free_running_lane4_matches_oracle():
    assert E.free_run4("cuda", 6, 1400, 300, seed=5) > 9000
    assert E.free_run4("cuda", 4, 600, 300, seed=8, vm=6) > 1500


@pytest.mark.parametrize("name", E.ROLLOUTS8)
def test_golden_rollout8_direct(name):
    scene = E.run_golden("cuda", name, lanes=8)
    assert scene.backend == "cuda-sm_100a" and scene.launch_info["dual"] is False


def test_free_running_lane8_matches_oracle():
    assert E.free_run4("cuda", 6, 1000, 300, seed=6, lanes=8) > 9000
    assert E.free_run4("cuda", 4, 500, 300, seed=9, vm=6, lanes=8) > 1500


def test_lane4_batch_of_4096_invariants_and_sampled_oracle():
    """4,096 4-lane intersections: a strided sample is compared with the oracle row by row every tick, the whole batch
    through invariants that do not depend on its size (dense offsets, vehicle conservation, counters)."""
    from oracle.scene4_oracle import Scene4Oracle
    B, ticks = 4096, 200
    tabs = synthetic_arrivals(B, 1200, ticks * 0.1 + 30.0, seed=3)[:, :, :4].copy()
    scene = E.make_scene4("cuda", B)
    scene.reset(tabs, warmup=True)
    sample = list(range(0, B, 512))
    orcs = {b: Scene4Oracle() for b in sample}
    for b, o in orcs.items():
        o.reset(tabs[b], warmup=True)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    rows = 0
    for t in range(ticks):
        act = ((torch.rand(B, scene.veh_cap, device="cuda", generator=gen) * 6 - 3) * scene.control_mask()).contiguous()
        a_host = act[sample].cpu().numpy()
        out = scene.step(act)
        off = out.agent_offset.cpu().numpy()
        assert off[0] == 0 and np.all(np.diff(off) >= 0)
        rows += int(off[-1])
        ids = out.ids[:int(off[-1])].cpu().numpy()
        assert np.array_equal(ids[:, 0], np.repeat(np.arange(B), np.diff(off)))
        for n, b in enumerate(sample):
            o = orcs[b]
            V = sum(len(x) for x in o.lanes)
            ref = o.step(a_host[n, :V])
            lo, hi = int(off[b]), int(off[b + 1])
            assert hi - lo == len(ref["ids"]), (b, t)
            np.testing.assert_array_equal(ids[lo:hi, 1:3], np.array(ref["ids"], np.int32).reshape(-1, 2))
            P.assert_rel(out.reward[lo:hi].cpu().numpy(), np.array(ref["reward"], np.float64), "env %d tick %d reward" % (b, t))
            P.assert_rel(out.obs[lo:hi].cpu().numpy(), np.array(ref["obs"], np.float64).reshape(-1, 7, 28), "env %d tick %d obs" % (b, t))
    st = scene.get_state()
    s = scene.stats()
    assert s["agent_steps"] == rows and s["overflow"] == 0 and s["env_steps"] == B * (ticks + 1)
    assert int(st["id_seq"].sum()) == s["spawned"] and int(st["n_veh"].sum()) == s["spawned"] - s["removed"]
    for b, o in orcs.items():
        E.check_state(st, o.snapshot(), b, "env %d final" % b)


def test_lane4_pipelined_host_path():
    B = 64
    tabs = synthetic_arrivals(B, 1000, 40.0, seed=9)[:, :, :4].copy()
    twin, scene = E.make_scene4("cuda", B), E.make_scene4("cuda", B)
    twin.reset(tabs, warmup=True)
    scene.reset(tabs, warmup=True)
    pairs = scene.make_async_buffers()
    rng = np.random.RandomState(2)
    acts = [torch.from_numpy(rng.uniform(-3, 3, size=(B, scene.veh_cap)).astype(np.float32)).pin_memory() for _ in range(4)]
    want = []
    for t in range(120):
        want.append(P.outputs_to_numpy(twin.step(acts[t % 4].cuda())))
        scene.step_host_async(acts[t % 4], pairs[t % 3], copy_obs=True)
        if t >= 2:
            n, host = scene.host_wait()
            w = want.pop(0)
            rec = host.records(n)
            assert n == len(w["reward"])
            np.testing.assert_array_equal(rec["reward"], w["reward"])
            np.testing.assert_array_equal(rec["status"], w["status"])
            np.testing.assert_array_equal(host.obs[:n].numpy(), w["obs"])
    scene.host_wait(); scene.host_wait()
