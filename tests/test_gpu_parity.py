"""GPU tier: the CUDA build (through the C ABI) against the CPU oracle and the golden traces."""
import numpy as np
import pytest
import torch

import parity as P
import test_kernel_logic_emul as E
from pve_mcc_for_unsignalized_intersection_b200.arrivals import stress_arrivals, synthetic_arrivals

pytestmark = pytest.mark.gpu


def test_backend_is_cuda():
    scene = P.make_scene("cuda", 4)
    assert scene.backend == "cuda-sm_100a"
    assert torch.cuda.get_device_capability(0)[0] == 10


@pytest.mark.parametrize("threads", [64, 96, 128, 256, 512])
def test_free_running_matches_oracle(threads):
    tabs = np.concatenate([synthetic_arrivals(4, lam, 50.0, seed=lam, rows=40) for lam in (400, 1000, 1200)])
    B = tabs.shape[0]
    scene = P.make_scene("cuda", B, vm=5, threads=threads)
    orc = P.make_oracle(B, vm=5, veh_cap=scene.veh_cap)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(threads)
    for t in range(420):
        st = orc.get_state()
        act = P.random_actions(rng, (st["flags"] & 1) != 0)
        o_ref = orc.step(act)
        o_dev = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        P.compare_outputs(o_dev, o_ref, "threads=%d tick %d" % (threads, t))
        if t % 20 == 0 or t == 419:
            P.compare_states(scene.get_state(), orc.get_state(), "tick %d" % t)


def test_train_setting_vm6_accel():
    E.free_run("cuda", synthetic_arrivals(3, 1000, 40.0, seed=77, rows=32), vm=6, ticks=330, seed=2, policy="accel")


@pytest.mark.parametrize("threads", [0, 128, 256])          # 0 = the class default (512 threads for the large classes)
def test_stress_occupancy_brake(threads):
    E.free_run("cuda", stress_arrivals(2, 40.0), vm=5, ticks=300, seed=3, veh_cap=384, agent_cap=320, policy="brake",
               threads=threads)


@pytest.mark.parametrize("name", E.ROLLOUTS)
def test_golden_rollout_direct(name, monkeypatch):
    monkeypatch.setattr(E, "BACKEND", "cuda")
    E.test_golden_rollout_direct(name)


def test_crafted_order_dependence_cases(monkeypatch):
    monkeypatch.setattr(E, "BACKEND", "cuda")
    E.test_crafted_order_dependence_cases()


def test_teacher_forced_every_tick():
    """North-star protocol: every tick starts from the oracle's state copied onto the device."""
    tabs = synthetic_arrivals(8, 1000, 45.0, seed=5, rows=40)
    B = tabs.shape[0]
    scene = P.make_scene("cuda", B, vm=5)
    orc = P.make_oracle(B, vm=5, veh_cap=scene.veh_cap)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(9)
    for t in range(400):
        st = orc.get_state()
        scene.set_state(P.oracle_state_for_device(st))
        act = P.random_actions(rng, (st["flags"] & 1) != 0)
        o_ref = orc.step(act)
        o_dev = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        P.compare_outputs(o_dev, o_ref, "teacher-forced tick %d" % t)
        P.compare_states(scene.get_state(), orc.get_state(), "teacher-forced tick %d" % t)


def test_full_size_4096_intersections_vs_oracle():
    """BASELINE config 2 size: 4,096 intersections, density 1000, random actions; every 50th tick
    compared row by row with the oracle running the same batch on the host cores."""
    B = 4096
    tabs = synthetic_arrivals(B, 1000, 32.0, seed=11, rows=24)
    scene = P.make_scene("cuda", B, vm=5)
    orc = P.make_oracle(B, vm=5, veh_cap=scene.veh_cap, n_threads=32)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(4)
    total = 0
    for t in range(260):
        st_ctrl = scene.control_mask().cpu().numpy()
        act = P.random_actions(rng, st_ctrl)
        o_ref = orc.step(act)
        out = scene.step(P.to_device_actions(scene, act))
        total += out.n_agents
        assert out.n_agents == len(o_ref["reward"]), t
        if t % 50 == 0 or t == 259:
            P.compare_outputs(P.outputs_to_numpy(out), o_ref, "B=4096 tick %d" % t)
    P.compare_states(scene.get_state(), orc.get_state(), "B=4096 final")
    s = scene.stats()
    assert s["agent_steps"] == total and s["overflow"] == 0


def test_step_host_end_to_end_matches_device_outputs(monkeypatch):
    monkeypatch.setenv("PVE_HOST_ZEROCOPY", "0")          # staged copies: host and device views both written
    B = 64
    tabs = synthetic_arrivals(B, 1000, 30.0, seed=3, rows=24)
    scene = P.make_scene("cuda", B)
    scene.reset(tabs, warmup=True)
    host = scene.make_host_outputs()
    act = torch.zeros(B, scene.veh_cap, dtype=torch.float32).pin_memory()
    rng = np.random.RandomState(0)
    for t in range(150):
        act.copy_(torch.from_numpy(rng.uniform(-3, 3, size=(B, scene.veh_cap)).astype(np.float32)))
        n = scene.step_host(act, host, copy_obs=True)
        assert n == int(host.agent_offset[-1])
        for f in ("reward", "cpv", "status", "jerk_sum", "ids", "obs"):
            assert torch.equal(getattr(host, f)[:n], getattr(scene.out, f)[:n].cpu()), (f, t)
    assert n > 0


def test_capacity_overflow_is_flagged_not_silent():
    tabs = stress_arrivals(1, 30.0)
    scene = P.make_scene("cuda", 1, veh_cap=64, agent_cap=48)
    scene.reset(tabs, warmup=True)
    act = torch.full((1, scene.veh_cap), -3.0, device="cuda")
    for _ in range(300):
        scene.step(act)
    st = scene.get_state()
    assert st["overflow"][0] > 0 and st["n_veh"][0] <= scene.veh_cap and st["n_ctrl"][0] <= scene.agent_cap
    assert scene.stats()["overflow"] > 0


def test_config3_teacher_forced_sample_of_a_large_shard():
    """BASELINE config 3 at reduced size (the full 8 x 8,192 run is tests/config3_check.py under torchrun):
    a shard of 2,048 intersections free-runs on the GPU; a strided sample is teacher-forced against the oracle."""
    import config3_check
    res, _ = config3_check.run_check(2048, ticks=120, sample=32, check_every=20)
    assert res["checks"] == 7 and res["agent_rows_compared"] > 3000


def host_outputs_to_numpy(host, n):
    g = lambda t: t[:n].numpy() if t.shape[0] != host.agent_offset.shape[0] - 1 else t.numpy()
    return {"agent_offset": host.agent_offset.numpy(), "obs": host.obs[:n].numpy(), "reward": host.reward[:n].numpy(),
            "ids": host.ids[:n].numpy(), "cpv": host.cpv[:n].numpy(), "status": host.status[:n].numpy(),
            "jerk_sum": host.jerk_sum[:n].numpy(), "collisions": host.env_collisions.numpy(),
            "lock": host.env_lock.numpy(), "n_removed": host.env_removed.numpy()}


@pytest.mark.parametrize("zerocopy", ["1", "0"])
def test_step_host_zero_copy_and_staged_paths_at_scale(zerocopy, monkeypatch):
    """pve_step_host with pinned host buffers lets the kernel read the actions and write the small outputs in
    place over PCIe (PVE_HOST_ZEROCOPY=1, default); =0 is the staged-copy path.  Either way the HOST buffers
    must follow the oracle tick by tick; the device-resident path may be mixed in."""
    monkeypatch.setenv("PVE_HOST_ZEROCOPY", zerocopy)
    B = 2304
    tabs = synthetic_arrivals(B, 1000, 30.0, seed=8, rows=24)
    scene = P.make_scene("cuda", B, vm=5)
    orc = P.make_oracle(B, vm=5, veh_cap=scene.veh_cap, n_threads=32)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    host = scene.make_host_outputs()
    act = torch.zeros(B, scene.veh_cap, dtype=torch.float32).pin_memory()
    rng = np.random.RandomState(1)
    for t in range(90):
        a = P.random_actions(rng, scene.control_mask().cpu().numpy())
        act.copy_(torch.from_numpy(a))
        o_ref = orc.step(a)
        n = scene.step_host(act, host, copy_obs=True)
        assert n == int(host.agent_offset[-1]) == len(o_ref["reward"])
        assert torch.equal(host.obs[:n], scene.out.obs[:n].cpu())          # observations always live on the device too
        if t % 10 == 0 or t > 85:
            P.compare_outputs(host_outputs_to_numpy(host, n), o_ref, "host step tick %d" % t)
        if t == 45:                             # the device-resident path in between must not disturb it
            a = P.random_actions(rng, scene.control_mask().cpu().numpy())
            o_ref = orc.step(a)
            P.compare_outputs(P.outputs_to_numpy(scene.step(P.to_device_actions(scene, a))), o_ref, "mixed")
    # pageable host actions: staged copy of the actions, same results
    a = P.random_actions(rng, scene.control_mask().cpu().numpy())
    o_ref = orc.step(a)
    n = scene.step_host(torch.from_numpy(a), host, copy_obs=True)
    P.compare_outputs(host_outputs_to_numpy(host, n), o_ref, "pageable actions")
    # pageable outputs: staged copies
    host2 = scene.make_host_outputs(pinned=False)
    a = P.random_actions(rng, scene.control_mask().cpu().numpy())
    o_ref = orc.step(a)
    act.copy_(torch.from_numpy(a))
    n = scene.step_host(act, host2, copy_obs=True)
    P.compare_outputs(host_outputs_to_numpy(host2, n), o_ref, "pageable outputs")
    P.compare_states(scene.get_state(), orc.get_state(), "host path final")


def test_neighbour_sources_name_the_copied_rows():
    E.check_neighbour_sources("cuda")


def test_ragged_and_empty_intersections():
    scene, n = E.free_run("cuda", E.edge_tables(), vm=5, ticks=320, seed=6)
    st = scene.get_state()
    assert st["id_seq"][1] == 36 and st["id_seq"][4] == 1 and st["tick"][3] > 30000 and n > 5000


def test_largest_capacity_class():
    scene, n = E.free_run("cuda", stress_arrivals(1, 60.0, headway=0.7), vm=5, ticks=420, seed=3, veh_cap=576,
                          agent_cap=416, policy="brake")
    st = scene.get_state()
    assert st["n_veh"][0] > 400 and st["n_ctrl"][0] > 384 and st["overflow"][0] == 0 and n > 100000
