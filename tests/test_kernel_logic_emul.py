"""CPU tier: the CUDA kernel SOURCE, compiled as a sequential emulation, against the oracle.

This checks the order-free reformulation in
pve_mcc_for_unsignalized_intersection_b200/csrc/scene_step.cuh without a GPU.  The same checks run
on the real CUDA build in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

import parity as P
from golden_io import ROLLOUTS, STATE_KEYS_E, load_crafted, load_rollout, snapshot_to_state
from pve_mcc_for_unsignalized_intersection_b200.arrivals import stress_arrivals, synthetic_arrivals

BACKEND = "emul"


def free_run(backend, tables, vm, ticks, seed, veh_cap=128, agent_cap=96, policy="uniform", threads=0):
    B = tables.shape[0]
    scene = P.make_scene(backend, B, vm=vm, veh_cap=veh_cap, agent_cap=agent_cap, threads=threads)
    orc = P.make_oracle(B, vm=vm, veh_cap=scene.veh_cap)
    scene.reset(tables, warmup=True)
    orc.reset(tables, warmup=True)
    P.compare_states(scene.get_state(), orc.get_state(), "after reset")
    rng = np.random.RandomState(seed)
    agent_steps = 0
    for t in range(ticks):
        st = orc.get_state()
        ctrl = (st["flags"] & 1) != 0
        if policy == "uniform":
            act = P.random_actions(rng, ctrl)
        elif policy == "brake":
            act = np.where(ctrl, -3.0, 0.0).astype(np.float32)
        else:
            act = np.where(ctrl, 3.0, 0.0).astype(np.float32)
        o_ref = orc.step(act)
        o_dev = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        P.compare_outputs(o_dev, o_ref, "tick %d" % t)
        assert o_ref["overflow"] == 0 and o_ref["q5_undefined"].sum() == 0
        agent_steps += len(o_ref["reward"])
        if t % 10 == 0 or t == ticks - 1:
            P.compare_states(scene.get_state(), orc.get_state(), "tick %d" % t)
    return scene, agent_steps


def test_free_running_matches_oracle_mixed_densities():
    tabs = np.concatenate([synthetic_arrivals(2, lam, 50.0, seed=lam, rows=40) for lam in (400, 1000, 1200)])
    scene, n = free_run(BACKEND, tabs, vm=5, ticks=420, seed=1)
    assert n > 20000
    s = scene.stats()
    assert s["agent_steps"] == n and s["overflow"] == 0


def test_free_running_train_setting_vm6_accel():
    tabs = synthetic_arrivals(3, 1000, 40.0, seed=77, rows=32)
    free_run(BACKEND, tabs, vm=6, ticks=330, seed=2, policy="accel")


def test_zero_uncontrolled_actions_option():
    """pve_config.zero_uncontrolled: garbage in the slots of uncontrolled vehicles == the reference driver's 0 (MAIN:401-405)."""
    tabs = synthetic_arrivals(2, 1000, 40.0, seed=5, rows=32)
    scene = P.make_scene(BACKEND, 2, vm=5, zero_uncontrolled_actions=True)
    orc = P.make_oracle(2, vm=5, veh_cap=scene.veh_cap)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(11)
    n_unctl = 0
    for t in range(300):
        ctrl = (orc.get_state()["flags"] & 1) != 0
        raw = rng.uniform(-3, 3, size=ctrl.shape).astype(np.float32)          # every slot filled
        o_ref = orc.step(np.where(ctrl, raw, 0).astype(np.float32))
        o_dev = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, raw)))
        P.compare_outputs(o_dev, o_ref, "tick %d" % t)
        n_unctl += int((~ctrl & (np.arange(ctrl.shape[1])[None, :] < orc.get_state()["lane_n"].sum(axis=1)[:, None])).sum())
    P.compare_states(scene.get_state(), orc.get_state(), "final")
    assert n_unctl > 1000


def test_stress_occupancy_brake():
    tabs = stress_arrivals(1, 40.0)
    free_run(BACKEND, tabs, vm=5, ticks=300, seed=3, veh_cap=384, agent_cap=320, policy="brake")


@pytest.mark.parametrize("name", ROLLOUTS)
def test_golden_rollout_direct(name):
    """Kernel logic against the reference's own trace (no oracle in between)."""
    z, r = load_rollout(name)
    big = name == "stress_brake"
    scene = P.make_scene(BACKEND, 1, vm=float(z["vm"]), collision_thr=float(z["collision_thr"]), veh_cap=384 if big else 128, agent_cap=320 if big else 96)
    scene.reset(z["table"], warmup=True)
    obs_at = {int(t): k for k, t in enumerate(z["obs_ticks"])}
    for t in range(int(z["n_ticks"])):
        act = np.zeros((1, scene.veh_cap), np.float32)
        a_in = r["actions_in", t]
        act[0, :len(a_in)] = a_in
        o = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        np.testing.assert_array_equal(o["ids"][:, 1:3], r["ids", t], err_msg="%s ids t=%d" % (name, t))
        np.testing.assert_array_equal(o["ids"][:, 3], r["uid", t])
        np.testing.assert_array_equal(o["cpv"], r["cpv", t][:, 0], err_msg="%s cpv t=%d" % (name, t))
        np.testing.assert_array_equal(o["status"] & 1, r["done", t])
        np.testing.assert_array_equal((o["status"] >> 1) & 1, r["removed", t])
        P.assert_rel(o["reward"], r["reward", t], "%s reward t=%d" % (name, t))
        P.assert_rel(o["jerk_sum"][(o["status"] & 4) != 0], r["jerks", t], "%s jerks t=%d" % (name, t))
        assert o["collisions"][0] == z["t_collisions"][t] and o["lock"][0] == z["t_lock"][t], (name, t)
        assert o["n_removed"][0] == z["t_n_removed"][t]
        if t in obs_at:
            P.assert_rel(o["obs"], r["obs", obs_at[t]], "%s obs t=%d" % (name, t))
    st = scene.get_state()
    t = int(z["n_ticks"]) - 1
    V = int(st["lane_n"][0].sum())
    for k in ("p", "v", "a", "jerk_sum", "collision", "step", "uid", "lock_a"):
        np.testing.assert_array_equal(st[k][0, :V], r["post_" + k, t], err_msg="%s final %s" % (name, k))
    np.testing.assert_array_equal(st["head_lane"][0], z["t_head_lane"][t])
    np.testing.assert_array_equal(st["head_j"][0], z["t_head_j"][t])
    assert int(st["passed_step_total"][0]) == int(z["t_passed_step_total"][t])


def test_crafted_order_dependence_cases():
    """Q1-Q6 crafted states from the reference, teacher-forced through set_state."""
    z, r = load_crafted()
    for c, name in enumerate(str(n) for n in z["names"]):
        snap = {k: r["in_" + k, c] for k in
                ["p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "control", "finish",
                 "lock", "lock_a", "row0"] + STATE_KEYS_E}
        scene = P.make_scene(BACKEND, 1, vm=5, veh_cap=64, agent_cap=64)
        cap = scene.veh_cap                      # rounded up to a capacity class
        scene.reset(r["table", c], warmup=False)
        scene.set_state(P.oracle_state_for_device(snapshot_to_state(snap, 1, cap)))
        act = np.zeros((1, cap), np.float32)
        a_in = r["actions_in", c]
        act[0, :len(a_in)] = a_in
        o = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        np.testing.assert_array_equal(o["ids"][:, 1:3], r["ids", c], err_msg=name)
        np.testing.assert_array_equal(o["cpv"], r["cpv", c][:, 0], err_msg=name + " cpv")
        np.testing.assert_array_equal(o["status"] & 1, r["done", c], err_msg=name + " done")
        np.testing.assert_array_equal((o["status"] >> 1) & 1, r["removed", c], err_msg=name + " removed")
        P.assert_rel(o["reward"], r["reward", c], name + " reward")
        P.assert_rel(o["obs"], r["obs", c], name + " obs")
        P.assert_rel(o["jerk_sum"][(o["status"] & 4) != 0], r["jerks", c], name + " jerks")
        assert o["collisions"][0] == r["collisions", c][0], name
        assert o["lock"][0] == r["lock", c][0], name
        assert o["n_removed"][0] == r["n_removed", c][0], name
        st = scene.get_state()
        V = int(st["lane_n"][0].sum())
        assert V == len(r["post_p", c]), name
        for k in ("p", "v", "a", "jerk_sum", "collision", "step", "uid", "lock_a"):
            np.testing.assert_array_equal(st[k][0, :V], r["post_" + k, c], err_msg="%s %s" % (name, k))
        np.testing.assert_array_equal(st["flags"][0, :V] & 1, r["post_control", c], err_msg=name)
        np.testing.assert_array_equal((st["flags"][0, :V] >> 2) & 1, r["post_lock", c], err_msg=name)
        for k in ("lane_n", "veh_rec", "head_lane", "head_j"):
            np.testing.assert_array_equal(st[k][0], r["post_" + k, c], err_msg="%s %s" % (name, k))
        P.assert_rel(st["row0"][0, :V], r["post_row0", c], name + " row0")


def check_neighbour_sources(backend, ticks=160):
    """pve_outputs.nbr_src: every observation row is bit for bit the row its source code names."""
    B = 3
    tabs = synthetic_arrivals(B, 1000, 40.0, seed=5, rows=32)
    scene = P.make_scene(backend, B, vm=6, neighbour_sources=True)
    scene.reset(tabs, warmup=True)
    rng = np.random.RandomState(4)
    seen = {"zero": 0, "new": 0, "prev": 0}
    for t in range(ticks):
        prev = scene.row0().detach().cpu().numpy().copy()           # rows stored by the previous tick, slot order
        ctrl = scene.control_mask().detach().cpu().numpy()
        out = scene.step(P.to_device_actions(scene, P.random_actions(rng, ctrl)))
        n = out.n_agents
        obs = out.obs[:n].detach().cpu().numpy()
        src = out.nbr_src[:n].detach().cpu().numpy().astype(np.int32)
        ids = out.ids[:n].detach().cpu().numpy()
        off = out.agent_offset.detach().cpu().numpy()
        assert np.all(src[:, 7] == -1)
        for r in range(n):
            b = ids[r, 0]
            assert src[r, 0] == r - off[b]                          # entry 0: the agent itself
            for k in range(1, 7):
                v = src[r, k]
                if v < 0:
                    want = np.zeros(28, np.float32); seen["zero"] += 1
                elif v & 0x4000:
                    want = prev[b, v & 0x3FFF]; seen["prev"] += 1
                else:
                    want = obs[off[b] + v, 0]; seen["new"] += 1
                assert np.array_equal(obs[r, k], want), (t, r, k, v)
    assert min(seen.values()) > 100, seen


def test_neighbour_sources_name_the_copied_rows():
    check_neighbour_sources(BACKEND)


def edge_tables():
    """Ragged batch: a normal intersection, one whose table runs out after three arrivals per lane, one fed on a
    single lane only, one whose first arrival is 3 000 s away (the warm-up jumps there; sparse traffic afterwards),
    one with a single early arrival followed by that sparse traffic."""
    base = synthetic_arrivals(5, 1000, 40.0, seed=31, rows=24)
    far = 3000.0                            # beyond the rollout, inside the clock table of to_spawn_ticks
    t = base.copy()
    t[1, 3:, :] = 0.0                       # zero-padded tail like the shipped fixtures: table exhausted
    late = far + 10.0 * np.arange(t.shape[1])[:, None]
    t[2, :, 1:] = late                      # only lane 0 receives traffic
    t[3] = late                             # first arrival far away: the warm-up jumps there (TIS:214-220)
    t[4] = late
    t[4, 0, 7] = 2.0                        # one vehicle in the whole rollout
    return t


def test_ragged_and_empty_intersections():
    scene, n = free_run(BACKEND, edge_tables(), vm=5, ticks=320, seed=6)
    st = scene.get_state()
    assert st["id_seq"][1] == 36 and st["id_seq"][4] == 1 and st["tick"][3] > 30000 and n > 5000


def test_largest_capacity_class():
    """Headway 0.7 s on all 12 lanes with the all-brake policy fills an intersection to 483 vehicles / 407 agents:
    only the 576/416 class holds that (maximum sizes; SURVEY 8: V <= 576, A <= 408)."""
    scene, n = free_run(BACKEND, stress_arrivals(1, 60.0, headway=0.7), vm=5, ticks=420, seed=3, veh_cap=576,
                        agent_cap=416, policy="brake")
    st = scene.get_state()
    assert st["n_veh"][0] > 400 and st["n_ctrl"][0] > 384 and st["overflow"][0] == 0 and n > 100000
