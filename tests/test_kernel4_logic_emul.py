"""CPU tier, row N3: the SOURCE of the 4-/8-lane kernel (csrc/scene_step4.cuh), compiled as a sequential emulation, against
rollouts of the unmodified reference scene with ``lane_num=4`` and ``lane_num=8`` (no oracle in between) and against the
4-/8-lane oracles on random tables.  The same checks run on the real CUDA build in tests/test_gpu_lane4.py."""
import numpy as np
import pytest
import torch

import parity as P
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene
from oracle.scene4_oracle import Scene4Oracle
from oracle.scene8_oracle import Scene8Oracle
from test_oracle4_golden import ROLLOUTS4, ROLLOUTS8, load4

BACKEND = "emul"


CAPS4 = None            # (veh_cap, agent_cap) of the scenes built here; None = the library's default class (128 / 96)


def make_scene4(backend, B, vm=5, collision_thr=2, lanes=4):
    cfg = SceneConfig(vm=vm, collision_thr=collision_thr, lane_num=lanes)
    caps = {} if CAPS4 is None else {"veh_cap": CAPS4[0], "agent_cap": CAPS4[1]}
    if backend == "cuda":
        return BatchedScene(B, cfg, device="cuda:0", **caps)
    from emul.build_emul import build_emul
    return BatchedScene(B, cfg, device="cpu", _library=build_emul(), **caps)


def check_state(st, snap, b, what):
    """Device state of intersection ``b`` against a flat snapshot (golden post-tick arrays or the oracle's)."""
    V = len(snap["p"])
    assert int(st["n_veh"][b]) == V, what
    for k in ("p", "v", "a", "jerk_sum"):
        np.testing.assert_array_equal(st[k][b, :V], np.asarray(snap[k], np.float64), err_msg=what + " " + k)
    for k in ("collision", "step", "uid", "lock_a", "intention"):
        np.testing.assert_array_equal(st[k][b, :V], np.asarray(snap[k]).astype(st[k].dtype), err_msg=what + " " + k)
    fl = np.asarray(snap["control"]).astype(np.uint8) | (np.asarray(snap["finish"]).astype(np.uint8) << 1) | (np.asarray(snap["lock"]).astype(np.uint8) << 2)
    np.testing.assert_array_equal(st["flags"][b, :V], fl, err_msg=what + " flags")


def run_golden(backend, name, lanes=4):
    z, r = load4(name, lanes)
    scene = make_scene4(backend, 1, vm=float(z["vm"]), collision_thr=float(z["collision_thr"]), lanes=lanes)
    if lanes == 4:
        scene.reset(z["table"], warmup=True)
    else:
        scene.reset(z["table"], warmup=True, intention_draws=z["draws"])
    obs_at = {int(t): k for k, t in enumerate(z["obs_ticks"])}
    rows = 0
    for t in range(int(z["n_ticks"])):
        act = np.zeros((1, scene.veh_cap), np.float32)
        a_in = r["actions_in", t]
        act[0, :len(a_in)] = a_in
        o = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        what = "%s t=%d" % (name, t)
        np.testing.assert_array_equal(o["ids"][:, 1:3], r["ids", t], err_msg=what + " ids")
        np.testing.assert_array_equal(o["ids"][:, 3], r["uid", t], err_msg=what)
        np.testing.assert_array_equal(o["cpv"], r["cpv", t][:, 0], err_msg=what + " cpv")
        np.testing.assert_array_equal(o["status"] & 1, r["done", t], err_msg=what)
        np.testing.assert_array_equal((o["status"] >> 1) & 1, r["removed", t], err_msg=what)
        P.assert_rel(o["reward"], r["reward", t], what + " reward")
        P.assert_rel(o["obs"][:, 0, :], r["row0", t], what + " row 0")
        if t in obs_at:
            P.assert_rel(o["obs"], r["obs", obs_at[t]], what + " obs")
        fin = (o["status"] & 4) != 0
        P.assert_rel(o["jerk_sum"][fin], r["jerks", t], what + " jerks")
        assert (int(o["collisions"][0]), int(o["lock"][0]), int(o["n_removed"][0])) == (
            int(z["t_collisions"][t]), int(z["t_lock"][t]), int(z["t_n_removed"][t])), what
        st = scene.get_state()
        snap = {k: r["post_" + k, t] for k in ("p", "v", "a", "jerk_sum", "collision", "step", "uid", "control", "finish", "lock",
                                                "lock_a", "intention")}
        check_state(st, snap, 0, what)
        np.testing.assert_array_equal(st["lane_n"][0, :lanes], z["t_lane_n"][t], err_msg=what)
        np.testing.assert_array_equal(st["veh_rec"][0, :lanes], z["t_veh_rec"][t], err_msg=what)
        # (lane_num = 8: the device keeps the heads of routes 0-11 of 16; step() reads those of routes 0-7, TIS:1517)
        np.testing.assert_array_equal(st["head_lane"][0], z["t_head_lane"][t][:12], err_msg=what + " head lane")
        np.testing.assert_array_equal(st["head_j"][0], z["t_head_j"][t][:12], err_msg=what + " head j")
        assert (int(st["tick"][0]), int(st["id_seq"][0]), int(st["passed_veh"][0]), int(st["passed_step_total"][0])) == (
            int(z["t_tick"][t]), int(z["t_id_seq"][t]), int(z["t_passed_veh"][t]), int(z["t_passed_step_total"][t])), what
        assert (lanes == 8 or int(st["intention_re"][0]) == int(z["t_intention_re"][t]) % 3) and int(st["overflow"][0]) == 0, what
        rows += len(o["reward"])
    assert rows > 3000
    return scene


@pytest.mark.parametrize("name", ROLLOUTS4)
def test_golden_rollout4_direct(name):
    run_golden(BACKEND, name)


@pytest.mark.parametrize("name", ROLLOUTS8)
def test_golden_rollout8_direct(name):
    run_golden(BACKEND, name, lanes=8)


def free_run4(backend, B, density, ticks, seed, vm=5, lanes=4):
    """B intersections with their own tables, random actions, against B instances of the 4-/8-lane oracle."""
    tabs = synthetic_arrivals(B, density, ticks * 0.1 + 30.0, seed=seed)[:, :, :lanes].copy()
    scene = make_scene4(backend, B, vm=vm, lanes=lanes)
    if lanes == 4:
        scene.reset(tabs, warmup=True)
        orcs = [Scene4Oracle(vm=vm) for _ in range(B)]
        for b, o in enumerate(orcs):
            o.reset(tabs[b], warmup=True)
    else:
        draws = np.random.RandomState(seed + 100).randint(0, 2, size=(B,) + tabs.shape[1:]).astype(np.uint8)
        scene.reset(tabs, warmup=True, intention_draws=draws)
        orcs = [Scene8Oracle(vm=vm) for _ in range(B)]
        for b, o in enumerate(orcs):
            o.reset(tabs[b], draws[b], warmup=True)
    rng = np.random.RandomState(seed)
    n = 0
    for t in range(ticks):
        act = np.zeros((B, scene.veh_cap), np.float32)
        for b, o in enumerate(orcs):
            m = np.array(o.control_mask(), bool)
            act[b, :len(m)] = np.where(m, rng.uniform(-3, 3, size=len(m)), 0.0)
        dev = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        st = scene.get_state() if (t % 20 == 0 or t == ticks - 1) else None
        for b, o in enumerate(orcs):
            V = sum(len(x) for x in o.lanes)
            ref = o.step(act[b, :V])
            lo, hi = int(dev["agent_offset"][b]), int(dev["agent_offset"][b + 1])
            what = "env %d tick %d" % (b, t)
            assert hi - lo == len(ref["ids"]), what
            np.testing.assert_array_equal(dev["ids"][lo:hi, 1:3], np.array(ref["ids"], np.int32).reshape(-1, 2), err_msg=what)
            np.testing.assert_array_equal(dev["cpv"][lo:hi], np.array(ref["cpv"], np.int32), err_msg=what)
            np.testing.assert_array_equal(dev["status"][lo:hi] & 1, np.array(ref["done"], np.uint8), err_msg=what)
            P.assert_rel(dev["reward"][lo:hi], np.array(ref["reward"], np.float64), what + " reward")
            P.assert_rel(dev["obs"][lo:hi], np.array(ref["obs"], np.float64).reshape(-1, 7, 28), what + " obs")
            assert (int(dev["collisions"][b]), int(dev["lock"][b])) == (ref["collisions"], ref["lock"]), what
            if st is not None:
                check_state(st, o.snapshot(), b, what)
            n += hi - lo
    return n


def test_free_running_lane4_matches_oracle():
    assert free_run4(BACKEND, 3, 1400, 260, seed=5) > 4000


def test_free_running_lane8_matches_oracle():
    assert free_run4(BACKEND, 3, 1000, 260, seed=6, lanes=8) > 4000


def test_lane8_needs_draws_at_the_c_abi():
    """pve_reset refuses an 8-lane scene without intention draws (the reference draws them from OS entropy, TIS:382/390)."""
    scene = make_scene4(BACKEND, 1, lanes=8)
    ticks = torch.zeros(1, 4, 12, dtype=torch.int32) + 5
    rc = scene.lib.pve_reset(scene._h, ticks.data_ptr(), 4, 1, None)
    assert rc != 0 and b"pve_set_intention_draws" in scene.lib.pve_last_error(scene._h)
