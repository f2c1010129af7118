"""Pin the CPU oracle against traces of the unmodified reference scene (tests/golden/*.npz)."""
import hashlib

import numpy as np
import pytest

from golden_io import ROLLOUTS, STATE_KEYS_E, load_crafted, load_rollout, snapshot_to_state
from oracle.oracle import OracleScene, scene_params

RTOL_TRANSCENDENTAL = 1e-12     # reward goes through tanh/log/pow of two different libms


def check_tick(o, z, r, t, what, obs_ref=None):
    A = len(r["ids", t])
    assert o["agent_offset"][1] == A, (what, t, o["agent_offset"][1], A)
    np.testing.assert_array_equal(o["ids"], r["ids", t], err_msg="%s ids t=%d" % (what, t))
    np.testing.assert_array_equal(o["uid"], r["uid", t])
    np.testing.assert_array_equal(o["nn"], r["nn", t], err_msg="%s nn t=%d" % (what, t))
    np.testing.assert_array_equal(o["cpv"], r["cpv", t][:, 0], err_msg="%s cpv t=%d" % (what, t))
    np.testing.assert_array_equal(o["status"] & 1, r["done", t], err_msg="%s done t=%d" % (what, t))
    np.testing.assert_array_equal((o["status"] >> 1) & 1, r["removed", t])
    np.testing.assert_array_equal(o["jerk_sum"][(o["status"] & 4) != 0], r["jerks", t])
    np.testing.assert_allclose(o["reward"], r["reward", t], rtol=RTOL_TRANSCENDENTAL, atol=0,
                               err_msg="%s reward t=%d" % (what, t))
    if obs_ref is not None:
        np.testing.assert_array_equal(o["obs"], obs_ref, err_msg="%s obs t=%d" % (what, t))


@pytest.mark.parametrize("name", ROLLOUTS)
def test_rollout_matches_reference(name):
    z, r = load_rollout(name)
    cap = 384 if name == "stress_brake" else 128
    orc = OracleScene(1, cap, scene_params(vm=float(z["vm"]), collision_thr=float(z["collision_thr"])))
    orc.reset(z["table"], warmup=True)
    st = orc.get_state()
    V0 = len(z["init_p"])
    assert st["lane_n"][0].sum() == V0
    np.testing.assert_array_equal(st["lane_n"][0], z["init_lane_n"])
    assert st["tick"][0] == int(z["init_tick"])
    np.testing.assert_array_equal(st["p"][0, :V0], z["init_p"])
    obs_at = {int(t): k for k, t in enumerate(z["obs_ticks"])}
    for t in range(int(z["n_ticks"])):
        act = np.zeros((1, cap), np.float32)
        a_in = r["actions_in", t]
        act[0, :len(a_in)] = a_in
        o = orc.step(act)
        assert o["overflow"] == 0 and o["q5_undefined"][0] == 0
        obs_ref = r["obs", obs_at[t]] if t in obs_at else None
        check_tick(o, z, r, t, name, obs_ref)
        sha = np.frombuffer(hashlib.sha256(o["obs"].astype("<f8").tobytes()).digest(), np.uint8)
        np.testing.assert_array_equal(sha, z["obs_sha256"][t], err_msg="%s obs sha t=%d" % (name, t))
        assert o["collisions"][0] == z["t_collisions"][t]
        assert o["lock"][0] == z["t_lock"][t], (name, t)
        assert o["n_removed"][0] == z["t_n_removed"][t]
        st = orc.get_state()
        V = int(st["lane_n"][0].sum())
        assert V == len(r["post_p", t])
        for k in ("p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "lock_a"):
            np.testing.assert_array_equal(st[k][0, :V], r["post_" + k, t], err_msg="%s %s t=%d" % (name, k, t))
        np.testing.assert_array_equal(st["flags"][0, :V] & 1, r["post_control", t])
        np.testing.assert_array_equal((st["flags"][0, :V] >> 1) & 1, r["post_finish", t])
        np.testing.assert_array_equal((st["flags"][0, :V] >> 2) & 1, r["post_lock", t])
        for k in ("lane_n", "veh_rec", "head_lane", "head_j"):
            np.testing.assert_array_equal(st[k][0], z["t_" + k][t], err_msg="%s %s t=%d" % (name, k, t))
        for k in ("tick", "id_seq", "passed_veh", "passed_step_total"):
            assert int(st[k][0]) == int(z["t_" + k][t]), (name, k, t)


def test_crafted_cases_match_reference():
    z, r = load_crafted()
    names = [str(n) for n in z["names"]]
    for c, name in enumerate(names):
        snap = {k: r["in_" + k, c] for k in
                ["p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "control", "finish",
                 "lock", "lock_a", "row0"] + STATE_KEYS_E}
        cap = 64
        orc = OracleScene(1, cap, scene_params(vm=5))
        orc.reset(r["table", c], warmup=False)
        orc.set_state(snapshot_to_state(snap, 1, cap))
        act = np.zeros((1, cap), np.float32)
        a_in = r["actions_in", c]
        act[0, :len(a_in)] = a_in
        o = orc.step(act)
        check_tick(o, z, r, c, name, obs_ref=r["obs", c])
        assert o["collisions"][0] == r["collisions", c][0], name
        assert o["lock"][0] == r["lock", c][0], name
        assert o["n_removed"][0] == r["n_removed", c][0], name
        st = orc.get_state()
        V = int(st["lane_n"][0].sum())
        assert V == len(r["post_p", c]), name
        for k in ("p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "lock_a"):
            np.testing.assert_array_equal(st[k][0, :V], r["post_" + k, c], err_msg="%s %s" % (name, k))
        np.testing.assert_array_equal(st["flags"][0, :V] & 1, r["post_control", c], err_msg=name)
        np.testing.assert_array_equal((st["flags"][0, :V] >> 2) & 1, r["post_lock", c], err_msg=name)
        for k in ("lane_n", "veh_rec", "head_lane", "head_j"):
            np.testing.assert_array_equal(st[k][0], r["post_" + k, c], err_msg="%s %s" % (name, k))
        np.testing.assert_array_equal(st["row0"][0, :V], r["post_row0", c], err_msg=name)
