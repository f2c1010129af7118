"""CPU tier: the 4-lane oracle (oracle/scene4_oracle.py) against rollouts of the UNMODIFIED reference scene with
``lane_num=4`` (tests/golden/rollout4_*.npz, minted by tests/golden/make_golden_n3.py) -- this is what pins it."""
import hashlib
import os

import numpy as np
import pytest

from golden_io import GOLDEN, Ragged
from oracle.scene4_oracle import Scene4Oracle
from oracle.scene8_oracle import Scene8Oracle

ROLLOUTS4 = ["mat1000_vm5", "mat1200_vm6", "mat400_vm5", "synth1800_brake", "mat1200_thr3"]
ROLLOUTS8 = ROLLOUTS4          # rollout8_*.npz: lane_num = 8, intentions from the recorded draw table


def load4(name, lanes=4):
    z = dict(np.load(os.path.join(GOLDEN, "rollout%d_%s.npz" % (lanes, name))))
    return z, Ragged(z)


@pytest.mark.parametrize("name", ROLLOUTS8)
def test_oracle8_reproduces_the_reference_rollout(name):
    test_oracle4_reproduces_the_reference_rollout(name, lanes=8)


@pytest.mark.parametrize("name", ROLLOUTS4)
def test_oracle4_reproduces_the_reference_rollout(name, lanes=4):
    z, r = load4(name, lanes)
    if lanes == 4:
        o = Scene4Oracle(vm=float(z["vm"]), collision_thr=float(z["collision_thr"]))
        o.reset(z["table"], warmup=True)
    else:
        o = Scene8Oracle(vm=float(z["vm"]), collision_thr=float(z["collision_thr"]))
        o.reset(z["table"], z["draws"], warmup=True)
    s = o.snapshot()
    assert s["tick"] == int(z["init_tick"]) and s["lane_n"] == z["init_lane_n"].tolist() and s["p"] == z["init_p"].tolist()
    obs_at = {int(t): k for k, t in enumerate(z["obs_ticks"])}
    rows = 0
    for t in range(int(z["n_ticks"])):
        out = o.step(r["actions_in", t])
        what = "%s tick %d" % (name, t)
        assert [list(x) for x in out["ids"]] == r["ids", t].tolist(), what
        assert out["uid"] == r["uid", t].tolist(), what
        assert out["cpv"] == r["cpv", t][:, 0].tolist(), what
        assert [int(x) for x in out["done"]] == r["done", t].tolist() and [int(x) for x in out["removed"]] == r["removed", t].tolist(), what
        assert [[list(c) for c in nb] for nb in out["nn"]] == r["nn", t].tolist(), what
        assert (out["collisions"], out["lock"], out["n_removed"]) == (int(z["t_collisions"][t]), int(z["t_lock"][t]), int(z["t_n_removed"][t])), what
        np.testing.assert_allclose(np.array(out["reward"], np.float64), r["reward", t], rtol=1e-12, atol=0, err_msg=what)
        np.testing.assert_array_equal(np.array(out["jerks"], np.float64), r["jerks", t], err_msg=what)
        obs = np.array(out["obs"], np.float64).reshape(-1, 7, 28)
        assert hashlib.sha256(obs.astype("<f8").tobytes()).digest() == z["obs_sha256"][t].tobytes(), what
        if t in obs_at:
            np.testing.assert_array_equal(obs, r["obs", obs_at[t]], err_msg=what)
        s = o.snapshot()
        for k in ("p", "v", "a", "jerk_sum"):
            np.testing.assert_array_equal(np.array(s[k], np.float64), r["post_" + k, t], err_msg=what + " " + k)
        for k in ("collision", "step", "uid", "control", "finish", "lock", "lock_a", "intention"):
            assert [int(x) for x in s[k]] == r["post_" + k, t].tolist(), what + " " + k
        for k in ("lane_n", "veh_rec", "head_lane", "head_j"):
            assert s[k] == z["t_" + k][t].tolist(), what + " " + k
        for k in ("tick", "id_seq", "passed_veh", "passed_step_total", "intention_re"):
            assert s[k] == int(z["t_" + k][t]), what + " " + k
        rows += len(out["ids"])
    assert rows > 3000
