"""Readers for the golden fixtures written by tests/golden/make_golden.py."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROLLOUTS = ["mat1000_vm5", "mat1200_vm6", "mat200_vm5", "stress_brake", "synth1000_accel", "synth800_mixed",
            "mat400_vm6", "mat600_vm5", "mat800_vm6", "mat900_vm5", "mat1200_thr15", "mat1000_thr3"]
STATE_KEYS_V = ["p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "control",
                "finish", "lock", "lock_a"]
STATE_KEYS_E = ["tick", "lane_n", "veh_rec", "head_lane", "head_j", "id_seq", "passed_veh",
                "passed_step_total"]


class Ragged:
    """``r[key, t]`` = the t-th part of a concatenated array with a ``key__off`` offsets table."""

    def __init__(self, npz):
        self.z = npz

    def __getitem__(self, kt):
        key, t = kt
        off = self.z[key + "__off"]
        return self.z[key][off[t]:off[t + 1]]

    def count(self, key):
        return len(self.z[key + "__off"]) - 1


def load_rollout(name):
    z = dict(np.load(os.path.join(GOLDEN, "rollout_%s.npz" % name)))
    return z, Ragged(z)


def load_crafted():
    z = dict(np.load(os.path.join(GOLDEN, "crafted.npz")))
    return z, Ragged(z)


def snapshot_to_state(snap, B, cap, row0_dtype=np.float64, env=0, state=None):
    """Place one golden snapshot (dict with STATE_KEYS_V/E + row0) into a batched flat state."""
    from oracle.oracle import empty_state
    st = state if state is not None else empty_state(B, cap)
    if row0_dtype != np.float64:
        st["row0"] = st["row0"].astype(row0_dtype)
    V = len(snap["p"])
    assert V <= cap, (V, cap)
    for k in ("tick", "id_seq", "passed_veh", "passed_step_total"):
        st[k][env] = int(np.asarray(snap[k]).reshape(-1)[0])
    for k in ("lane_n", "veh_rec", "head_lane", "head_j"):
        st[k][env] = np.asarray(snap[k]).reshape(12)
    for k in ("p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "lock_a"):
        st[k][env, :V] = snap[k]
        st[k][env, V:] = 0
    st["flags"][env, :V] = (snap["control"].astype(np.uint8) | (snap["finish"].astype(np.uint8) << 1)
                            | (snap["lock"].astype(np.uint8) << 2))
    st["flags"][env, V:] = 0
    st["row0"][env, :V] = snap["row0"]
    st["row0"][env, V:] = 0
    return st
