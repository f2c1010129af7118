"""Mint the golden fixtures under tests/golden/ by RUNNING THE REFERENCE SCENE.

Run in the build container only (needs /root/reference and scipy):

    python tests/golden/make_golden.py

Each ``rollout_*.npz`` is a free-running rollout of the unmodified reference
``TrafficInteraction`` (lane_num=12) from its constructor, driven like main.py:397-441 with
seeded actions; each tick's outputs and the post-tick state are recorded.  ``crafted.npz`` holds
single ticks started from hand-built states that exercise the order-dependent rules of
SURVEY.md section 3.3 (Q1-Q6).  The oracle (oracle/scene_oracle.c) is pinned against these files
by tests/test_oracle_golden.py; the CUDA path is compared with the oracle and with these files.

Ragged per-tick arrays are stored concatenated along axis 0 with an ``<name>__off`` offsets
array.  Full 7x28 observations are kept every ``OBS_EVERY`` ticks; every tick keeps a SHA-256 of
the float64 observation bytes.
"""
import hashlib
import os
import sys

import numpy as np
import scipy.io as scio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ref_harness import RefEnv, load_reference  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import stress_arrivals, synthetic_arrivals  # noqa: E402

OBS_EVERY = 20
MAT_DIR = "/root/reference/data/test"

STATE_KEYS_V = ["p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane", "uid", "control",
                "finish", "lock", "lock_a"]
STATE_KEYS_E = ["tick", "lane_n", "veh_rec", "head_lane", "head_j", "id_seq", "passed_veh",
                "passed_step_total"]
OUT_KEYS_A = ["ids", "reward", "cpv", "nn", "done", "removed", "uid", "row0"]
OUT_KEYS_E = ["collisions", "lock", "n_removed"]


class Ragged:
    def __init__(self):
        self.parts = {}

    def add(self, key, arr):
        self.parts.setdefault(key, []).append(np.asarray(arr))

    def pack(self, prefix=""):
        out = {}
        for key, parts in self.parts.items():
            off = np.zeros(len(parts) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(x) for x in parts])
            out[prefix + key] = np.concatenate(parts, axis=0) if parts else np.zeros(0)
            out[prefix + key + "__off"] = off
        return out


def mat_table(density, rows):
    arr = scio.loadmat(os.path.join(MAT_DIR, "arvTimeNewVeh_new_%d_12.mat" % density))["arvTimeNewVeh"]
    return np.ascontiguousarray(arr[:rows].astype(np.float64))


def draw_actions(policy, rng, snap):
    V = len(snap["p"])
    ctrl = snap["control"].astype(bool)
    if policy == "uniform":
        a = rng.uniform(-3.0, 3.0, size=V).astype(np.float32)
    elif policy == "brake":
        a = np.full(V, -3.0, dtype=np.float32)
    elif policy == "accel":
        a = np.full(V, 3.0, dtype=np.float32)
    elif policy == "mixed":
        # out-of-range values exercise the clamp (TIS:1502); extremes exercise safety/lock rules
        a = rng.uniform(-4.0, 4.0, size=V).astype(np.float32)
        pick = rng.random(V)
        a[pick < 0.15] = -3.0
        a[pick > 0.85] = 3.0
    else:
        raise ValueError(policy)
    if policy != "mixed":
        a[~ctrl] = 0.0          # main.py:401-405: uncontrolled vehicles are stepped with 0
    else:
        a[~ctrl & (rng.random(V) < 0.7)] = 0.0
    return a


def rollout(mod, name, table, n_ticks, policy, vm, seed, collision_thr=2):
    rng = np.random.RandomState(seed)
    env = RefEnv(mod, table, vm=vm, collision_thr=collision_thr)
    per_tick = Ragged()
    obs_full = Ragged()
    obs_ticks = []
    sha = []
    scal = {k: [] for k in OUT_KEYS_E + ["tick", "id_seq", "passed_veh", "passed_step_total"]}
    envs = {k: [] for k in ["lane_n", "veh_rec", "head_lane", "head_j"]}
    init = env.snapshot()
    n_agent_steps = 0
    max_v = 0
    for t in range(n_ticks):
        snap = env.snapshot() if t else init
        act = draw_actions(policy, rng, snap)
        out = env.tick(act)
        post = env.snapshot()
        per_tick.add("actions_in", act)
        for k in OUT_KEYS_A[:-1]:
            per_tick.add(k, out[k])
        per_tick.add("jerks", out["jerks"])
        for k in STATE_KEYS_V:
            per_tick.add("post_" + k, post[k])
        for k in OUT_KEYS_E:
            scal[k].append(int(out[k]))
        for k in ["tick", "id_seq", "passed_veh", "passed_step_total"]:
            scal[k].append(int(post[k]))
        for k in envs:
            envs[k].append(post[k])
        sha.append(np.frombuffer(hashlib.sha256(out["obs"].astype("<f8").tobytes()).digest(), np.uint8))
        if t % OBS_EVERY == 0 or t == n_ticks - 1:
            obs_full.add("obs", out["obs"])
            obs_ticks.append(t)
        n_agent_steps += len(out["ids"])
        assert int(post["veh_rec"].max()) < len(table) - 1, "arrival table too short for rollout"
        max_v = max(max_v, len(post["p"]))
    data = {"table": table, "vm": np.float64(vm), "collision_thr": np.float64(collision_thr),
            "n_ticks": np.int64(n_ticks), "obs_sha256": np.stack(sha),
            "obs_ticks": np.array(obs_ticks, dtype=np.int64)}
    for k in STATE_KEYS_V + STATE_KEYS_E + ["row0"]:
        data["init_" + k] = init[k]
    data.update(per_tick.pack())
    data.update(obs_full.pack())
    for k, v in scal.items():
        data["t_" + k] = np.array(v, dtype=np.int64)
    for k, v in envs.items():
        data["t_" + k] = np.stack(v).astype(np.int32)
    path = os.path.join(HERE, "rollout_%s.npz" % name)
    np.savez_compressed(path, **data)
    print("%-22s ticks=%d agent_steps=%d max_V=%d collisions_sum=%d locks=%d passed=%d  %.0f KB" % (
        name, n_ticks, n_agent_steps, max_v, sum(scal["collisions"]), sum(scal["lock"]),
        scal["passed_veh"][-1], os.path.getsize(path) / 1024))


# --------------------------------------------------------------------------------------------
# crafted single-tick cases
# --------------------------------------------------------------------------------------------
def blank_state(tick=500, veh_rec=40):
    s = {"tick": np.int64(tick), "lane_n": np.zeros(12, np.int32),
         "veh_rec": np.full(12, veh_rec, np.int32),
         "head_lane": np.full(12, -1, np.int32), "head_j": np.full(12, -1, np.int32),
         "id_seq": np.int64(1000), "passed_veh": np.int64(7), "passed_step_total": np.int64(1234)}
    s["_veh"] = [[] for _ in range(12)]
    return s


def add_veh(s, lane, p, v=8.0, a=0.0, control=True, collision=0, lock=False, lock_a=0, step=50,
            jerk_sum=3.5, finish=None, row0=None):
    rec = dict(p=p, v=v, a=a, control=control, collision=collision, lock=lock, lock_a=lock_a,
               step=step, jerk_sum=jerk_sum, finish=(not control) if finish is None else finish,
               row0=row0)
    s["_veh"][lane].append(rec)


def finalize(s, rng):
    vs = s.pop("_veh")
    flat = [(i, r) for i in range(12) for r in vs[i]]
    V = len(flat)
    s["lane_n"] = np.array([len(vs[i]) for i in range(12)], np.int32)
    for k, dt in [("p", float), ("v", float), ("a", float), ("jerk_sum", float), ("collision", np.int32),
                  ("step", np.int32), ("control", np.uint8), ("finish", np.uint8), ("lock", np.uint8),
                  ("lock_a", np.int8)]:
        s[k] = np.array([r[k] for _, r in flat], dtype=dt).reshape(V)
    s["uid"] = (900 + np.arange(V)).astype(np.int32)
    s["seq_in_lane"] = np.zeros(V, np.int32)
    k = 0
    for i in range(12):
        for j in range(len(vs[i])):
            s["seq_in_lane"][k] = int(s["veh_rec"][i]) - len(vs[i]) + j
            k += 1
    row0 = rng.uniform(-1, 1, size=(V, 28))
    for k, (_, r) in enumerate(flat):
        if r["row0"] is not None:
            row0[k] = r["row0"]
        if not r["control"]:
            row0[k] = 0.0
    s["row0"] = row0
    return s


def crafted_cases(rng):
    """Yield (name, state, actions)."""
    far = np.full((64, 12), 1e6)      # arrival table with no spawns in range ...
    L0, L1, L2 = 3.1415 / 2 * 7 * 2.5, 30.0, 3.1415 / 2 * 2.5

    # Q1: rear-end safety reads the already-stepped leader; three-car chain on a straight lane
    s = blank_state()
    add_veh(s, 1, 60.0, v=5.2); add_veh(s, 1, 63.0, v=9.0); add_veh(s, 1, 66.5, v=12.0)
    add_veh(s, 1, 90.0, v=7.0); add_veh(s, 1, 90.9, v=7.0000001)
    yield "q1_chain", finalize(s, rng), None, far

    # Q1b: leader uncontrolled => no safety override; follower much faster
    s = blank_state()
    add_veh(s, 4, -3.0, v=6.0, control=False); add_veh(s, 4, 2.0, v=12.5); add_veh(s, 4, 4.0, v=12.9)
    yield "q1_uncontrolled_leader", finalize(s, rng), None, far

    # Q2: stale head identity. head of lane 0 recorded as j=1; also a head pointing at another lane
    s = blank_state()
    add_veh(s, 0, 40.0); add_veh(s, 0, 55.0); add_veh(s, 0, 70.0)
    add_veh(s, 10, 50.0); add_veh(s, 10, 80.0)
    add_veh(s, 3, 45.0)
    s["head_lane"][0] = 0; s["head_j"][0] = 1            # fires on the "wrong" vehicle
    s["head_lane"][10] = 0; s["head_j"][10] = 0          # head of VL 10 is a lane-0 car: no fire
    s["head_lane"][3] = 3; s["head_j"][3] = 0
    s["head_lane"][7] = 7; s["head_j"][7] = 0            # lane 7 empty: head must stay stale
    yield "q2_stale_head", finalize(s, rng), None, far

    # Q2b: empty lane keeps its stale head through the tick and a spawn refills it
    s = blank_state(tick=500, veh_rec=3)
    add_veh(s, 1, 100.0)
    s["head_lane"][4] = 4; s["head_j"][4] = 0
    tab = np.full((64, 12), 1e6)
    tab[0:3] = np.array([[1.0], [2.0], [3.0]])      # ascending prefix: the table stays valid
    tab[3, 4] = 49.95   # spawns on lane 4 at this tick (tick 501 -> t=50.1)
    tab[3, 6] = 50.1000001
    tab[3, 7] = 50.0
    yield "q2_empty_lane_spawn", finalize(s, rng), None, tab

    # Q3/Q4: mutual nearest neighbours across conflicting lanes within 2 m in world space
    # lane 1 (straight W->E?) and lane 10 (straight, crossing); conflict point of (d=1,k=0).
    s = blank_state()
    add_veh(s, 1, 22.6, v=9.0); add_veh(s, 1, 40.0, v=9.0)
    add_veh(s, 10, 7.6 + 0.9, v=9.0); add_veh(s, 10, 30.0)
    add_veh(s, 4, 22.0, v=6.0); add_veh(s, 7, 23.0, v=6.0)
    yield "q4_cross_collision", finalize(s, rng), None, far

    # Q4b: same-lane overlap: three cars within 2 m, collision counts already > 0 on one
    s = blank_state()
    add_veh(s, 7, 80.0, v=9.0, collision=0); add_veh(s, 7, 81.0, v=9.0, collision=2)
    add_veh(s, 7, 81.5, v=9.0); add_veh(s, 7, 120.0)
    yield "q4_same_lane_pileup", finalize(s, rng), None, far

    # Q5: an uncontrolled vehicle carrying collision>0 overwrites the previous agent's reward
    s = blank_state()
    add_veh(s, 0, 90.0)
    add_veh(s, 1, -20.0, control=False, collision=1, v=10.0)
    add_veh(s, 1, 50.0)
    add_veh(s, 2, -40.0, control=False, collision=3, v=10.0)
    add_veh(s, 5, 60.0)
    yield "q5_ghost_collision", finalize(s, rng), None, far

    # finish / removal thresholds: p crossing 0 and -135 on every movement type
    s = blank_state()
    for lane in (0, 1, 2, 9, 10, 11):
        add_veh(s, lane, -134.2, control=False, v=10.0)
        add_veh(s, lane, -133.9, control=False, v=10.0)
        add_veh(s, lane, 0.4, v=9.0)
        add_veh(s, lane, 1.5, v=9.0)
        add_veh(s, lane, 20.0, v=9.0)
    yield "finish_and_remove", finalize(s, rng), None, far

    # Q6: exact position ties inside a virtual lane and exact |delta| ties in the kNN
    s = blank_state()
    for p in (30.0, 30.0, 34.0, 38.0, 42.0, 46.0, 46.0, 50.0, 54.0, 54.0, 58.0, 62.0):
        add_veh(s, 4, p, v=8.0, a=0.0)
    # lane 1 conflicts with lane 4 (k=0 of d=4 is lane 1): vd = 22.5 + (p1 - 7.5)
    for p in (19.0, 23.0, 27.0, 31.0, 35.0, 39.0):
        add_veh(s, 1, p, v=8.0)
    yield "q6_ties", finalize(s, rng), np.zeros(18, np.float32), far

    # deadlock ring: hand-built positions so that vir_header forms a cycle across 4 left lanes
    s = blank_state()
    for lane, p in ((0, 30.0), (3, 30.0), (6, 30.0), (9, 30.0)):
        add_veh(s, lane, p, v=5.0)
        add_veh(s, lane, p + 9.0, v=5.0)
    for lane, p in ((1, 33.0), (4, 33.0), (7, 33.0), (10, 33.0)):
        add_veh(s, lane, p, v=5.0)
    yield "lock_ring", finalize(s, rng), np.full(12, -3.0, np.float32), far

    # lock break: lock + lock_a set, p above and below 70 (TIS:1503-1505)
    s = blank_state()
    add_veh(s, 0, 95.0, a=1.5, lock=True, lock_a=1); add_veh(s, 0, 120.0, a=2.5, lock=True, lock_a=1)
    add_veh(s, 3, 60.0, a=-1.0, lock=True, lock_a=-1); add_veh(s, 3, 99.0, a=-2.7, lock=True, lock_a=-1)
    add_veh(s, 6, 99.0, a=0.25, lock=True, lock_a=0); add_veh(s, 9, 99.0, a=0.25, lock=False, lock_a=1)
    yield "lock_break", finalize(s, rng), None, far

    # arc geometry: vehicles on the left-turn and right-turn arcs of all four approaches
    s = blank_state()
    for ap in range(4):
        for frac in (0.1, 0.35, 0.6, 0.9):
            add_veh(s, 3 * ap + 0, L0 * frac, v=7.0)
        for frac in (0.2, 0.5, 0.8):
            add_veh(s, 3 * ap + 2, L2 * frac, v=7.0)
        add_veh(s, 3 * ap + 1, L1 * 0.5, v=7.0)
    yield "arc_geometry", finalize(s, rng), None, far


def crafted(mod):
    rng = np.random.RandomState(20261017)
    rag = Ragged()
    names = []
    for name, st, act, table in crafted_cases(rng):
        env = RefEnv(mod, np.full((4, 12), 0.05), vm=5)      # any table; state is overwritten
        env.env.arrive_time = np.asarray(table, dtype=np.float64)
        env.inject(st)
        V = len(st["p"])
        if act is None:
            act = rng.uniform(-3, 3, size=V).astype(np.float32)
            act[~st["control"].astype(bool)] = 0.0
        back = env.snapshot()
        for k in STATE_KEYS_V + ["row0"]:
            assert np.array_equal(np.asarray(back[k]), np.asarray(st[k])), (name, k)
        out = env.tick(act)
        post = env.snapshot()
        names.append(name)
        rag.add("table", table.reshape(-1, 12))
        rag.add("actions_in", act)
        for k in STATE_KEYS_V + ["row0"]:
            rag.add("in_" + k, st[k])
            rag.add("post_" + k, post[k])
        for k in STATE_KEYS_E:
            rag.add("in_" + k, np.atleast_1d(st[k]))
            rag.add("post_" + k, np.atleast_1d(post[k]))
        for k in OUT_KEYS_A[:-1] + ["obs", "jerks"]:
            rag.add(k, out[k])
        for k in OUT_KEYS_E:
            rag.add(k, np.atleast_1d(out[k]))
        print("crafted %-24s V=%d A=%d collisions=%d lock=%d removed=%d" % (
            name, V, len(out["ids"]), out["collisions"], out["lock"], out["n_removed"]))
    data = rag.pack()
    data["names"] = np.array(names)
    path = os.path.join(HERE, "crafted.npz")
    np.savez_compressed(path, **data)
    print("crafted.npz %.0f KB" % (os.path.getsize(path) / 1024))


SPECS = {
    # name: (table factory, ticks, policy, vm, seed) -- every shipped .mat fixture is covered
    "mat1000_vm5": (lambda: mat_table(1000, 40), 560, "uniform", 5, 11),
    "mat1200_vm6": (lambda: mat_table(1200, 40), 420, "uniform", 6, 12),
    "mat200_vm5": (lambda: mat_table(200, 12), 700, "uniform", 5, 13),
    "stress_brake": (lambda: stress_arrivals(1, 40.0)[0], 330, "brake", 5, 14),
    "synth1000_accel": (lambda: synthetic_arrivals(1, 1000, 60.0, seed=5)[0], 420, "accel", 5, 15),
    "synth800_mixed": (lambda: synthetic_arrivals(1, 800, 70.0, seed=6)[0], 520, "mixed", 6, 16),
    "mat400_vm6": (lambda: mat_table(400, 16), 520, "mixed", 6, 17),
    "mat600_vm5": (lambda: mat_table(600, 20), 460, "uniform", 5, 18),
    "mat800_vm6": (lambda: mat_table(800, 28), 440, "uniform", 6, 19),
    "mat900_vm5": (lambda: mat_table(900, 30), 440, "accel", 5, 20),
    # args.collision_thr, the one CLI flag that reaches the scene (main.py:104 -> TIS:32, 332, 1495)
    "mat1200_thr15": (lambda: mat_table(1200, 40), 420, "uniform", 5, 21, 1.5),
    "mat1000_thr3": (lambda: mat_table(1000, 40), 420, "mixed", 6, 22, 3.0),
}


def main(names):
    """``python make_golden.py``: everything; ``python make_golden.py name ...``: only those rollouts
    (``crafted`` = the crafted single-tick cases)."""
    mod = load_reference()
    for name, spec in SPECS.items():
        table, ticks, policy, vm, seed = spec[:5]
        if not names or name in names:
            rollout(mod, name, table(), ticks, policy, vm=vm, seed=seed, collision_thr=spec[5] if len(spec) > 5 else 2)
    if not names or "crafted" in names:
        crafted(mod)


if __name__ == "__main__":
    main(sys.argv[1:])
