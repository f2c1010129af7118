"""Mint the golden rollouts of the 4-lane and 8-lane intersections by RUNNING THE
REFERENCE SCENE with ``lane_num=4`` / ``lane_num=8`` (build container only: needs /root/reference):

    python tests/golden/make_golden_n3.py

Each ``rollout4_*.npz`` is a free-running rollout of the unmodified ``TrafficInteraction(..., lane_num=4)`` from its
constructor, driven like main.py:397-441 with seeded actions.  Per tick: the actions fed, ids / uid / reward / cpv /
done / removed / the six neighbours of every agent, the scalar outputs, a SHA-256 of the float64 observations (full
observations every ``OBS_EVERY`` ticks) and the whole post-tick state (vehicles incl. their intention, per-lane counts,
the heads of the twelve virtual lanes, the spawn counter ``intention_re``).  Layout as tests/golden/make_golden.py:
ragged arrays concatenated along axis 0 with ``<name>__off`` offsets.

``lane_num=8`` draws every new vehicle's intention with ``random.seed(); random.randint(0, 1)`` (TIS:382, 390), i.e. from
OS entropy: those rollouts (``rollout8_*.npz``) run with ``random.seed`` disabled and ``random.randint`` returning
``draws[k][i]`` for the k-th arrival of lane i -- a seeded table stored in the file; nothing else of the reference is
touched.
"""
import hashlib
import os
import sys

import numpy as np
import scipy.io as scio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import Ragged  # noqa: E402
from ref_harness import load_reference, ref_args  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals  # noqa: E402

OBS_EVERY = 20
NL, ND = 4, 12          # rebound by rollout() for the 8-lane scene


def snapshot(env, ticks):
    lane_n = np.array([len(env.veh_info[i]) for i in range(NL)], np.int32)
    vehs = [v for i in range(NL) for v in env.veh_info[i]]
    s = {"tick": np.int64(ticks), "lane_n": lane_n, "veh_rec": np.array(env.veh_rec, np.int32),
         "head_lane": np.full(ND, -1, np.int32), "head_j": np.full(ND, -1, np.int32),
         "id_seq": np.int64(env.id_seq), "passed_veh": np.int64(env.passed_veh),
         "passed_step_total": np.int64(env.passed_veh_step_total), "intention_re": np.int64(env.intention_re)}
    for d in range(ND):
        if len(env.virtual_lane_4[d]) > 0:
            s["head_lane"][d], s["head_j"][d] = env.virtual_lane_4[d][0][1], env.virtual_lane_4[d][0][2]
    for k, dt in (("p", float), ("v", float), ("a", float), ("jerk_sum", float), ("collision", np.int32), ("step", np.int32),
                  ("control", np.uint8), ("finish", np.uint8), ("lock", np.uint8), ("lock_a", np.int8), ("intention", np.uint8)):
        s[k] = np.array([v[k] for v in vehs], dtype=dt)
    s["uid"] = np.array([v["id_info"][0] for v in vehs], np.int32)
    s["row0"] = np.array([np.asarray(v["state"])[0] for v in vehs], np.float64).reshape(len(vehs), 28)
    return s


def rollout(mod, name, table, n_ticks, policy, vm, seed, collision_thr=2, lane_num=4):
    global NL, ND
    NL, ND = (4, 12) if lane_num == 4 else (8, 16)
    rng = np.random.RandomState(seed)
    nn_log = []
    draws = rng.randint(0, 2, size=(len(table), NL)).astype(np.uint8) if lane_num == 8 else None
    ctx = [0, 0]
    saved = (mod.random.seed, mod.random.randint)
    if lane_num == 8:
        mod.random.seed = lambda *a: None
        mod.random.randint = lambda a, b: int(draws[ctx[1]][ctx[0]])

    class Counted(mod.TrafficInteraction):
        n_updates = 0

        def scene_update(self):
            Counted.n_updates += 1
            return super().scene_update()

        def get_state(self, i, j, vl, direction):
            out = super().get_state(i, j, vl, direction)
            nn_log.append([list(c) for c in self.closer_cars])
            return out

        def add_new_veh(self, i):
            ctx[0], ctx[1] = i, self.veh_rec[i]
            return super().add_new_veh(i)

    env = Counted(np.asarray(table, np.float64), 150, ref_args(collision_thr), vm=vm, lane_num=NL)
    rag, obs_full, obs_ticks, sha = Ragged(), Ragged(), [], []
    scal = {k: [] for k in ("collisions", "lock", "n_removed", "tick", "id_seq", "passed_veh", "passed_step_total", "intention_re")}
    envs = {k: [] for k in ("lane_n", "veh_rec", "head_lane", "head_j")}
    init = snapshot(env, Counted.n_updates)
    n_rows = max_v = 0
    for t in range(n_ticks):
        vehs = [(i, j, v) for i in range(NL) for j, v in enumerate(env.veh_info[i])]
        if policy == "uniform":
            act = rng.uniform(-3, 3, size=len(vehs)).astype(np.float32)
        elif policy == "brake":
            act = np.full(len(vehs), -3.0, np.float32)
        else:
            act = rng.uniform(-4, 4, size=len(vehs)).astype(np.float32)
            pick = rng.random(len(vehs))
            act[pick < 0.15] = -3.0
            act[pick > 0.85] = 3.0
        for k, (i, j, v) in enumerate(vehs):
            if not v["control"]:
                act[k] = 0.0                                   # main.py:401-405
        for k, (i, j, v) in enumerate(vehs):
            env.step(i, j, float(act[k]))
        del nn_log[:]
        ids, st, rew, acts, coll, estm, cpv, jerks, lock = env.scene_update()
        A = len(ids)
        obs = np.array(st, np.float64).reshape(A, 7, 28)
        rag.add("actions_in", act)
        rag.add("ids", np.array(ids, np.int32).reshape(A, 2))
        rag.add("uid", np.array([env.veh_info[i][j]["id_info"][0] for i, j in ids], np.int32))
        rag.add("reward", np.array([float(r) for r in rew], np.float64))
        rag.add("cpv", np.array(cpv, np.int32).reshape(A, 2))
        rag.add("nn", np.array(nn_log, np.int32).reshape(A, 6, 2))
        rag.add("done", np.array([bool(env.veh_info[i][j]["Done"]) for i, j in ids], np.uint8))
        rag.add("removed", np.array([[i, j] in env.delete_veh for i, j in ids], np.uint8))
        rag.add("row0", obs[:, 0, :])
        rag.add("jerks", np.array([float(x) for x in jerks], np.float64))
        n_removed = len(env.delete_veh)
        env.delete_vehicle()
        post = snapshot(env, Counted.n_updates)
        for k in ("p", "v", "a", "jerk_sum", "collision", "step", "uid", "control", "finish", "lock", "lock_a", "intention"):
            rag.add("post_" + k, post[k])
        scal["collisions"].append(int(coll)); scal["lock"].append(int(lock)); scal["n_removed"].append(n_removed)
        for k in ("tick", "id_seq", "passed_veh", "passed_step_total", "intention_re"):
            scal[k].append(int(post[k]))
        for k in envs:
            envs[k].append(post[k])
        sha.append(np.frombuffer(hashlib.sha256(obs.astype("<f8").tobytes()).digest(), np.uint8))
        if t % OBS_EVERY == 0 or t == n_ticks - 1:
            obs_full.add("obs", obs)
            obs_ticks.append(t)
        n_rows += A
        max_v = max(max_v, len(post["p"]))
        assert int(post["veh_rec"].max()) < len(table) - 1, "arrival table too short for the rollout"
    mod.random.seed, mod.random.randint = saved
    data = {"table": np.asarray(table, np.float64), "vm": np.float64(vm), "collision_thr": np.float64(collision_thr),
            "lane_num": np.int64(NL), "n_ticks": np.int64(n_ticks), "obs_sha256": np.stack(sha), "obs_ticks": np.array(obs_ticks, np.int64)}
    if draws is not None:
        data["draws"] = draws
    for k, v in init.items():
        data["init_" + k] = v
    data.update(rag.pack())
    data.update(obs_full.pack())
    for k, v in scal.items():
        data["t_" + k] = np.array(v, np.int64)
    for k, v in envs.items():
        data["t_" + k] = np.stack(v).astype(np.int32)
    path = os.path.join(HERE, "rollout%d_%s.npz" % (lane_num, name))
    np.savez_compressed(path, **data)
    print("%-18s ticks=%d agent rows=%d max_V=%d collisions=%d locks=%d passed=%d  %.0f KB" % (
        name, n_ticks, n_rows, max_v, sum(scal["collisions"]), sum(scal["lock"]), scal["passed_veh"][-1], os.path.getsize(path) / 1024))


def mat4(density, rows, lanes=4):
    arr = scio.loadmat("/root/reference/data/test/arvTimeNewVeh_new_%d_12.mat" % density)["arvTimeNewVeh"]
    return np.ascontiguousarray(arr[:rows, :lanes].astype(np.float64))


SPECS = {
    "mat1000_vm5": (lambda: mat4(1000, 60), 600, "uniform", 5, 31, 2),
    "mat1200_vm6": (lambda: mat4(1200, 60), 500, "mixed", 6, 32, 2),
    "mat400_vm5": (lambda: mat4(400, 30), 600, "uniform", 5, 33, 2),
    "synth1800_brake": (lambda: synthetic_arrivals(1, 1800, 60.0, seed=9)[0][:, :NL].copy(), 420, "brake", 5, 34, 2),
    "mat1200_thr3": (lambda: mat4(1200, 60), 420, "uniform", 5, 35, 3.0),
}

SPECS8 = {
    "mat1000_vm5": (lambda: mat4(1000, 60, 8), 600, "uniform", 5, 41, 2),
    "mat1200_vm6": (lambda: mat4(1200, 60, 8), 500, "mixed", 6, 42, 2),
    "mat400_vm5": (lambda: mat4(400, 30, 8), 600, "uniform", 5, 43, 2),
    "synth1800_brake": (lambda: synthetic_arrivals(1, 1800, 60.0, seed=9)[0][:, :8].copy(), 420, "brake", 5, 44, 2),
    "mat1200_thr3": (lambda: mat4(1200, 60, 8), 420, "uniform", 5, 45, 3.0),
}

if __name__ == "__main__":
    mod = load_reference()
    for lanes, specs in ((4, SPECS), (8, SPECS8)):
        for name, (table, ticks, policy, vm, seed, thr) in specs.items():
            if len(sys.argv) < 2 or ("%d:%s" % (lanes, name)) in sys.argv[1:] or (lanes == 4 and name in sys.argv[1:]):
                rollout(mod, name, table(), ticks, policy, vm, seed, thr, lane_num=lanes)
