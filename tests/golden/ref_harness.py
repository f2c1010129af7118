"""Harness around the UNMODIFIED reference scene, used only to mint golden fixtures.

This file runs only in the build container (it needs /root/reference).  Nothing in the
GPU tests, smoke() or bench.py imports it.  It loads
``/root/reference/traffic_interaction_scene.py`` from source with two neutral edits
(SURVEY.md section 8(c)):

* ``matplotlib`` is stubbed (the module is imported at TIS:2 but only ``Visible`` uses it);
* the logging-only loop TIS:371-375 is blanked, because it raises ``IndexError`` at low
  densities.  The edit is output-neutral.

It then offers ``snapshot`` / ``inject`` between the reference's list-of-dict vehicle
records and the flat arrays every other component of this repo uses, and ``ref_tick``
which drives one tick exactly the way main.py:398-407 + 441 does.
"""
import argparse
import sys
import types

import numpy as np

REF_SCENE = "/root/reference/traffic_interaction_scene.py"
NLANE = 12
OBS_W = 28


def load_reference(path=REF_SCENE):
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    src = open(path, encoding="utf-8").read().split("\n")
    assert src[370].strip().startswith("for v in self.virtual_lane_4[0]:"), src[370]
    for k in range(370, 375):
        src[k] = "        pass"
    mod = types.ModuleType("ref_scene")
    exec(compile("\n".join(src), path, "exec"), mod.__dict__)
    return mod


def ref_args(collision_thr=2):
    return argparse.Namespace(collision_thr=collision_thr, o_agent_num=6, c_mode="closer")


class RefEnv:
    """The reference ``TrafficInteraction`` plus a tick counter and neighbour capture."""

    def __init__(self, mod, arrive_time, vm=5, collision_thr=2):
        self.mod = mod
        self.ticks = 0
        self.nn_log = []
        outer = self

        class Counted(mod.TrafficInteraction):
            def scene_update(self_inner):
                outer.ticks += 1
                return super().scene_update()

            def get_state(self_inner, i, j, vl, direction):
                out = super().get_state(i, j, vl, direction)
                outer.nn_log.append([list(c) for c in self_inner.closer_cars])
                return out

        self.env = Counted(np.asarray(arrive_time, dtype=np.float64), 150, ref_args(collision_thr),
                           vm=vm, lane_num=NLANE)

    # ---- flat snapshot of the live state (SURVEY.md section 3.2 + H7) -------------------
    def snapshot(self):
        e = self.env
        lane_n = np.array([len(e.veh_info[i]) for i in range(NLANE)], dtype=np.int32)
        V = int(lane_n.sum())
        s = {
            "tick": np.int64(self.ticks),
            "lane_n": lane_n,
            "veh_rec": np.array(e.veh_rec, dtype=np.int32),
            "head_lane": np.full(NLANE, -1, dtype=np.int32),
            "head_j": np.full(NLANE, -1, dtype=np.int32),
            "id_seq": np.int64(e.id_seq),
            "passed_veh": np.int64(e.passed_veh),
            "passed_step_total": np.int64(e.passed_veh_step_total),
            "p": np.zeros(V), "v": np.zeros(V), "a": np.zeros(V), "jerk_sum": np.zeros(V),
            "collision": np.zeros(V, np.int32), "step": np.zeros(V, np.int32),
            "seq_in_lane": np.zeros(V, np.int32), "uid": np.zeros(V, np.int32),
            "control": np.zeros(V, np.uint8), "finish": np.zeros(V, np.uint8),
            "lock": np.zeros(V, np.uint8), "lock_a": np.zeros(V, np.int8),
            "row0": np.zeros((V, OBS_W)),
        }
        for d in range(NLANE):
            if len(e.virtual_lane_4[d]) > 0:
                s["head_lane"][d] = e.virtual_lane_4[d][0][1]
                s["head_j"][d] = e.virtual_lane_4[d][0][2]
        k = 0
        for i in range(NLANE):
            for veh in e.veh_info[i]:
                s["p"][k] = veh["p"]; s["v"][k] = veh["v"]; s["a"][k] = veh["a"]
                s["jerk_sum"][k] = veh["jerk_sum"]
                s["collision"][k] = veh["collision"]; s["step"][k] = veh["step"]
                s["seq_in_lane"][k] = veh["seq_in_lane"]; s["uid"][k] = veh["id_info"][0]
                s["control"][k] = bool(veh["control"]); s["finish"][k] = bool(veh["finish"])
                s["lock"][k] = bool(veh["lock"]); s["lock_a"][k] = int(veh["lock_a"])
                s["row0"][k] = np.asarray(veh["state"])[0]
                k += 1
        return s

    # ---- overwrite the reference's live state from a flat snapshot ------------------------
    def inject(self, s):
        e = self.env
        self.ticks = int(s["tick"])
        # the reference clock is an accumulated float64 sum of 0.1 (TIS:223)
        t = 0
        for _ in range(self.ticks):
            t += e.deltaT
        e.current_time = t
        e.veh_rec = [int(x) for x in s["veh_rec"]]
        e.veh_num = [int(x) for x in s["lane_n"]]
        e.id_seq = int(s["id_seq"])
        e.passed_veh = int(s["passed_veh"])
        e.passed_veh_step_total = int(s["passed_step_total"])
        e.virtual_lane.clear()
        for d in range(NLANE):
            if s["head_lane"][d] >= 0:
                # only element [0][1:3] of the stale list is ever read (TIS:1517)
                e.virtual_lane_4[d] = [[0.0, int(s["head_lane"][d]), int(s["head_j"][d]), 0.0, d]]
            else:
                e.virtual_lane_4[d] = []
        k = 0
        for i in range(NLANE):
            e.veh_info[i] = []
            for j in range(int(s["lane_n"][i])):
                m = i % 3
                state = np.zeros((7, OBS_W))
                state[0] = s["row0"][k]
                e.veh_info[i].append({
                    "intention": m, "buffer": [], "route": i, "count": 0,
                    "Done": bool(s["finish"][k]),
                    "p": float(s["p"][k]), "jerk": 0, "jerk_sum": float(s["jerk_sum"][k]),
                    "lock_a": int(s["lock_a"][k]), "lock": bool(s["lock"][k]),
                    "vir_header": [-1, -1], "vir_dis": 100,
                    "v": float(s["v"][k]), "a": float(s["a"][k]), "action": 0, "closer_p": 150,
                    "lane": i, "header": False, "reward": 10, "dis_front": 50,
                    "seq_in_lane": int(s["seq_in_lane"][k]), "control": bool(s["control"][k]),
                    "state": state, "step": int(s["step"][k]), "collision": int(s["collision"][k]),
                    "finish": bool(s["finish"][k]), "estm_collision": 0, "estm_arrive_time": 0.0,
                    "id_info": [int(s["uid"][k]), j],
                })
                k += 1
            # TIS:283 indexes this log by seq_in_lane; TIS:432 appends one slot per spawn
            e.veh_info_record[i] = [[] for _ in range(max(int(s["veh_rec"][i]), 1) + 1)]

    # ---- one tick, driven like main.py:398-407 then 441 ----------------------------------
    def tick(self, actions):
        """``actions``: one float per vehicle in (lane asc, j asc) order."""
        e = self.env
        k = 0
        for lane in range(NLANE):
            for ind, veh in enumerate(e.veh_info[lane]):
                e.step(lane, ind, float(actions[k]))
                k += 1
        assert k == len(actions)
        self.nn_log = []
        ids, st, rew, acts, coll, estm, cpv, jerks, lock = e.scene_update()
        A = len(ids)
        out = {
            "ids": np.array(ids, dtype=np.int32).reshape(A, 2),
            "obs": np.array(st, dtype=np.float64).reshape(A, 7, OBS_W),
            "reward": np.array([float(r) for r in rew], dtype=np.float64),
            "actions": np.array(acts, dtype=np.float64).reshape(A, 7),
            "collisions": np.int64(coll),
            "estm": np.int64(estm),
            "cpv": np.array(cpv, dtype=np.int32).reshape(A, 2),
            "jerks": np.array([float(x) for x in jerks], dtype=np.float64),
            "lock": np.int64(lock),
            "nn": np.array(self.nn_log, dtype=np.int32).reshape(A, 6, 2),
            "done": np.array([bool(e.veh_info[i][j]["Done"]) for i, j in ids], dtype=np.uint8),
            "removed": np.array([[i, j] in e.delete_veh for i, j in ids], dtype=np.uint8),
            "uid": np.array([e.veh_info[i][j]["id_info"][0] for i, j in ids], dtype=np.int32),
            "n_removed": np.int64(len(e.delete_veh)),
        }
        e.delete_vehicle()
        return out
