"""Mint the long closed-loop fixtures of the evaluation row (main.py:543-583 ``batch_test``).
Build container only: needs /root/reference.  For two of the seven density files of ``batch_test`` -- 400 veh/h (the
region where the reference's logging loop raises) and 1200 veh/h (the densest) -- the UNMODIFIED
reference scene is driven for 6000 ticks by the shipped checkpoint's actor (numpy restatement, oracle/actor_oracle.py)
exactly as main.py:553-575 does; the per-tick integer trace and the report quantities of main.py:576-581 are stored in
tests/golden/eval_mat<density>_6000.npz together with the part of the arrival table the run can reach.
"""
import os
import sys

import numpy as np
import scipy.io as scio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import actor_oracle  # noqa: E402
import ref_harness as H  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights  # noqa: E402

CKPT_DIR = "/root/reference/model_data/baseline"


def run(density, ticks):
    w = ActorWeights.from_checkpoint(CKPT_DIR)
    arr = scio.loadmat("/root/reference/data/test/arvTimeNewVeh_new_%d_12.mat" % density)["arvTimeNewVeh"]
    env = H.RefEnv(H.load_reference(), arr, vm=5).env
    coll = lock_total = 0
    jerk_total = 0.0
    trace = np.zeros((ticks, 6), dtype=np.int64)   # agents, id_seq, passed, lock (cum), collided agents (cum), passed_step_total
    for i in range(ticks):
        rows, where = [], []
        for lane in range(12):
            for ind, veh in enumerate(env.veh_info[lane]):
                if veh["control"]:
                    rows.append(np.asarray(veh["state"][0], dtype=np.float64))
                    where.append((lane, ind))
        acts = actor_oracle.actor_forward(w, np.asarray(rows).reshape(-1, 28), np.float32) if rows else []
        amap = {wh: float(a) for wh, a in zip(where, acts)}
        for lane in range(12):
            for ind, veh in enumerate(env.veh_info[lane]):
                env.step(lane, ind, amap.get((lane, ind), 0))                 # main.py:559-565
        ids, _, rew, actions, _, _, cpv, jerks, lock = env.scene_update()
        jerk_total += sum(jerks)
        lock_total += lock
        coll += sum(1 for k in range(len(actions)) if cpv[k][0] > 0)           # main.py:569-571
        trace[i] = (len(ids), env.id_seq, env.passed_veh, lock_total, coll, env.passed_veh_step_total)
        env.delete_vehicle()
    keep = int(np.max(np.sum((arr > 0) & (arr < ticks * 0.1 + 20.0), axis=0))) + 2
    out = os.path.join(HERE, "eval_mat%d_%d.npz" % (density, ticks))
    np.savez_compressed(out, arrive_time=arr[:keep].astype(np.float64), trace=trace, jerk_total=jerk_total,
                        outcome=np.array([env.id_seq, coll, env.passed_veh, lock_total, env.passed_veh_step_total], dtype=np.int64))
    print(density, "vehicles %d collided-agent-steps %d passed %d lock %d jerk %.3f, table rows kept %d, %.0f KB" % (
        env.id_seq, coll, env.passed_veh, lock_total, jerk_total, keep, os.path.getsize(out) / 1024))


if __name__ == "__main__":
    for d in ([int(x) for x in sys.argv[1:]] or [400, 1200]):
        run(d, 6000)
