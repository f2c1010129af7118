"""The reference's ``batch_test`` (main.py:543-583) run here, in the build container: the UNMODIFIED scene, the shipped
checkpoint's actor (numpy restatement), 36 000 ticks on each of the seven density files, one process per file.
Writes tests/golden/batch_test_reference.json (the report quantities per file) and tests/golden/batch_test_tables.npz
(the arrival tables up to the horizon of the run, so that the GPU box can run the same evaluation)."""
import json
import multiprocessing as mp
import os
import sys

import numpy as np
import scipy.io as scio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
DENSITIES = (1200, 1000, 900, 800, 600, 400, 200)          # main.py:545
TICKS = 36000


def run(density):
    import actor_oracle
    import ref_harness as H
    from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights
    w = ActorWeights.from_checkpoint("/root/reference/model_data/baseline")
    arr = scio.loadmat("/root/reference/data/test/arvTimeNewVeh_new_%d_12.mat" % density)["arvTimeNewVeh"]
    env = H.RefEnv(H.load_reference(), arr, vm=5).env
    coll = lock_total = 0
    jerk_total = 0.0
    for i in range(TICKS):
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
                env.step(lane, ind, amap.get((lane, ind), 0))
        ids, _, rew, actions, _, _, cpv, jerks, lock = env.scene_update()
        jerk_total += sum(jerks)
        lock_total += lock
        coll += sum(1 for k in range(len(actions)) if cpv[k][0] > 0)
        env.delete_vehicle()
    keep = int(np.max(np.sum((arr > 0) & (arr < TICKS * 0.1 + 20.0), axis=0))) + 2
    return density, dict(vehicles=int(env.id_seq), collisions=int(coll), passed=int(env.passed_veh),
                         passed_step_total=int(env.passed_veh_step_total), jerk_total=float(jerk_total),
                         lock_total=int(lock_total)), arr[:keep].astype(np.float64)


if __name__ == "__main__":
    with mp.get_context("fork").Pool(min(7, os.cpu_count() or 1)) as pool:
        res = pool.map(run, DENSITIES)
    json.dump({str(d): r for d, r, _ in res}, open(os.path.join(HERE, "batch_test_reference.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "batch_test_tables.npz"), **{"d%d" % d: t for d, _, t in res})
    for d, r, t in res:
        print(d, r, t.shape)
