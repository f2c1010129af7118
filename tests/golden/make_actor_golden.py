"""Mint the fixtures of the actor row.  Build container only: needs
/root/reference (checkpoint, scene source, .mat fixture).  Writes

* tests/golden/actor_agent1.npz          -- the twelve tensors of ``agent1actor`` from the shipped checkpoint
                                            model_data/baseline/66.cptk (read by checkpoint.read_bundle)
* tests/golden/actor_rollout_mat1000.npz -- config 1 of BASELINE.json: the UNMODIFIED reference scene driven for
                                            1000 ticks by that actor (numpy restatement, oracle/actor_oracle.py)
                                            on arvTimeNewVeh_new_1000_12.mat, exactly as main.py:397-441 does:
                                            per-tick integer trace, reward sums, the final report quantities of
                                            main.py:523-526, and a sample of (observation row, action) pairs.

The outcome must equal the one recorded in BASELINE.md section 2 (323 vehicles, 0 collisions, 281 passed,
pT-m 12.294 s, 548 lock events); the script asserts it.
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
from pve_mcc_for_unsignalized_intersection_b200.actor import PARAM_SPECS, ActorWeights  # noqa: E402

TICKS = 1000
MAT = "/root/reference/data/test/arvTimeNewVeh_new_1000_12.mat"
CKPT_DIR = "/root/reference/model_data/baseline"


def main():
    w = ActorWeights.from_checkpoint(CKPT_DIR)
    w.save_npz(os.path.join(HERE, "actor_agent1.npz"))
    arr = scio.loadmat(MAT)["arvTimeNewVeh"]
    mod = H.load_reference()
    env = H.RefEnv(mod, arr, vm=5).env                      # main.py:394: vm defaults to 5
    coll = lock_total = 0
    jerk_total = 0.0
    trace = np.zeros((TICKS, 6), dtype=np.int64)            # agents, id_seq, passed, lock (cum), collisions (cum), passed_step_total
    rsum = np.zeros(TICKS)
    rows, acts32, acts64 = [], [], []
    for i in range(TICKS):
        for lane in range(12):
            for ind, veh in enumerate(env.veh_info[lane]):
                a = 0
                if veh["control"]:
                    row = np.asarray(veh["state"][0], dtype=np.float64)[None, :]
                    a = float(actor_oracle.actor_forward(w, row, np.float32)[0])
                    if (i * 7 + ind) % 11 == 0 and len(rows) < 4000:
                        rows.append(row[0].copy())
                        acts32.append(a)
                        acts64.append(float(actor_oracle.actor_forward(w, row, np.float64)[0]))
                env.step(lane, ind, a)
        ids, _, rew, actions, _, _, cpv, jerks, lock = env.scene_update()
        jerk_total += sum(jerks)
        lock_total += lock
        coll += sum(1 for k in range(len(actions)) if cpv[k][0] > 0)
        trace[i] = (len(ids), env.id_seq, env.passed_veh, lock_total, coll, env.passed_veh_step_total)
        rsum[i] = float(np.sum(rew))
        env.delete_vehicle()
    ptm = float(env.passed_veh_step_total) / (env.passed_veh + 0.0001) * env.deltaT
    outcome = dict(vehicles=env.id_seq, collisions=coll, passed=env.passed_veh, ptm=ptm, lock=lock_total,
                   jerk_total=jerk_total)
    print(outcome)
    assert (env.id_seq, coll, env.passed_veh, lock_total) == (323, 0, 281, 548) and abs(ptm - 12.294) < 5e-4, outcome
    # only the part of the table the rollout can reach (TICKS * 0.1 s, plus slack) travels with the repo
    keep = int(np.max(np.sum((arr > 0) & (arr < TICKS * 0.1 + 20.0), axis=0))) + 2
    np.savez_compressed(os.path.join(HERE, "actor_rollout_mat1000.npz"), arrive_time=arr[:keep].astype(np.float64),
                        trace=trace, reward_sum=rsum, rows=np.asarray(rows), actions_f32=np.asarray(acts32),
                        actions_f64=np.asarray(acts64), ptm=ptm, jerk_total=jerk_total,
                        outcome=np.array([env.id_seq, coll, env.passed_veh, lock_total], dtype=np.int64))
    print("rows sampled:", len(rows), "table rows kept:", keep)


if __name__ == "__main__":
    main()
