#!/usr/bin/env python
"""The reference's ``batch_test`` (main.py:543-583: 36 000 ticks on each of the seven density files, pretrained actor in
the loop) on the GPU, next to the same run of the unmodified reference scene recorded in the build container
(tests/golden/batch_test_reference.json, tables in tests/golden/batch_test_tables.npz).  GPU box:

    python tools/batch_test_gpu.py > gpurun_out/r02_batch_test.md
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pve_mcc_for_unsignalized_intersection_b200 import evaluate  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
ref = json.load(open(os.path.join(G, "batch_test_reference.json")))
tabs = np.load(os.path.join(G, "batch_test_tables.npz"))
w = ActorWeights.from_npz(os.path.join(G, "actor_agent1.npz"))
dens = list(evaluate.DENSITIES)
torch.zeros(1, device="cuda").item()          # the CUDA context exists before the clock starts
from pve_mcc_for_unsignalized_intersection_b200.actor import BatchedActor  # noqa: E402
BatchedActor(w).close()                        # ... and the library is loaded
torch.cuda.synchronize()
t0 = time.perf_counter()
res = evaluate.evaluate_tables([tabs["d%d" % d] for d in dens], w, ticks=36000, veh_cap=192, agent_cap=128)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("# `batch_test` (main.py:543-583): 36 000 ticks x 7 density files, pretrained actor in the loop\n")
print("GPU: all seven files side by side as the intersections of one scene (class 192/128), actor + step per tick, tallies from the "
      "per-intersection device counters: **%.2f s** for the whole evaluation (%.0f ticks/s).  Reference: the unmodified scene + numpy "
      "actor, one process per file, recorded by tests/golden/make_batch_test_golden.py (about 6 minutes per file on one core).\n" % (dt, 36000 / dt))
print("The closed loop is chaotic (the fp32 policy differs in the last bits between numpy, TensorFlow and the GPU), so the rows agree "
      "exactly only where the arrival table fixes them (vehicles) or traffic is light; tests/test_gpu_actor.py pins 400 veh/h exactly "
      "over 6000 ticks and holds 1200 veh/h to the spread of such perturbations.\n")
print("| file | | vehicles | collided agent-steps | passed | pT-m (s) | jerks / passed | lock_num |")
print("|---|---|---|---|---|---|---|---|")
for d, r in zip(dens, res):
    g = ref[str(d)]
    for who, x in (("reference", g), ("GPU", r)):
        print("| arvTimeNewVeh_new_%d_12.mat | %s | %d | %d | %d | %.4f | %.3f | %d |" % (
            d, who, x["vehicles"], x["collisions"], x["passed"], x["passed_step_total"] / (x["passed"] + 0.0001) * 0.1,
            x["jerk_total"] / x["passed"], x["lock_total"]))
print("\nGPU report lines (the format of main.py:576-581):\n\n```")
for d, r in zip(dens, res):
    print("./data/test/arvTimeNewVeh_new_%d_12.mat\n%s" % (d, r["report"]))
print("```")
