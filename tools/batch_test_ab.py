#!/usr/bin/env python
"""Where the time of a small closed-loop evaluation goes: the seven density files side by side (B = 7, class 192/128),
6000 ticks, driven tick by tick from Python or by pve_rollout, with each actor implementation.  GPU only.
Usage: PVE_ACTOR_IMPL=tc5|mma python tools/batch_test_ab.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig, evaluate  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights, BatchedActor  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
tabs = np.load(os.path.join(G, "batch_test_tables.npz"))
w = ActorWeights.from_npz(os.path.join(G, "actor_agent1.npz"))
tables = [tabs["d%d" % d] for d in evaluate.DENSITIES]
T = 6000
for mode in ("python loop", "pve_rollout", "step only", "actor only"):
    scene = BatchedScene(len(tables), SceneConfig(vm=5), veh_cap=192, agent_cap=128, device="cuda:0")
    actor = BatchedActor(w)
    scene.reset(evaluate.stack_tables(tables), warmup=True)
    acts = torch.zeros(len(tables), scene.veh_cap, device="cuda")
    actor.rollout(scene, 600, out=acts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if mode == "python loop":
        for _ in range(T):
            scene.step(actor.act(scene, out=acts))
    elif mode == "pve_rollout":
        actor.rollout(scene, T, out=acts)
    elif mode == "step only":
        for _ in range(T):
            scene.step(acts)
    else:
        for _ in range(T):
            actor.act(scene, out=acts)
    torch.cuda.synchronize()
    print("%-12s %6.1f us per tick (PVE_ACTOR_IMPL=%s)" % (mode, (time.perf_counter() - t0) / T * 1e6, os.environ.get("PVE_ACTOR_IMPL", "unset: by size")))
