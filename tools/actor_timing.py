#!/usr/bin/env python
"""Device time of the actor kernel on the bench workload: N back-to-back launches between two CUDA events
(the Python/ctypes launch overhead hides behind the previous kernel).  GPU only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights, BatchedActor  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scene = BatchedScene(B, SceneConfig(vm=5), device="cuda:0")
actor = BatchedActor(ActorWeights.from_npz(os.path.join(ROOT, "tests", "golden", "actor_agent1.npz")))
scene.reset(synthetic_arrivals(B, 1000, 120.0, seed=1000), warmup=True)
acts = torch.empty(B, scene.veh_cap, device="cuda")
for t in range(400):
    scene.step(actor.act(scene, out=acts))
n_agents = int(scene.control_mask().sum())
flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    flush.fill_(rep)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        actor.act(scene, out=acts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("actor kernel: %.4f ms per launch, %d controlled vehicles, %.2f TFLOP/s fp32, %.1f GB/s of rows"
          % (ms, n_agents, 2 * 5952 * n_agents / (ms * 1e-3) / 1e12, n_agents * 116 / (ms * 1e-3) / 1e9))
# the pair actor + step, back to back
for rep in range(2):
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        scene.step(actor.act(scene, out=acts))
    e1.record()
    torch.cuda.synchronize()
    print("actor + step: %.4f ms per tick (Python-driven)" % (e0.elapsed_time(e1) / 20))
