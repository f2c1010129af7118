#!/usr/bin/env python
"""Timeline of the tcgen05 actor kernel (build with -DPVT_X_TRACE, see tools/ab_build.sh): clock64 stamps of the first
rounds of every group of the first CTAs, printed in microseconds at 1.9 GHz relative to the CTA's first stamp.  GPU only."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig, _native  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights, BatchedActor  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene  # noqa: E402

B = 4096
scene = BatchedScene(B, SceneConfig(vm=5), device="cuda:0")
actor = BatchedActor(ActorWeights.from_npz(os.path.join(ROOT, "tests", "golden", "actor_agent1.npz")))
scene.reset(synthetic_arrivals(B, 1000, 120.0, seed=1000), warmup=True)
acts = torch.empty(B, scene.veh_cap, device="cuda")
for t in range(300):
    scene.step(actor.act(scene, out=acts))
torch.cuda.synchronize()
actor.act(scene, out=acts)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["PVE_MCC_LIBRARY"])
buf = np.zeros(8 * 3 * 16 * 12, dtype=np.int64)
assert lib.pve_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
tr = buf.reshape(8, 3, 16, 12)
names = ["claim0", "claimed", "A1 stored", "mma1 issue", "mma1 issued", "mma1 done", "tmem read", "A2 stored", "mma2 issue", "mma2 done", "end"]
for cta in (0, 5):
    t0 = tr[cta][tr[cta] > 0].min()
    for g in range(3):
        for r in range(16):
            if tr[cta, g, r, 0] == 0:
                break
            print("cta %d group %d round %d: " % (cta, g, r) + "  ".join("%s %.2f" % (names[e], (tr[cta, g, r, e] - t0) / 1900.0) for e in range(11) if tr[cta, g, r, e] > 0))
