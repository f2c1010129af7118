#!/usr/bin/env python
"""Wall-clock (SM cycles) spent between the kernel's phase boundaries, measured by CTA thread 0.

Builds a DEBUG copy of the library with -DPVE_PHASE_TIMING into gpurun_out/ (never the product
.so), runs the bench workload and prints the median cycles per phase.  GPU only."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig, _native  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.build import CSRC, INCLUDE, NVCC_FLAGS, find_nvcc  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
lib = os.path.join(out_dir, "libpve_mcc_timing.so")
subprocess.check_call([find_nvcc()] + NVCC_FLAGS + ["-DPVE_PHASE_TIMING", "-I", INCLUDE, "-I", CSRC,
                                                    os.path.join(CSRC, "pve_mcc.cu"), "-o", lib])
threads = int(sys.argv[1]) if len(sys.argv) > 1 else 128
veh_cap = int(sys.argv[2]) if len(sys.argv) > 2 else 128
agent_cap = int(sys.argv[3]) if len(sys.argv) > 3 else 96
B = 4096
scene = BatchedScene(B, SceneConfig(vm=6, zero_uncontrolled_actions=True), veh_cap=veh_cap, agent_cap=agent_cap, device="cuda:0",
                     threads=threads, _library=lib)
scene.reset(synthetic_arrivals(B, 1000, 120.0, seed=1000), warmup=True)
acts = [(torch.rand(B, scene.veh_cap, device="cuda") * 6 - 3).contiguous() for _ in range(8)]
for t in range(420):
    scene.step(acts[t % 8])
torch.cuda.synchronize()
scene.lib.pve_debug_stamps.restype = C.c_void_p
scene.lib.pve_debug_stamps.argtypes = [C.c_void_p]
ptr = scene.lib.pve_debug_stamps(scene._h)
st = scene._wrap(ptr, (B, 48), torch.int64, "<i8").cpu().numpy()
ext = st[:, 40:].copy()
st = st[:, :40]
n = int((st[0] >= 0).sum())
d = np.diff(st[:, :n], axis=1)
src = open(os.path.join(CSRC, "scene_step.cuh")).read()
print("phases recorded:", n - 1, " median total cycles per CTA:", int(np.median(st[:, n - 1])))
tot = np.median(st[:, n - 1])
for i in range(n - 1):
    print("boundary %2d -> %2d : median %6d  p90 %6d  (%4.1f%%)" % (i, i + 1, np.median(d[:, i]), np.percentile(d[:, i], 90),
                                                                   100 * np.median(d[:, i]) / tot))

# CTA lifetime on the SM: team end vs row movers' end, and how busy each SM slot was
mover_end = ext[:, 7] - ext[:, 0]
team_end = st[:, n - 1]
print("team end   (cycles from CTA start): median %d  p90 %d" % (np.median(team_end), np.percentile(team_end, 90)))
print("mover end  (cycles from CTA start): median %d  p90 %d" % (np.median(mover_end), np.percentile(mover_end, 90)))
print("mover ends after team in %.1f%% of CTAs; median (mover - team) %d" % (100 * (mover_end > team_end).mean(),
                                                                          np.median(mover_end - team_end)))
gt0, gt1, sm = ext[:, 1], ext[:, 2], ext[:, 3]
t_begin, t_end = gt0.min(), gt1.max()
print("kernel span by globaltimer: %.1f us; CTA lifetime (thread 0) median %.2f us, mean %.2f us" % (
    (t_end - t_begin) / 1e3, np.median(gt1 - gt0) / 1e3, (gt1 - gt0).mean() / 1e3))
busy = []
for s_ in np.unique(sm):
    m = sm == s_
    busy.append((gt1[m] - gt0[m]).sum() / 1e3)
print("per SM: CTAs %.1f, sum of CTA lifetimes %.1f us (=> mean resident CTAs %.2f over the span); last CTA start %.1f us, "
      "first CTA end %.1f us" % (len(sm) / len(busy), np.mean(busy), np.mean(busy) / ((t_end - t_begin) / 1e3),
                                 (gt0.max() - t_begin) / 1e3, (gt1.min() - t_begin) / 1e3))
A_, V_ = ext[:, 4], ext[:, 5]
print("A mean %.1f max %d; V mean %.1f max %d; corr(team_end, A) %.2f" % (A_.mean(), A_.max(), V_.mean(), V_.max(),
                                                                        np.corrcoef(team_end, A_)[0, 1]))
for lo, hi in ((0, 20), (20, 40), (40, 50), (50, 60), (60, 100)):
    m = (A_ >= lo) & (A_ < hi)
    if m.any():
        print("A in [%d,%d): %d CTAs, team end median %d cycles, lifetime %.2f us" % (lo, hi, m.sum(), np.median(team_end[m]),
                                                                                     np.median((gt1 - gt0)[m]) / 1e3))
