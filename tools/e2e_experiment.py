"""Host-buffer path of one tick under the three PVE_HOST_ZEROCOPY modes (1: kernel reads actions and writes the small
outputs in pinned host memory; 2: actions staged by a DMA copy, outputs written in place; 0: everything staged), plus the
launch + synchronise floor of a device-resident tick.  Usage (GPU box): python tools/e2e_experiment.py [ticks]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = 4096
tabs = synthetic_arrivals(B, 1000, (400 + 3 * K) * 0.1 + 60.0, seed=1000)

def copy_rate(nbytes, h2d, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        (d if h2d else h).copy_(h if h2d else d, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        (d if h2d else h).copy_(h if h2d else d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


for nb in (2 << 20, 64 << 20):
    print("DMA copy %3d MiB: H2D %.1f GB/s  D2H %.1f GB/s" % (nb >> 20, copy_rate(nb, True), copy_rate(nb, False)), flush=True)

for mode in ("1", "2", "0", "floor"):
    os.environ["PVE_HOST_ZEROCOPY"] = "1" if mode == "floor" else mode
    scene = BatchedScene(B, SceneConfig(vm=6), device="cuda:0")
    scene.reset(tabs, warmup=True)
    pool = [(torch.rand(B, scene.veh_cap, device="cuda") * 6 - 3).contiguous() for _ in range(4)]
    for t in range(400):
        scene.step(pool[t % 4])
    torch.cuda.synchronize()
    host = scene.make_host_outputs()
    hact = [p.cpu().pin_memory() for p in pool]
    rows = 0
    if mode == "floor":
        for t in range(10):
            scene.step(pool[t % 4])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(K):
            out = scene.step(pool[t % 4])
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rows = K * out.n_agents
    else:
        for t in range(10):
            scene.step_host(hact[t % 4], host)
        t0 = time.perf_counter()
        for t in range(K):
            rows += scene.step_host(hact[t % 4], host)
        dt = time.perf_counter() - t0
    print("mode %-7s ms_per_tick %.4f  agent-steps/s %.3e" % (mode, dt / K * 1e3, rows / dt), flush=True)
    scene.close()
