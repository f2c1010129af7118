#!/usr/bin/env python
"""BASELINE config 3: 65,536 intersections sharded across the GPUs of one box (8,192 per GPU), with a
teacher-forced check against the CPU oracle on a strided sample of every shard.

    python tests/config3_check.py                      # one GPU: its shard of 8,192
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/config3_check.py

Every rank free-runs its shard on the GPU with U(-3,3) actions.  Every CHECK_EVERY ticks the live state of
SAMPLE intersections (stride through the shard) is copied into the oracle, both sides take the same actions
from that identical state, and the outputs and the resulting states are compared: integers, flags and the
float64 state bit-exact, float32 outputs within 1e-5 relative (tests/parity.py).  Rank 0 prints one JSON
line; the only collective is the final reduction of the counters."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))      # test infrastructure: this script uses the CPU oracle as checker
import parity as P  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.distributed import shard_range  # noqa: E402


def slice_outputs(o, off, envs):
    """Rows of the dense outputs that belong to the intersections `envs`, re-packed like a small batch."""
    rows = np.concatenate([np.arange(off[e], off[e + 1]) for e in envs]) if len(envs) else np.zeros(0, np.int64)
    new_off = np.concatenate([[0], np.cumsum([off[e + 1] - off[e] for e in envs])]).astype(np.int64)
    ids = o["ids"][rows].copy()
    ids[:, 0] = np.repeat(np.arange(len(envs)), np.diff(new_off))
    return {"agent_offset": new_off, "obs": o["obs"][rows], "reward": o["reward"][rows], "ids": ids,
            "cpv": o["cpv"][rows], "status": o["status"][rows], "jerk_sum": o["jerk_sum"][rows],
            "collisions": o["collisions"][envs], "lock": o["lock"][envs], "n_removed": o["n_removed"][envs]}


def run_check(total_envs, ticks, sample, check_every, rank=0, world=1, device="cuda:0", seed=3, vm=6):
    lo, hi = shard_range(total_envs, rank, world)
    B = hi - lo
    horizon = ticks * 0.1 + 40.0
    tabs = synthetic_arrivals(B, 1000, horizon, seed=seed * 1000 + rank)
    scene = P.BatchedScene(B, P.SceneConfig(vm=vm), device=device)
    scene.reset(tabs, warmup=True)
    envs = np.unique(np.linspace(0, B - 1, sample).astype(np.int64))
    orc = P.OracleScene(len(envs), scene.veh_cap, P.scene_params(vm=vm), n_threads=min(16, os.cpu_count() or 1))
    orc.reset(tabs[envs], warmup=True)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + rank)
    checks = agent_rows = 0
    t0 = time.perf_counter()
    for t in range(ticks):
        act = (torch.rand(B, scene.veh_cap, device=device, generator=gen) * 6.0 - 3.0) * scene.control_mask()
        if t % check_every == 0 or t == ticks - 1:
            st = scene.get_state()
            sub = {k: np.ascontiguousarray(v[envs]) for k, v in st.items()}
            sub["row0"] = sub["row0"].astype(np.float64)
            orc.set_state(sub)                                        # identical state on both sides
            o_ref = orc.step(act[envs].cpu().numpy())
            out = scene.step(act)
            o_dev = P.outputs_to_numpy(out)
            P.compare_outputs(slice_outputs(o_dev, o_dev["agent_offset"], envs), o_ref,
                              "rank %d tick %d (teacher-forced sample)" % (rank, t))
            st2 = scene.get_state()
            P.compare_states({k: v[envs] for k, v in st2.items()}, orc.get_state(), "rank %d tick %d" % (rank, t))
            checks += 1
            agent_rows += len(o_ref["reward"])
        else:
            scene.step(act)
    torch.cuda.synchronize()
    s = scene.stats()
    assert s["overflow"] == 0
    return {"rank": rank, "envs": B, "ticks": ticks, "checks": checks, "sampled_envs": int(len(envs)),
            "agent_rows_compared": int(agent_rows), "agent_steps": s["agent_steps"], "seconds": time.perf_counter() - t0}, scene


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    total = int(os.environ.get("PVE_CONFIG3_ENVS", str(8192 * world)))
    ticks = int(os.environ.get("PVE_CONFIG3_TICKS", "400"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res, scene = run_check(total, ticks, sample=64, check_every=25, rank=rank, world=world, device="cuda:%d" % local)
    vec = torch.tensor([res["checks"], res["agent_rows_compared"], res["agent_steps"], res["envs"]],
                       dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(vec)                                          # end-of-rollout reduction only
    if rank == 0:
        print(json.dumps({"config": "BASELINE config 3", "n_gpus": world, "intersections": int(vec[3].item()),
                          "ticks": ticks, "teacher_forced_checks": int(vec[0].item()),
                          "agent_rows_compared_bit_exact_ints_1e-5_floats": int(vec[1].item()),
                          "agent_steps_simulated": vec[2].item(), "result": "parity ok on every rank"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
