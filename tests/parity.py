"""Shared parity checks: product path (CUDA on the GPU box, or the test-only kernel-logic
emulation on CPU) against the CPU oracle and the golden traces."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200 import _native as N  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene  # noqa: E402
from oracle.oracle import OracleScene, scene_params  # noqa: E402

RTOL = 1e-5          # north-star tolerance: fp32 outputs against the float64 reference


def make_scene(backend, B, vm=5, collision_thr=2, veh_cap=128, agent_cap=96, threads=0, out_cap=None,
               neighbour_sources=False, zero_uncontrolled_actions=False):
    cfg = SceneConfig(vm=vm, collision_thr=collision_thr, zero_uncontrolled_actions=zero_uncontrolled_actions)
    if backend == "cuda":
        return BatchedScene(B, cfg, veh_cap=veh_cap, agent_cap=agent_cap, out_cap=out_cap, device="cuda:0",
                            threads=threads, neighbour_sources=neighbour_sources)
    from emul.build_emul import build_emul
    return BatchedScene(B, cfg, veh_cap=veh_cap, agent_cap=agent_cap, out_cap=out_cap, device="cpu",
                        _library=build_emul(), neighbour_sources=neighbour_sources)


def make_oracle(B, vm=5, collision_thr=2, veh_cap=128, n_threads=4):
    return OracleScene(B, veh_cap, scene_params(vm=vm, collision_thr=collision_thr), n_threads=n_threads)


def outputs_to_numpy(out):
    n = out.n_agents
    g = lambda t: t.detach().cpu().numpy()
    return {"agent_offset": g(out.agent_offset), "obs": g(out.obs[:n]), "reward": g(out.reward[:n]),
            "ids": g(out.ids[:n]), "cpv": g(out.cpv[:n]), "status": g(out.status[:n]),
            "jerk_sum": g(out.jerk_sum[:n]), "collisions": g(out.env_collisions),
            "lock": g(out.env_lock), "n_removed": g(out.env_removed)}


def assert_rel(x, y, what):
    """|x - y| <= RTOL * |y| elementwise (pure relative error; exact zeros must match)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    assert x.shape == y.shape, (what, x.shape, y.shape)
    err = np.abs(x - y)
    bad = err > RTOL * np.abs(y)
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError("%s: %d values beyond rtol=%g, first at %s: got %r want %r" % (
            what, int(bad.sum()), RTOL, tuple(i), x[tuple(i)], y[tuple(i)]))


def compare_outputs(dev, orc, what=""):
    """``dev``: outputs_to_numpy(StepOutputs); ``orc``: OracleScene.step() dict."""
    np.testing.assert_array_equal(dev["agent_offset"], orc["agent_offset"], err_msg=what + " agent_offset")
    np.testing.assert_array_equal(dev["ids"][:, 1:3], orc["ids"], err_msg=what + " ids")
    np.testing.assert_array_equal(dev["ids"][:, 3], orc["uid"], err_msg=what + " uid")
    B = len(orc["collisions"])
    env = np.repeat(np.arange(B), np.diff(orc["agent_offset"]))
    np.testing.assert_array_equal(dev["ids"][:, 0], env, err_msg=what + " env index")
    np.testing.assert_array_equal(dev["cpv"], orc["cpv"], err_msg=what + " cpv")
    np.testing.assert_array_equal(dev["status"], orc["status"], err_msg=what + " status")
    np.testing.assert_array_equal(dev["collisions"], orc["collisions"], err_msg=what + " collisions")
    np.testing.assert_array_equal(dev["lock"], orc["lock"], err_msg=what + " lock")
    np.testing.assert_array_equal(dev["n_removed"], orc["n_removed"], err_msg=what + " n_removed")
    assert_rel(dev["reward"], orc["reward"], what + " reward")
    assert_rel(dev["obs"], orc["obs"], what + " obs")
    fin = (orc["status"] & 4) != 0
    assert_rel(dev["jerk_sum"][fin], orc["jerk_sum"][fin], what + " jerks")


STATE_INT_KEYS = ("tick", "lane_n", "veh_rec", "head_lane", "head_j", "id_seq", "passed_veh",
                  "passed_step_total", "collision", "step", "uid", "flags", "lock_a")
STATE_F64_KEYS = ("p", "v", "a", "jerk_sum")


def compare_states(dev, orc, what="", exact=True):
    for k in STATE_INT_KEYS:
        np.testing.assert_array_equal(dev[k], orc[k], err_msg="%s state %s" % (what, k))
    for k in STATE_F64_KEYS:
        if exact:      # float64 state is computed with the reference's exact operation order
            np.testing.assert_array_equal(dev[k], orc[k], err_msg="%s state %s" % (what, k))
        else:
            assert_rel(dev[k], orc[k], "%s state %s" % (what, k))
    # the stored row 0 is live only for controlled vehicles (it is the actor input, MAIN:234-240, and
    # the neighbour row source, TIS:1332); rows of vehicles past the exit are dead in both worlds
    ctrl = ((orc["flags"] & 1) != 0)[:, :, None]
    assert_rel(np.where(ctrl, dev["row0"], 0), np.where(ctrl, orc["row0"], 0), what + " state row0")


def oracle_state_for_device(st):
    """Oracle state (float64 row0) -> the dict BatchedScene.set_state takes."""
    out = dict(st)
    out["row0"] = st["row0"].astype(np.float32)
    return out


def random_actions(rng, ctrl_mask, low=-3.0, high=3.0):
    a = rng.uniform(low, high, size=ctrl_mask.shape).astype(np.float32)
    a[~ctrl_mask] = 0.0            # MAIN:401-405
    return a


def to_device_actions(scene, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(scene.device)
