"""GPU tier, second file: configurations of the C ABI that the first parity file does not reach -- output capacity too
small, a non-default stream, two handles on one device, teacher forcing in the large capacity classes, the
dual-kernel mode against the single-kernel mode.  Everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest
import torch

import parity as P
from pve_mcc_for_unsignalized_intersection_b200.arrivals import stress_arrivals, synthetic_arrivals

pytestmark = pytest.mark.gpu


def _free_run_pair(scene, orc, tabs, ticks, seed, check_every=1, policy="uniform"):
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(seed)
    n = 0
    for t in range(ticks):
        ctrl = orc.control_mask()
        act = P.random_actions(rng, ctrl) if policy == "uniform" else np.where(ctrl, -3.0, 0.0).astype(np.float32)
        o_ref = orc.step(act)
        out = scene.step(P.to_device_actions(scene, act))
        if t % check_every == 0 or t == ticks - 1:
            P.compare_outputs(P.outputs_to_numpy(out), o_ref, "tick %d" % t)
        n += len(o_ref["reward"])
    P.compare_states(scene.get_state(), orc.get_state(), "final")
    return n


def test_out_cap_too_small_suppresses_rows_and_raises_overflow():
    """pve_config.out_cap smaller than a tick's rows: the intersections whose block does not fit emit nothing, the
    sticky overflow counter says so, the rows that fit are right and the state is unaffected (scene_step.cuh: out_ok)."""
    B = 8
    tabs = synthetic_arrivals(B, 1000, 40.0, seed=21, rows=32)
    scene = P.make_scene("cuda", B, vm=5, out_cap=120)
    orc = P.make_oracle(B, vm=5, veh_cap=scene.veh_cap)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(3)
    guard = torch.full((4096,), 7, dtype=torch.uint8, device="cuda")       # allocated right after the outputs
    suppressed = 0
    for t in range(300):
        act = P.random_actions(rng, orc.control_mask())
        o_ref = orc.step(act)
        out = scene.step(P.to_device_actions(scene, act))
        off = out.agent_offset.cpu().numpy()
        np.testing.assert_array_equal(off, o_ref["agent_offset"])             # offsets are always the full prefix
        fit = int(np.searchsorted(off, 120, side="right")) - 1                # intersections [0, fit) have their rows
        n_fit = int(off[fit])
        suppressed += B - fit
        for f, ref in (("reward", o_ref["reward"]), ("cpv", o_ref["cpv"]), ("status", o_ref["status"])):
            got = getattr(out, f)[:n_fit].cpu().numpy()
            if f == "reward":
                P.assert_rel(got, ref[:n_fit], "reward t=%d" % t)
            else:
                np.testing.assert_array_equal(got, ref[:n_fit], err_msg="%s t=%d" % (f, t))
        P.assert_rel(out.obs[:n_fit].cpu().numpy(), o_ref["obs"][:n_fit], "obs t=%d" % t)
    assert suppressed > 100
    assert bool((guard == 7).all())
    st = scene.get_state()
    assert st["overflow"].sum() > 0 and scene.stats()["overflow"] > 0
    ref = orc.get_state()
    for k in P.STATE_INT_KEYS + P.STATE_F64_KEYS:                           # the trajectory itself never deviates
        np.testing.assert_array_equal(st[k], ref[k], err_msg=k)


def test_non_default_stream():
    """Every entry point takes the caller's stream: run a rollout on a side stream."""
    B = 16
    tabs = synthetic_arrivals(B, 1000, 30.0, seed=8, rows=24)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        scene = P.make_scene("cuda", B, vm=6)
        orc = P.make_oracle(B, vm=6, veh_cap=scene.veh_cap)
        n = _free_run_pair(scene, orc, tabs, 220, seed=5, check_every=7)
        side.synchronize()
    assert n > 40000


def test_two_handles_on_one_device_do_not_interfere():
    """Two scenes with different tables and settings stepped alternately (own state, own side stream, own counters)."""
    ta = synthetic_arrivals(6, 1200, 30.0, seed=1, rows=30)
    tb = synthetic_arrivals(10, 600, 30.0, seed=2, rows=20)
    sa, sb = P.make_scene("cuda", 6, vm=5), P.make_scene("cuda", 10, vm=6, collision_thr=3)
    oa, ob = P.make_oracle(6, vm=5, veh_cap=sa.veh_cap), P.make_oracle(10, vm=6, collision_thr=3, veh_cap=sb.veh_cap)
    for s, o, t in ((sa, oa, ta), (sb, ob, tb)):
        s.reset(t, warmup=True)
        o.reset(t, warmup=True)
    rng = np.random.RandomState(12)
    for t in range(240):
        for s, o in ((sa, oa), (sb, ob)):
            act = P.random_actions(rng, o.control_mask())
            o_ref = o.step(act)
            P.compare_outputs(P.outputs_to_numpy(s.step(P.to_device_actions(s, act))), o_ref, "tick %d" % t)
    P.compare_states(sa.get_state(), oa.get_state(), "a")
    P.compare_states(sb.get_state(), ob.get_state(), "b")


@pytest.mark.parametrize("veh_cap,agent_cap,headway,fill", [(384, 320, 1.0, 260), (576, 416, 0.7, 330)])
def test_teacher_forced_stress_in_the_large_classes(veh_cap, agent_cap, headway, fill):
    """Classes 384/320 and 576/416 (512-thread CTAs, 16-bit lane prefix): state injected from the oracle every tick."""
    tabs = stress_arrivals(2, 60.0, headway=headway)
    scene = P.make_scene("cuda", 2, vm=5, veh_cap=veh_cap, agent_cap=agent_cap)
    orc = P.make_oracle(2, vm=5, veh_cap=scene.veh_cap)
    scene.reset(tabs, warmup=True)
    orc.reset(tabs, warmup=True)
    rng = np.random.RandomState(2)
    for t in range(fill):                                                   # fill the intersection (brake: worst occupancy)
        orc.step(np.where(orc.control_mask(), -3.0, 0.0).astype(np.float32))
    for t in range(60):
        st = orc.get_state()
        scene.set_state(P.oracle_state_for_device(st))
        ctrl = (st["flags"] & 1) != 0
        act = P.random_actions(rng, ctrl) if t % 2 else np.where(ctrl, -3.0, 0.0).astype(np.float32)
        o_ref = orc.step(act)
        o_dev = P.outputs_to_numpy(scene.step(P.to_device_actions(scene, act)))
        P.compare_outputs(o_dev, o_ref, "teacher-forced stress tick %d" % t)
        P.compare_states(scene.get_state(), orc.get_state(), "teacher-forced stress tick %d" % t)
    assert len(o_ref["reward"]) > 2 * 0.6 * agent_cap and o_ref["overflow"] == 0


@pytest.mark.parametrize("caps", [(128, 96), (192, 128)])
def test_dual_kernel_mode_equals_single_kernel_mode(caps, monkeypatch):
    """The default launch (two concurrent kernels: small-class CTAs + the intersections that do not fit) and the
    single-kernel launch (PVE_DUAL=0) give identical bytes, at a density where both kernels have work every tick;
    for both capacity classes that run in dual mode."""
    B = 192
    tabs = synthetic_arrivals(B, 1200, 45.0, seed=31, rows=48)
    monkeypatch.setenv("PVE_DUAL", "1")
    dual = P.make_scene("cuda", B, vm=5, veh_cap=caps[0], agent_cap=caps[1])
    monkeypatch.setenv("PVE_DUAL", "0")
    single = P.make_scene("cuda", B, vm=5, veh_cap=caps[0], agent_cap=caps[1])
    assert dual.launch_info["dual"] is True and single.launch_info["dual"] is False and dual.veh_cap == caps[0]
    dual.reset(tabs, warmup=True)
    single.reset(tabs, warmup=True)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    big_seen = 0
    for t in range(420):
        act = (torch.rand(B, dual.veh_cap, device="cuda", generator=gen) * 6 - 3) * dual.control_mask()
        a, b = dual.step(act.contiguous()), single.step(act.contiguous())
        n = a.n_agents
        assert n == b.n_agents
        for f in ("agent_offset", "env_collisions", "env_lock", "env_removed"):
            assert torch.equal(getattr(a, f), getattr(b, f)), (f, t)
        for f in ("obs", "reward", "ids", "cpv", "status", "jerk_sum"):
            assert torch.equal(getattr(a, f)[:n], getattr(b, f)[:n]), (f, t)
        st = dual.get_state() if t % 60 == 59 else None
        if st is not None:
            big_seen += int(((st["n_veh"] > 84) | (st["n_ctrl"] > 52)).sum())
    sa, sb = dual.get_state(), single.get_state()
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    assert big_seen > 0 and dual.stats()["overflow"] == 0


def test_pipelined_host_path_delivers_the_same_rows():
    """pve_step_host_async / pve_host_wait (three ticks in flight, DMA in and out, 16-byte agent records) against the
    plain device step of a twin scene fed with the same actions."""
    B = 96
    tabs = synthetic_arrivals(B, 1000, 30.0, seed=17, rows=24)
    twin, scene = P.make_scene("cuda", B, vm=6), P.make_scene("cuda", B, vm=6)
    twin.reset(tabs, warmup=True)
    scene.reset(tabs, warmup=True)
    pairs = scene.make_async_buffers()
    rng = np.random.RandomState(1)
    acts = [torch.from_numpy(rng.uniform(-3, 3, size=(B, scene.veh_cap)).astype(np.float32)).pin_memory() for _ in range(5)]
    want = []

    def check(tick):
        n, host = scene.host_wait()
        w = want.pop(0)
        assert n == len(w["reward"]) and n == int(host.agent_offset[-1]), tick
        rec = host.records(n)
        np.testing.assert_array_equal(rec["reward"], w["reward"])
        np.testing.assert_array_equal(rec["uid"], w["ids"][:, 3])
        np.testing.assert_array_equal(rec["lane"], w["ids"][:, 1])
        np.testing.assert_array_equal(rec["j"], w["ids"][:, 2])
        np.testing.assert_array_equal(rec["status"], w["status"])
        np.testing.assert_array_equal(rec["cpv"], np.minimum(w["cpv"], 255))
        np.testing.assert_array_equal(rec["jerk_sum"], w["jerk_sum"])
        np.testing.assert_array_equal(host.obs[:n].numpy(), w["obs"])
        np.testing.assert_array_equal(host.agent_offset.numpy(), w["agent_offset"])
        np.testing.assert_array_equal(host.env_lock.numpy(), w["lock"])

    for t in range(160):
        a = acts[t % 5]
        want.append(P.outputs_to_numpy(twin.step(a.cuda())))
        scene.step_host_async(a, pairs[t % 3], copy_obs=True)
        if t >= 2:
            check(t - 2)                       # tick t - 2 landed while ticks t - 1 and t are under way
    check(158)
    check(159)
    sa, sb = scene.get_state(), twin.get_state()
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    assert scene.stats()["agent_steps"] == twin.stats()["agent_steps"] > 100000
