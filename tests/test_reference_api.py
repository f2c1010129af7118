"""The B = 1 `TrafficInteraction` view driven exactly like main.py:397-441, against the oracle."""
import argparse

import numpy as np
import pytest

import parity as P
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
from pve_mcc_for_unsignalized_intersection_b200.reference_api import TrafficInteraction


def drive(backend, ticks=260):
    args = argparse.Namespace(collision_thr=2, o_agent_num=6, c_mode="closer")
    table = synthetic_arrivals(1, 1000, 40.0, seed=21, rows=32)[0]
    if backend == "cuda":
        env = TrafficInteraction(table, 150, args, vm=6, lane_num=12, device="cuda:0")
    else:
        from emul.build_emul import build_emul
        env = TrafficInteraction(table, 150, args, vm=6, lane_num=12, device="cpu", _library=build_emul())
    orc = P.make_oracle(1, vm=6, veh_cap=env.scene.veh_cap)
    orc.reset(table[None], warmup=True)
    rng = np.random.RandomState(3)
    passed_jerks = 0
    for t in range(ticks):
        act = np.zeros((1, env.scene.veh_cap), np.float32)
        k = 0
        for lane in range(12):                                   # MAIN:398-406
            for ind, veh in enumerate(env.veh_info[lane]):
                a = float(np.float32(rng.uniform(-3, 3))) if veh["control"] else 0.0
                if veh["control"]:
                    assert veh["state"].shape == (7, 28)
                env.step(lane, ind, a)
                act[0, k] = a
                k += 1
        ids, state_next, reward, actions, collisions, estm, cpv, jerks, lock = env.scene_update()   # MAIN:407
        o = orc.step(act)
        assert ids == [list(map(int, x)) for x in o["ids"]], t
        P.assert_rel(np.array(reward), o["reward"], "reward t=%d" % t)
        if len(ids):
            P.assert_rel(np.array(state_next), o["obs"], "obs t=%d" % t)
            P.assert_rel(np.array(actions), o["obs"][:, :, 2], "actions t=%d" % t)
        assert collisions == int(o["collisions"][0]) and estm == 0 and lock == int(o["lock"][0]), t
        assert [c[0] for c in cpv] == [int(c) for c in o["cpv"]], t
        P.assert_rel(np.array(jerks), o["jerk_sum"][(o["status"] & 4) != 0], "jerks t=%d" % t)
        passed_jerks += len(jerks)
        for (lane, j), st in zip(ids, o["status"]):              # MAIN:243-249 reads Done through ids
            assert env.veh_info[lane][j]["Done"] == bool(st & 1), t
        assert len(env.delete_veh) == int(o["n_removed"][0]), t
        env.delete_vehicle()                                     # MAIN:441
        ost = orc.get_state()
        assert [len(x) for x in env.veh_info] == [int(x) for x in ost["lane_n"][0]], t
        assert env.id_seq == int(ost["id_seq"][0]) and env.passed_veh == int(ost["passed_veh"][0])
        assert env.passed_veh_step_total == int(ost["passed_step_total"][0])
    assert passed_jerks > 0 and env.passed_veh > 0


def test_main_style_loop_on_kernel_emulation():
    drive("emul")


@pytest.mark.gpu
def test_main_style_loop_on_gpu():
    drive("cuda")
