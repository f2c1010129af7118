"""CPU tier: host-side logic, the C ABI's exported symbols, the oracle's own invariants."""
import ctypes
import os
import re

import numpy as np
import pytest

import parity as P
from golden_io import load_rollout
from pve_mcc_for_unsignalized_intersection_b200 import SceneConfig, _native, arrivals
from pve_mcc_for_unsignalized_intersection_b200.build import LIB, build_cuda
from pve_mcc_for_unsignalized_intersection_b200.distributed import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_clock_is_the_python_accumulation():
    t, ref = 0.0, [0.0]
    for _ in range(100000):
        t += 0.1
        ref.append(t)
    np.testing.assert_array_equal(arrivals.reference_clock(100000), np.array(ref))


def test_spawn_ticks_follow_the_float_comparison():
    # TIS:379 `current_time >= arrive_time`: 1.0 s spawns at tick 11 (ten additions give 0.9999999999999999)
    tab = np.array([[1.0] + [0.3] * 11, [2.0] + [0.65] * 11, [0.0] * 12], dtype=np.float64)
    ticks = arrivals.to_spawn_ticks(tab)
    assert ticks[0, 0] == 11 and ticks[1, 0] == 20 and ticks[0, 1] == 3 and ticks[1, 1] == 7
    assert (ticks[2] == arrivals.NEVER).all()            # zero padding = no more arrivals (Q10)
    clock = arrivals.reference_clock(40)
    for k in range(2):
        for i in range(12):
            n = ticks[k, i]
            assert clock[n] >= tab[k, i] and clock[n - 1] < tab[k, i]


def test_spawn_ticks_match_reference_trace():
    z, r = load_rollout("mat1000_vm5")
    ticks = arrivals.to_spawn_ticks(z["table"])
    # the reference's first vehicles appear at its warm-up tick
    assert int(z["init_tick"]) == ticks[0].min()
    assert (np.nonzero(z["init_lane_n"])[0] == np.nonzero(ticks[0] == ticks[0].min())[0]).all()


def test_synthetic_arrivals_have_fixture_statistics():
    arr = arrivals.synthetic_arrivals(64, 1000, 600.0, seed=3)
    head = np.diff(arr, axis=1)
    assert head.min() >= 1.0 - 1e-9 and (head > 0).all()  # cumulative sums: headways are exact only to rounding
    assert abs(np.mean(np.abs(head - 1.0) < 1e-9) - 0.242) < 0.02      # share of minimum headways
    assert abs(head.mean() - (1 + 3.6 * np.exp(-1 / 3.6))) < 0.05
    assert arr[:, -1, :].min() > 600.0
    st = arrivals.stress_arrivals(2, 30.0)
    assert np.allclose(np.diff(st, axis=1), 1.0) and st[0, 0, 0] == 1.0


def test_config_constants_match_the_reference_values():
    c = SceneConfig(vm=6)
    assert c.lane_len() == [3.1415 / 2 * 7 * 2.5, 30.0, 3.1415 / 2 * 2.5] and c.lane_in() == 135.0
    assert c.remove_p() == -135.0 and c.spawn_p(1) == 165.0
    a1, a2, b = c.virtual_distance_table()
    # threshold T = a1 - a2, offset C = b - a1 + a2
    T = [[a1[m][k] - a2[m][k] for k in range(4)] for m in range(2)]
    Cc = [[b[m][k] - a1[m][k] + a2[m][k] for k in range(4)] for m in range(2)]
    np.testing.assert_allclose(T[0], [14.188612, 9.469242, 18.019694, 15.811388], atol=1e-6)
    np.testing.assert_allclose(Cc[0], [5.549381, 8.550452, -8.550452, -8.060445], atol=1e-6)
    np.testing.assert_allclose(T[1], [7.5, 7.750943, 19.737992, 22.5], atol=1e-6)
    np.testing.assert_allclose(Cc[1], [15, 8.060445, -5.549381, -15], atol=1e-6)
    with pytest.raises(NotImplementedError):
        SceneConfig(lane_num=3)          # the reference's T-junction branch dies in its own constructor
    c8 = SceneConfig(lane_num=8)         # TIS:100-103, 341-342
    assert c8.lane_in() == 140.0 and c8.remove_p() == -140.0
    np.testing.assert_allclose(c8.lane_len(), [19.634375, 20.0, 3.926875], atol=1e-9)
    from oracle.scene8_oracle import geometry8
    lane_in, L, T, C1, C2 = geometry8(2.5)
    T8, C8, C28 = c8.eight_lane_tables()
    assert (lane_in, L) == (c8.lane_in(), c8.lane_len())
    for r in range(3):
        assert T8[r] == T[r] and C8[r] == C1[r] and C28[r] == C2[r]
    # the oracle derives the same constants independently
    from oracle.oracle import scene_params
    op = scene_params(vm=6)
    n = c.to_native(1, 128, 96, 96)
    assert op.lane_in == n.lane_in and list(op.lane_len) == list(n.lane_len) and op.remove_p == n.remove_p
    assert list(op.rot_cos) == list(n.rot_cos) and list(op.rot_sin) == list(n.rot_sin)
    assert n.vd_a1[1][1] == op.beta * 7 * 2.5 and n.vd_b[0][2] == op.gama * 7 * 2.5


def test_cuda_library_builds_loads_and_exports_every_declared_symbol():
    build_cuda()
    lib = _native.load_library()                 # dlopen works without a GPU (cudart is linked statically)
    assert lib.pve_backend().decode() == "cuda-sm_100a"
    header = open(os.path.join(ROOT, "include", "pve_mcc.h")).read()
    declared = sorted(set(re.findall(r"\b(pve_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 18
    raw = ctypes.CDLL(LIB)
    for name in declared:
        assert hasattr(raw, name), "libpve_mcc.so does not export %s" % name
    cfg = _native.PveConfig()
    assert lib.pve_default_config(ctypes.byref(cfg), 4, 5.0) == 0
    py = SceneConfig(vm=5).to_native(4, 160, 96, 384)
    assert cfg.lane_in == py.lane_in and list(cfg.lane_len) == list(py.lane_len)
    for m in range(2):
        assert list(cfg.vd_a1[m]) == list(py.vd_a1[m]) and list(cfg.vd_b[m]) == list(py.vd_b[m])


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch
    from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_native.NativeError):
        BatchedScene(2, device="cuda:0")
    with pytest.raises(_native.NativeError):
        BatchedScene(2, device="cpu")             # the CUDA library never runs on host memory


def test_pack_unpack_state_round_trip():
    rng = np.random.RandomState(0)
    B, cap = 3, 128
    st = _native.empty_state(B, cap)
    st["lane_n"][:] = rng.randint(0, 6, size=(B, 12))
    nv = st["lane_n"].sum(1)
    for k in ("p", "v", "a", "jerk_sum"):
        st[k][:] = rng.uniform(-100, 100, size=(B, cap))
    st["collision"][:] = rng.randint(0, 4, size=(B, cap)); st["step"][:] = rng.randint(0, 500, size=(B, cap))
    st["uid"][:] = rng.randint(0, 10**6, size=(B, cap)); st["flags"][:] = rng.randint(0, 8, size=(B, cap))
    st["lock_a"][:] = rng.randint(-1, 2, size=(B, cap)); st["tick"][:] = [5, 77, 1234]
    st["head_lane"][:] = rng.randint(-1, 12, size=(B, 12)); st["head_j"][:] = rng.randint(0, 9, size=(B, 12))
    st["head_j"][st["head_lane"] < 0] = -1
    st["veh_rec"][:] = rng.randint(0, 300, size=(B, 12))
    packed = _native.pack_state(st, B, cap)
    packed["hdr"]["n_veh"] = nv
    back = _native.unpack_state(packed, B, cap)
    live = np.arange(cap)[None, :] < nv[:, None]
    for k in ("p", "v", "a", "jerk_sum", "collision", "step", "uid", "flags", "lock_a"):
        np.testing.assert_array_equal(back[k], np.where(live, st[k], 0), err_msg=k)
    for k in ("tick", "lane_n", "veh_rec", "head_lane", "head_j"):
        np.testing.assert_array_equal(back[k], st[k], err_msg=k)


def test_shard_ranges_partition_the_batch():
    for n, w in ((65536, 8), (4096, 3), (7, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_emulation_is_labelled_and_refused_by_the_product_path():
    from emul.build_emul import build_emul
    from pve_mcc_for_unsignalized_intersection_b200.scene import BatchedScene
    lib = _native.load_library(build_emul())
    assert "emulation" in lib.pve_backend().decode()
    with pytest.raises(_native.NativeError):
        BatchedScene(2, device="cuda:0", _library=build_emul())   # emulation never pairs with a cuda device


def test_capacity_overflow_is_flagged_in_kernel_logic():
    scene = P.make_scene("emul", 1, veh_cap=64, agent_cap=48)
    scene.reset(arrivals.stress_arrivals(1, 40.0), warmup=True)
    import torch
    act = torch.full((1, scene.veh_cap), -3.0)
    for _ in range(330):
        scene.step(act)
    st = scene.get_state()
    assert st["overflow"][0] > 0 and st["n_veh"][0] <= scene.veh_cap and st["n_ctrl"][0] <= scene.agent_cap
    assert scene.stats()["overflow"] > 0
