"""ctypes binding of the CPU oracle (oracle/scene_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs.  The product package never imports this module.

Parity status: PINNED against traces of the unmodified reference scene
(tests/golden/*.npz, minted by tests/golden/make_golden.py; checked by
tests/test_oracle_golden.py).
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libscene_oracle.so")
NLANE, OBS_W, OBS_H, NN = 12, 28, 7, 6


def build(force=False):
    """Compile the C restatement (gcc only; a few seconds)."""
    src = [os.path.join(HERE, f) for f in ("scene_oracle.c", "scene_oracle.h", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(f) for f in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-B"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("dis_ctl", "lane_cw", "dt", "dt2", "vm", "vM", "am", "aM", "v0", "collision_thr",
                 "lane_in")] + [
        ("lane_len", C.c_double * 3), ("remove_p", C.c_double),
        ("cita", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
        ("gama", C.c_double), ("gama2", C.c_double),
        ("rot_cos", C.c_double * 4), ("rot_sin", C.c_double * 4)]


def scene_params(vm=5, collision_thr=2, dis_ctl=150, deltaT=0.1, vM=13, am=-3, aM=3, v0=10, lane_cw=2.5):
    """Constants of the reference constructor for lane_num=12, derived with its own expressions."""
    P = OrcParams()
    P.dis_ctl, P.lane_cw, P.dt, P.dt2 = dis_ctl, lane_cw, deltaT, pow(deltaT, 2)        # TIS:1529
    P.vm, P.vM, P.am, P.aM, P.v0, P.collision_thr = vm, vM, am, aM, v0, collision_thr
    P.lane_in = dis_ctl - 6 * lane_cw                                                    # TIS:149
    P.lane_len[0] = 3.1415 / 2 * 7 * lane_cw                                             # TIS:149
    P.lane_len[1] = 12 * lane_cw                                                         # TIS:150
    P.lane_len[2] = 3.1415 / 2 * lane_cw                                                 # TIS:151
    P.remove_p = -dis_ctl + int((12 + 1) / 2) * lane_cw                                  # TIS:341-342
    P.cita = (2 * math.sqrt(10) - 6) * lane_cw                                           # TIS:182
    P.alpha = math.atan((6 * lane_cw + P.cita) / (3 * lane_cw))                          # TIS:183
    P.beta = math.pi / 2 - P.alpha                                                       # TIS:184
    P.gama = math.atan((math.sqrt(13) * lane_cw) / (6 * lane_cw))                        # TIS:185
    P.gama2 = math.pi / 2 - P.gama                                                       # TIS:186
    for k in range(4):
        rot = 3.141593 / 2 * k                                                           # TIS:1251
        P.rot_cos[k] = float(np.cos(rot))                                                # TIS:1287
        P.rot_sin[k] = float(np.sin(rot))
    return P


_PTR = C.c_void_p


class _StateView(C.Structure):
    _fields_ = [(n, _PTR) for n in
                ("tick", "lane_n", "veh_rec", "head_lane", "head_j", "id_seq", "passed_veh",
                 "passed_step_total", "p", "v", "a", "jerk_sum", "collision", "step", "seq_in_lane",
                 "uid", "flags", "lock_a", "row0")]


class _Outputs(C.Structure):
    _fields_ = [(n, _PTR) for n in
                ("agent_offset", "ids", "uid", "obs", "reward", "cpv", "status", "jerk_sum", "nn",
                 "collisions", "lock", "n_removed", "q5_undefined")]


STATE_SPEC = {      # name -> (dtype, trailing shape builder)
    "tick": (np.int32, lambda B, c: (B,)), "lane_n": (np.int32, lambda B, c: (B, NLANE)),
    "veh_rec": (np.int32, lambda B, c: (B, NLANE)), "head_lane": (np.int32, lambda B, c: (B, NLANE)),
    "head_j": (np.int32, lambda B, c: (B, NLANE)), "id_seq": (np.int32, lambda B, c: (B,)),
    "passed_veh": (np.int32, lambda B, c: (B,)), "passed_step_total": (np.int64, lambda B, c: (B,)),
    "p": (np.float64, lambda B, c: (B, c)), "v": (np.float64, lambda B, c: (B, c)),
    "a": (np.float64, lambda B, c: (B, c)), "jerk_sum": (np.float64, lambda B, c: (B, c)),
    "collision": (np.int32, lambda B, c: (B, c)), "step": (np.int32, lambda B, c: (B, c)),
    "seq_in_lane": (np.int32, lambda B, c: (B, c)), "uid": (np.int32, lambda B, c: (B, c)),
    "flags": (np.uint8, lambda B, c: (B, c)), "lock_a": (np.int8, lambda B, c: (B, c)),
    "row0": (np.float64, lambda B, c: (B, c, OBS_W)),
}


def empty_state(B, cap):
    st = {k: np.zeros(shape(B, cap), dtype=dt) for k, (dt, shape) in STATE_SPEC.items()}
    st["head_lane"][:] = -1
    st["head_j"][:] = -1
    return st


def valid_rows(arrive):
    """Rows usable per lane: the positive, non-decreasing prefix (zero padding ends the table; equal consecutive times
    are two arrivals, spawned on consecutive ticks as TIS:379 does)."""
    arr = np.asarray(arrive, dtype=np.float64)
    K = arr.shape[-2]
    if K == 0:
        return np.zeros(arr.shape[:-2] + (NLANE,), np.int32)
    bad = arr <= 0
    bad[..., 1:, :] |= arr[..., 1:, :] < arr[..., :-1, :]
    bad = np.logical_or.accumulate(bad, axis=-2)
    return (~bad).sum(axis=-2).astype(np.int32)


class OracleScene:
    """B independent reference-faithful intersections on the CPU."""

    def __init__(self, n_envs, veh_cap, params=None, n_threads=1):
        build()
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p, _PTR, _PTR, C.c_int32, C.c_int32]
        L.orc_set_state.argtypes = [C.c_void_p, C.POINTER(_StateView)]
        L.orc_get_state.argtypes = [C.c_void_p, C.POINTER(_StateView)]
        L.orc_count_agents.restype = C.c_int64
        L.orc_count_agents.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, _PTR, C.POINTER(_Outputs), C.c_int32]
        L.orc_overflow.argtypes = [C.c_void_p]
        L.orc_control_mask.argtypes = [C.c_void_p, _PTR]
        L.orc_control_mask.restype = None
        self.B, self.cap, self.n_threads = int(n_envs), int(veh_cap), int(n_threads)
        self.params = params if params is not None else scene_params()
        self.h = L.orc_create(self.B, self.cap, C.byref(self.params))
        if not self.h:
            raise MemoryError("orc_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def reset(self, arrive, warmup=True):
        arr = np.ascontiguousarray(np.asarray(arrive, dtype=np.float64))
        if arr.ndim == 2:
            arr = np.ascontiguousarray(np.broadcast_to(arr, (self.B,) + arr.shape))
        assert arr.shape[0] == self.B and arr.shape[2] == NLANE, arr.shape
        kv = np.ascontiguousarray(valid_rows(arr))
        rc = self.lib.orc_reset(self.h, arr.ctypes.data, kv.ctypes.data, arr.shape[1], int(bool(warmup)))
        assert rc == 0

    def _view(self, st):
        v = _StateView()
        for k, (dt, shape) in STATE_SPEC.items():
            a = st[k]
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"] and a.shape == shape(self.B, self.cap), (
                k, a.dtype, a.shape)
            setattr(v, k, a.ctypes.data)
        return v

    def set_state(self, st):
        st = {k: np.ascontiguousarray(st[k], dtype=STATE_SPEC[k][0]) for k in STATE_SPEC}
        rc = self.lib.orc_set_state(self.h, C.byref(self._view(st)))
        assert rc == 0, "state does not fit veh_cap"

    def get_state(self):
        st = empty_state(self.B, self.cap)
        self.lib.orc_get_state(self.h, C.byref(self._view(st)))
        return st

    def control_mask(self):
        """bool ``[B, veh_cap]``: ``veh["control"]`` of every slot (main.py:399-405)."""
        m = np.zeros((self.B, self.cap), np.uint8)
        self.lib.orc_control_mask(self.h, m.ctypes.data)
        return m.astype(bool)

    def count_agents(self):
        return int(self.lib.orc_count_agents(self.h))

    def step(self, actions, reuse_buffers=False):
        """``actions``: float32 ``[B, veh_cap]``, one per vehicle slot in (lane, j) order.

        ``reuse_buffers``: write into arrays allocated once for ``B * veh_cap`` rows and return views of them
        (valid until the next call) instead of allocating 1.7 KB of zeros per agent every tick -- used by the
        timed CPU baseline so that it measures the scene, not numpy's allocator."""
        act = np.ascontiguousarray(actions, dtype=np.float32)
        assert act.shape == (self.B, self.cap), act.shape
        A = self.count_agents()
        if reuse_buffers:
            if getattr(self, "_buf", None) is None:
                self._buf = self._alloc_outputs(self.B * self.cap)
            full = self._buf
            o = {k: (a if k in ("agent_offset", "collisions", "lock", "n_removed", "q5_undefined") else a[:A])
                 for k, a in full.items()}
        else:
            full = o = self._alloc_outputs(A)
        ov = _Outputs()
        for k, a in full.items():
            setattr(ov, k, a.ctypes.data)
        self.lib.orc_step(self.h, act.ctypes.data, C.byref(ov), self.n_threads)
        o["overflow"] = int(self.lib.orc_overflow(self.h))
        return o

    def _alloc_outputs(self, A):
        return {
            "agent_offset": np.zeros(self.B + 1, np.int64),
            "ids": np.zeros((A, 2), np.int32), "uid": np.zeros(A, np.int32),
            "obs": np.zeros((A, OBS_H, OBS_W), np.float64), "reward": np.zeros(A, np.float64),
            "cpv": np.zeros(A, np.int32), "status": np.zeros(A, np.uint8),
            "jerk_sum": np.zeros(A, np.float64), "nn": np.zeros((A, NN, 2), np.int32),
            "collisions": np.zeros(self.B, np.int32), "lock": np.zeros(self.B, np.int32),
            "n_removed": np.zeros(self.B, np.int32), "q5_undefined": np.zeros(self.B, np.int32),
        }
