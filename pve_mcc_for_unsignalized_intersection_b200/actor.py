"""Batched MADDPG actor inference on the GPU.

The reference evaluates its policy one vehicle at a time (main.py:398-404, 557-565):

    if veh["control"]:
        agent1_action = get_agents_action(o_n[0], sess, agent1_ddpg_test, noise_range=0)
    env.step(lane, ind, agent1_action[0][0])

``BatchedActor.act(scene)`` produces the whole ``[B, veh_cap]`` action tensor of the next
``BatchedScene.step`` in one kernel launch: the actor of model_agent_maddpg.py:23-49 applied to the
stored observation row 0 of every controlled vehicle, 0 elsewhere.  Weights come from the reference's
own checkpoint (``checkpoint.read_bundle``, no TensorFlow needed) or from an ``.npz`` with the same names.

The kernel is ``csrc/actor.cuh`` behind ``pve_actor_create / pve_actor_forward / pve_act``
(include/pve_mcc.h).  There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from .checkpoint import latest_checkpoint, read_bundle
from .config import OBS_W

# variable names below the actor's scope (tf.layers / tf.contrib.layers naming, NET:23-47), in the order
# of the flat parameter vector of include/pve_mcc.h
PARAM_SPECS = (
    ("LayerNorm/gamma", (28,)), ("LayerNorm/beta", (28,)),
    ("dense/kernel", (28, 64)), ("dense/bias", (64,)),
    ("LayerNorm_1/gamma", (64,)), ("LayerNorm_1/beta", (64,)),
    ("dense_1/kernel", (64, 64)), ("dense_1/bias", (64,)),
    ("LayerNorm_2/gamma", (64,)), ("LayerNorm_2/beta", (64,)),
    ("dense_2/kernel", (64, 1)), ("dense_2/bias", (1,)),
)
ACTOR_FLOATS = sum(int(np.prod(s)) for _, s in PARAM_SPECS)      # 6393


class ActorWeights:
    """The twelve tensors of one actor network, float32, keyed by the names of PARAM_SPECS."""

    def __init__(self, tensors):
        self.tensors = {}
        for name, shape in PARAM_SPECS:
            if name not in tensors:
                raise KeyError("actor tensor %s is missing" % name)
            arr = np.asarray(tensors[name], dtype=np.float32)
            if arr.shape != shape:
                raise ValueError("actor tensor %s has shape %s, expected %s" % (name, arr.shape, shape))
            self.tensors[name] = np.ascontiguousarray(arr)

    @classmethod
    def from_checkpoint(cls, prefix_or_dir, scope="agent1actor"):
        """``prefix_or_dir``: a bundle prefix (``.../66.cptk``) or the directory holding the reference's
        ``checkpoint`` file (main.py:541).  ``scope``: the variable scope of the actor (MADDPG('agent1'))."""
        import os
        prefix = latest_checkpoint(prefix_or_dir) if os.path.isdir(prefix_or_dir) else prefix_or_dir
        names = ["%s/%s" % (scope, n) for n, _ in PARAM_SPECS]
        got = read_bundle(prefix, names)
        return cls({n: got["%s/%s" % (scope, n)] for n, _ in PARAM_SPECS})

    @classmethod
    def from_npz(cls, path):
        with np.load(path) as z:
            return cls({n: z[n.replace("/", "__")] for n, _ in PARAM_SPECS})

    def save_npz(self, path):
        np.savez(path, **{n.replace("/", "__"): self.tensors[n] for n, _ in PARAM_SPECS})

    @classmethod
    def random(cls, seed=0):
        """The reference's initialisation (NET:28-38: kernels U(-3e-3, 3e-3), LN gamma 1 / beta 0, biases 0)."""
        rng = np.random.default_rng(seed)
        t = {}
        for name, shape in PARAM_SPECS:
            if name.endswith("kernel"):
                t[name] = rng.uniform(-3e-3, 3e-3, shape).astype(np.float32)
            elif name.endswith("gamma"):
                t[name] = np.ones(shape, np.float32)
            else:
                t[name] = np.zeros(shape, np.float32)
        return cls(t)

    def flat(self):
        out = np.concatenate([self.tensors[n].reshape(-1) for n, _ in PARAM_SPECS]).astype(np.float32)
        assert out.size == ACTOR_FLOATS
        return out


class BatchedActor:
    """Device copy of one actor network + the two launches that use it."""

    def __init__(self, weights, device="cuda:0", _library=None):
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise N.NativeError("the actor kernel needs a CUDA device (there is no CPU fallback)")
        self.lib = N.load_library(_library)
        self.weights = weights
        flat = weights.flat()
        self._h = C.c_void_p()
        rc = self.lib.pve_actor_create(flat.ctypes.data_as(C.c_void_p), flat.size, self.device.index or 0,
                                       C.byref(self._h))
        if rc != 0:
            raise N.NativeError("pve_actor_create failed with %d" % rc)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pve_actor_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self, rows, out=None):
        """``rows`` float32 ``[n, 28]`` on the device -> actions ``[n]`` (agent.action, NET:120-121)."""
        if rows.device != self.device or rows.dtype != torch.float32 or rows.dim() != 2 or rows.shape[1] != OBS_W:
            raise ValueError("rows must be a float32 [n, 28] tensor on %s" % self.device)
        rows = rows.contiguous()
        if out is None:
            out = torch.empty(rows.shape[0], dtype=torch.float32, device=self.device)
        rc = self.lib.pve_actor_forward(self._h, rows.data_ptr(), rows.shape[0], out.data_ptr(), self._stream())
        if rc != 0:
            raise N.NativeError("pve_actor_forward failed with %d" % rc)
        return out

    def act(self, scene, out=None, noise=None, noise_scale=0.0):
        """The action tensor ``[B, veh_cap]`` for ``scene.step``: the policy on every controlled vehicle's
        stored row 0, 0 for the others (main.py:398-404).  ``noise`` (``[B, veh_cap]`` standard normal
        draws) and ``noise_scale`` reproduce ``+ np.random.randn(1) * noise_range`` of main.py:44."""
        if out is None:
            out = torch.empty(scene.B, scene.veh_cap, dtype=torch.float32, device=self.device)
        if noise is not None and (noise.shape != out.shape or noise.dtype != torch.float32 or noise.device != self.device):
            raise ValueError("noise must be a float32 [B, veh_cap] tensor on %s" % self.device)
        rc = self.lib.pve_act(scene._h, self._h, noise.data_ptr() if noise is not None else None,
                              C.c_float(float(noise_scale)), out.data_ptr(), self._stream())
        scene._check(rc)
        return out

    def rollout(self, scene, n_ticks, out=None, noise=None, noise_scale=0.0):
        """``n_ticks`` x (``act``; ``scene.step``) enqueued by the library (``pve_rollout``): the test drivers' loop
        main.py:553-575 without a Python round trip per tick.  Returns ``scene.out`` (the last tick's outputs)."""
        if out is None:
            out = torch.empty(scene.B, scene.veh_cap, dtype=torch.float32, device=self.device)
        scene.out._n = None
        rc = self.lib.pve_rollout(scene._h, self._h, int(n_ticks), noise.data_ptr() if noise is not None else None,
                                  C.c_float(float(noise_scale)), out.data_ptr(), C.byref(scene._out_native), self._stream())
        scene._check(rc)
        return scene.out
