"""Shared pieces of the n-step / replay tests (row N2): the recorded reference trace and its replay through the
CPU scene oracle."""
import hashlib
import os

import numpy as np

import parity  # noqa: F401  (sys.path)
from oracle import nstep_oracle
from pve_mcc_for_unsignalized_intersection_b200.actor import PARAM_SPECS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ST_DONE = 1


def load_nets():
    with np.load(os.path.join(GOLD, "nstep_nets.npz")) as z:
        actor = {n: z["actor__" + n.replace("/", "__")] for n, _ in PARAM_SPECS}
        critic = {n: z["critic__" + n.replace("/", "__")] for n, _ in nstep_oracle.CRITIC_SPECS}
    return actor, critic


def load_trace():
    return np.load(os.path.join(GOLD, "nstep_mat1000.npz"))


def digest(state, action, next_state):
    """tests/golden/make_nstep_golden.py::digest"""
    h = hashlib.blake2b(digest_size=16)
    for a in (state, action, next_state):
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def dense_actions(z, tick, cap, copies=1):
    """The recorded per-vehicle actions of one tick as the ``[copies, cap]`` tensor of the batched scene."""
    a = z["actions"][z["action_offset"][tick]:z["action_offset"][tick + 1]]
    out = np.zeros((copies, cap), dtype=np.float32)
    out[:, :a.size] = a
    return out
