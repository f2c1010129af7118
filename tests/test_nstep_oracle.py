"""CPU tier of row N2: the restatement of main.py:243-266 + replay_buffer.py:45-53 (oracle/nstep_oracle.py)
against the trace recorded from the reference's own lines (tests/golden/make_nstep_golden.py)."""
import numpy as np

import nstep_common as K
import parity
from oracle import nstep_oracle

TOL = 2e-5


def test_trace_shape():
    z = K.load_trace()
    assert int(z["num_experiences"]) == int(z["n_added"].sum()) == z["rec_target"].size == 13913
    # replay_buffer.py:47-53: the deque never reaches buffer_size
    assert z["final_target"].size == int(z["buffer_size"]) - 1
    assert abs(float(z["gamma"]) - np.tanh(26 / 12.0) * 0.9) < 1e-15


def test_oracle_reproduces_the_reference_lines():
    z = K.load_trace()
    actor, critic = K.load_nets()
    ticks = z["n_added"].size
    scene = parity.make_oracle(1, vm=int(z["vm"]), veh_cap=128, n_threads=1)
    scene.reset(z["arrive_time"][None], warmup=True)
    fold = nstep_oracle.NStepOracle(actor, critic, int(z["seq_max_step"]), int(z["buffer_size"]))
    gamma = float(z["gamma"])
    k = 0
    for t in range(ticks):
        o = scene.step(K.dense_actions(z, t, 128))
        A = int(o["agent_offset"][1])
        added = fold.push(np.zeros(A, np.int64), o["uid"], o["obs"], o["reward"], (o["status"] & K.ST_DONE) != 0, gamma)
        assert len(added) == int(z["n_added"][t]), t
        for row, state, action, target, nxt in added:
            assert int(o["uid"][row]) == int(z["rec_uid"][k]), (t, k)
            assert np.array_equal(K.digest(state, action, nxt), z["rec_digest"][k]), (t, k)
            # the bootstrap is evaluated in one numpy batch here and row by row in the trace: the fp32 networks
            # are conditioned at the 1e-4 level (tests/test_actor_oracle.py), and Q enters times gamma^13 = 0.18
            assert abs(target - z["rec_target"][k]) <= TOL * max(1.0, abs(z["rec_target"][k])), (t, k)
            k += 1
    assert k == z["rec_target"].size and fold.memory.num_experiences == int(z["num_experiences"])
    assert len(fold.memory.buffer) == z["final_target"].size
    for e, want_t, want_d in zip(fold.memory.buffer, z["final_target"], z["final_digest"]):
        assert np.array_equal(K.digest(e[0], e[1], e[3]), want_d)
        assert abs(e[2] - want_t) <= TOL * max(1.0, abs(want_t)) and e[4] is False


def test_done_uses_the_plain_reward_and_frees_the_vehicle():
    actor, critic = K.load_nets()
    fold = nstep_oracle.NStepOracle(actor, critic, seq_max_step=3, buffer_size=10)
    rng = np.random.RandomState(1)
    obs = rng.randn(5, 1, 7, 28)
    rew = [1.0, 2.0, 4.0, 8.0, 16.0]
    g = 0.5
    out = []
    for t in range(5):
        out.append(fold.push([0], [7], obs[t], [rew[t]], [t == 4], g))
    assert [len(a) for a in out] == [0, 0, 0, 1, 1]
    q = nstep_oracle.bootstrap_q(actor, critic, obs[3])[0]
    assert abs(out[3][0][3] - (1 + g * (2 + g * (4 + g * (8 + g * q))))) < 1e-9       # 4 buffered rewards + bootstrap
    assert np.all(out[3][0][1] == 0) and np.array_equal(out[3][0][4], obs[0, 0])       # zeros -> first observation
    assert abs(out[4][0][3] - (2 + g * (4 + g * (8 + g * 16)))) < 1e-12                # Done: no bootstrap
    assert np.array_equal(out[4][0][1], obs[0, 0]) and np.array_equal(out[4][0][4], obs[1, 0])
    assert (0, 7) not in fold.buffers


def test_critic_weight_layout_and_host_errors():
    import pytest
    import torch
    from pve_mcc_for_unsignalized_intersection_b200 import _native
    from pve_mcc_for_unsignalized_intersection_b200.nstep import (CRITIC_FLOATS, CRITIC_SPECS, BatchedCritic,
                                                                   CriticWeights)
    assert CRITIC_SPECS == nstep_oracle.CRITIC_SPECS and CRITIC_FLOATS == 6841
    _, critic = K.load_nets()
    w = CriticWeights(critic)
    flat = w.flat()
    # order of the flat vector = order of include/pve_mcc.h (PVE_CRITIC_FLOATS)
    assert np.array_equal(flat[:28], critic["LayerNorm/gamma"])
    off_w2 = 56 + 28 * 64 + 64 + 128
    assert np.array_equal(flat[off_w2:off_w2 + 71 * 64].reshape(71, 64), critic["dense_1/kernel"])
    assert flat[-1] == critic["dense_2/bias"][0]
    with pytest.raises(ValueError):
        CriticWeights({n: np.zeros((2,)) for n, _ in CRITIC_SPECS})
    header = open(__import__("os").path.join(parity.ROOT, "include", "pve_mcc.h")).read()
    assert "#define PVE_CRITIC_FLOATS 6841" in header
    if not torch.cuda.is_available():
        with pytest.raises(_native.NativeError):           # no CPU fallback
            BatchedCritic(w)
    # a random critic in the reference's initialisation has the right shapes
    q = nstep_oracle.critic_forward(CriticWeights.random(1), np.zeros((3, 28)), np.zeros((3, 7)))
    assert q.shape == (3,)
