"""Mint the fixture of the n-step / replay row.  Build container only: needs
/root/reference.  Writes

* tests/golden/nstep_nets.npz      -- ``agent1_targetactor`` and ``agent1_target_critic`` of the shipped checkpoint
* tests/golden/nstep_mat1000.npz   -- the UNMODIFIED reference scene on arvTimeNewVeh_new_1000_12.mat (vm = 6 as in
                                      training, main.py:230) driven like the training loop main.py:231-241 (policy +
                                      0.2 * N(0,1) noise), with the reference's OWN lines main.py:243-266 executed after
                                      every scene_update: they are read from /root/reference/main.py when this script
                                      runs (never copied into the repo) and compiled into a function.  The replay memory
                                      is the reference's own ``ReplayBuffer(rand_s=True)`` (small, so that the deque
                                      wraps).  ``sess`` / ``agent1_ddpg_target`` are numpy stand-ins for the TensorFlow
                                      networks (oracle/actor_oracle.py, oracle/nstep_oracle.py::critic_forward).

Recorded per tick: the float32 action of every vehicle (the inputs), the number of records added, and per record
the uid of the vehicle, the n-step target and a 16-byte BLAKE2b digest of the float64 state / action / next-state
arrays;
at the end the digests and targets of the deque in order.
"""
import argparse
import hashlib
import os
import sys
import textwrap

import numpy as np
import scipy.io as scio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import actor_oracle  # noqa: E402
import nstep_oracle  # noqa: E402
import ref_harness as H  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.actor import PARAM_SPECS  # noqa: E402
from pve_mcc_for_unsignalized_intersection_b200.checkpoint import read_bundle  # noqa: E402

TICKS = 400
EPOCH = 20                       # main.py:227: gamma = tanh((epoch + 6) / 12) * 0.9
BUFFER_SIZE = 3000               # small enough for the deque to wrap within TICKS
MAT = "/root/reference/data/test/arvTimeNewVeh_new_1000_12.mat"
CKPT = "/root/reference/model_data/baseline/66.cptk"
MAIN = "/root/reference/main.py"


def digest(state, action, next_state):
    """16-byte BLAKE2b over the float64 bytes of the three arrays of one replay record."""
    h = hashlib.blake2b(digest_size=16)
    for a in (state, action, next_state):
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
    return h.digest()


def reference_fold_function():
    """main.py:243-266 as a function of the names those lines use."""
    src = open(MAIN, encoding="utf-8").read().split("\n")
    assert src[242].strip() == "for seq, car_index in enumerate(ids):", src[242]
    assert src[265].strip().endswith('["count"] -= 1'), src[265]
    body = textwrap.dedent("\n".join(src[242:266]))
    code = ("def fold(ids, env, state_now, actions, reward, state_next, seq_max_step, args, agent1_ddpg_target, sess, "
            "agent1_memory_seq, np):\n" + textwrap.indent(body, "    "))
    ns = {}
    exec(compile(code, MAIN, "exec"), ns)
    return ns["fold"]


class TargetNets:
    """agent1_ddpg_target of main.py:201 with the two sess.run calls of NET:120-125 in numpy."""

    def __init__(self, actor, critic):
        self.actor, self.critic = actor, critic

    def action(self, state, sess):
        return actor_oracle.actor_forward(self.actor, np.asarray(state, dtype=np.float64), np.float32).reshape(-1, 1)

    def Q(self, state, action, other_action, sess):
        a7 = np.concatenate([np.asarray(action, dtype=np.float32).reshape(-1, 1),
                             np.asarray(other_action, dtype=np.float32).reshape(-1, 6)], axis=1)      # NET:81-83
        return nstep_oracle.critic_forward(self.critic, np.asarray(state, dtype=np.float64), a7, np.float32).reshape(-1, 1)


def main():
    from replay_buffer import ReplayBuffer                      # the reference's own class
    names = lambda scope, specs: ["%s/%s" % (scope, n) for n, _ in specs]
    got = read_bundle(CKPT, names("agent1actor", PARAM_SPECS) + names("agent1_targetactor", PARAM_SPECS)
                      + names("agent1_target_critic", nstep_oracle.CRITIC_SPECS))
    online = {n: got["agent1actor/" + n] for n, _ in PARAM_SPECS}
    t_actor = {n: got["agent1_targetactor/" + n] for n, _ in PARAM_SPECS}
    t_critic = {n: got["agent1_target_critic/" + n] for n, _ in nstep_oracle.CRITIC_SPECS}
    np.savez(os.path.join(HERE, "nstep_nets.npz"),
             **{"actor__" + n.replace("/", "__"): v for n, v in t_actor.items()},
             **{"critic__" + n.replace("/", "__"): v for n, v in t_critic.items()})

    fold = reference_fold_function()
    arr = scio.loadmat(MAT)["arvTimeNewVeh"]
    env = H.RefEnv(H.load_reference(), arr, vm=6).env
    args = argparse.Namespace(gamma=np.tanh(float(EPOCH + 6) / 12.0) * 0.90, o_agent_num=6)
    memory = ReplayBuffer(BUFFER_SIZE, 128, 1000, 50000, rand_s=True)
    nets = TargetNets(t_actor, t_critic)
    rng = np.random.RandomState(20261017)

    act_log, act_off = [], [0]
    n_added, rec_uid, rec_target, rec_digest = [], [], [], []
    uid_of_obj = {}
    orig_add = memory.add

    def logging_add(state, action, reward, next_state, done):
        assert done is False
        rec_target.append(float(reward))
        rec_digest.append(digest(state, action, next_state))
        orig_add(state, action, reward, next_state, done)

    memory.add = logging_add
    for i in range(TICKS):
        state_now = []
        for lane in range(12):
            for ind, veh in enumerate(env.veh_info[lane]):
                a = 0.0
                if veh["control"]:
                    o_n = veh["state"]
                    a = actor_oracle.actor_forward(online, np.asarray(o_n[0], dtype=np.float64)[None, :], np.float32)[0] \
                        + rng.randn() * 0.2                                              # main.py:44, 239
                    state_now.append(o_n)
                a = float(np.float32(a))                 # the batched scene takes float32 actions
                act_log.append(a)
                env.step(lane, ind, a)
        act_off.append(len(act_log))
        ids, state_next, reward, actions, _, _, cpv, jerks, lock = env.scene_update()
        before = len(rec_target)
        uids = [env.veh_info[l][j]["id_info"][0] for l, j in ids]
        counts_before = [len(env.veh_info[l][j]["buffer"]) for l, j in ids]
        dones = [env.veh_info[l][j]["Done"] for l, j in ids]
        fold(ids, env, state_now, actions, reward, state_next, 12, args, nets, None, memory, np)
        for k, (l, j) in enumerate(ids):                 # which rows added a record (main.py:247-248)
            if dones[k] or counts_before[k] + 1 > 12:
                rec_uid.append(uids[k])
        n_added.append(len(rec_target) - before)
        assert len(rec_uid) == len(rec_target)
        env.delete_vehicle()
    final_digest = [digest(e[0], e[1], e[3]) for e in memory.buffer]
    final_target = [float(e[2]) for e in memory.buffer]
    keep = int(np.max(np.sum((arr > 0) & (arr < TICKS * 0.1 + 20.0), axis=0))) + 2
    np.savez_compressed(
        os.path.join(HERE, "nstep_mat1000.npz"), arrive_time=arr[:keep].astype(np.float64),
        gamma=np.float64(args.gamma), seq_max_step=np.int64(12), buffer_size=np.int64(BUFFER_SIZE), vm=np.int64(6),
        actions=np.asarray(act_log, dtype=np.float32), action_offset=np.asarray(act_off, dtype=np.int64),
        n_added=np.asarray(n_added, dtype=np.int64), rec_uid=np.asarray(rec_uid, dtype=np.int64),
        rec_target=np.asarray(rec_target, dtype=np.float64),
        rec_digest=np.frombuffer(b"".join(rec_digest), dtype=np.uint8).reshape(-1, 16),
        final_target=np.asarray(final_target, dtype=np.float64),
        final_digest=np.frombuffer(b"".join(final_digest), dtype=np.uint8).reshape(-1, 16),
        num_experiences=np.int64(memory.num_experiences), id_seq=np.int64(env.id_seq))
    print("ticks", TICKS, "records", len(rec_target), "deque", len(memory.buffer), "num_experiences",
          memory.num_experiences, "vehicles", env.id_seq, "gamma", float(args.gamma))


if __name__ == "__main__":
    main()
