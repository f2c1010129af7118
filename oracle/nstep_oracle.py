"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs) -- never imported by
the product path.

CPU restatement of the reference's n-step return folding and replay writer:

* main.py:243-266 (MAIN) -- after every ``scene_update`` each agent row ``seq`` appends
  ``[state_now, actions, reward, state_next, Done]`` to its vehicle's ``buffer``; when the vehicle is
  ``Done`` or its ``count`` exceeds ``seq_max_step`` the buffered rewards are folded into one n-step
  target and the OLDEST transition is added to the replay memory with that target.
* replay_buffer.py:45-53 (RB) -- ``ReplayBuffer.add`` with ``rand_s=True`` (main.py:212): a deque that
  never holds more than ``buffer_size - 1`` items (the counter is incremented before the comparison).
* model_agent_maddpg.py:52-76 (NET) -- the critic used for the bootstrap term
  (``agent1_ddpg_target.Q``), restated in numpy like oracle/actor_oracle.py restates the actor.

Facts of the reference this restatement relies on (each checked by the golden trace, see below):
  * ``veh["count"]`` (TIS:292) and ``len(veh["buffer"])`` move together, so ``count > seq_max_step`` is
    ``len(buffer) > seq_max_step`` after the append.
  * ``state_now[seq]`` (main.py:235-240) is the vehicle's stored 7 x 28 state: ``state_next`` of its
    previous tick (TIS:288) or zeros for a vehicle that has not been an agent yet (TIS:380, 420).
  * ``actions[seq]`` is column 2 of ``state_next[seq]`` (TIS:290).
  * a ``Done`` vehicle is never an agent again (TIS:336, 353), so only the oldest transition of its
    buffer ever reaches the replay memory.

PARITY PINNING: tests/golden/make_nstep_golden.py executes the reference's OWN lines main.py:243-266 (read from
/root/reference at generation time, not copied) around the unmodified reference scene and the reference's own
``ReplayBuffer``, with the two TensorFlow networks replaced by the numpy restatements (TensorFlow 1.12 is not
installable here: single activations of the networks stay "parity unpinned", exactly as for the actor row).
tests/test_nstep_oracle.py replays this module against that trace: every replay record (hash of the float64
state / action / next-state arrays, n-step target) and the final deque.
"""
from collections import deque

import numpy as np

try:
    from .actor_oracle import actor_forward, layer_norm
except ImportError:                      # oracle/ itself on sys.path (tests/golden/make_nstep_golden.py)
    from actor_oracle import actor_forward, layer_norm

CRITIC_SPECS = (
    ("LayerNorm/gamma", (28,)), ("LayerNorm/beta", (28,)),
    ("dense/kernel", (28, 64)), ("dense/bias", (64,)),
    ("LayerNorm_1/gamma", (64,)), ("LayerNorm_1/beta", (64,)),
    ("dense_1/kernel", (71, 64)), ("dense_1/bias", (64,)),
    ("LayerNorm_2/gamma", (64,)), ("LayerNorm_2/beta", (64,)),
    ("dense_2/kernel", (64, 1)), ("dense_2/bias", (1,)),
)


def critic_forward(weights, state_rows, actions7, dtype=np.float32):
    """NET:52-76.  ``state_rows`` [n, 28] (row 0 of the observation), ``actions7`` [n, 7] =
    concat(action_input, other_action_input) (NET:81-83)  ->  Q [n]."""
    t = weights.tensors if hasattr(weights, "tensors") else weights
    x = np.asarray(state_rows, dtype=dtype).reshape(-1, 28)
    a = np.asarray(actions7, dtype=dtype).reshape(-1, 7)
    x = layer_norm(x, t["LayerNorm/gamma"], t["LayerNorm/beta"], dtype)                          # NET:58-59
    x = x @ t["dense/kernel"].astype(dtype) + t["dense/bias"].astype(dtype)                      # NET:60-61
    x = np.maximum(layer_norm(x, t["LayerNorm_1/gamma"], t["LayerNorm_1/beta"], dtype), 0)       # NET:62-64
    x = np.concatenate([x, a], axis=1)                                                           # NET:66
    x = x @ t["dense_1/kernel"].astype(dtype) + t["dense_1/bias"].astype(dtype)                  # NET:67-68
    x = np.maximum(layer_norm(x, t["LayerNorm_2/gamma"], t["LayerNorm_2/beta"], dtype), 0)       # NET:69-71
    x = x @ t["dense_2/kernel"].astype(dtype) + t["dense_2/bias"].astype(dtype)                  # NET:73
    return x.reshape(-1)


def bootstrap_q(target_actor, target_critic, state_next, dtype=np.float32):
    """main.py:253-260 for a batch: Q'(s'[0], mu'(s'[0]), [mu'(s'[1]) ... mu'(s'[6])])."""
    s = np.asarray(state_next, dtype=np.float64).reshape(-1, 7, 28)
    if s.shape[0] == 0:
        return np.zeros(0, dtype=dtype)
    acts = actor_forward(target_actor, s.reshape(-1, 28), dtype).reshape(-1, 7)
    return critic_forward(target_critic, s[:, 0, :], acts, dtype)


class ReplayDeque:
    """ReplayBuffer(rand_s=True), RB:8-9, 45-53."""

    def __init__(self, buffer_size):
        self.buffer_size = int(buffer_size)
        self.num_experiences = 0
        self.buffer = deque()

    def add(self, state, action, reward, next_state, done):
        self.num_experiences += 1                                   # RB:47
        if self.num_experiences < self.buffer_size:                 # RB:49
            self.buffer.append((state, action, reward, next_state, done))
        else:
            self.buffer.popleft()                                   # RB:52
            self.buffer.append((state, action, reward, next_state, done))


class NStepOracle:
    """The per-vehicle buffers of main.py:243-266 for any number of intersections.  Vehicles are keyed by
    ``(intersection, uid)`` (``id_info[0]``, TIS:426), which is what ``env.veh_info[lane][j]`` resolves to."""

    def __init__(self, target_actor, target_critic, seq_max_step=12, buffer_size=500000, dtype=np.float32):
        self.target_actor, self.target_critic = target_actor, target_critic
        self.seq_max_step = int(seq_max_step)                       # main.py:91, 224
        self.memory = ReplayDeque(buffer_size)                      # main.py:212
        self.dtype = dtype
        self.buffers = {}
        self.stored_state = {}

    def reset(self):
        """A new episode: main.py:230 builds a new scene, so every vehicle record (and its buffer) is gone; the
        replay memory of main.py:212 lives on."""
        self.buffers = {}
        self.stored_state = {}

    def push(self, env_of_row, uid, state_next, reward, done, gamma):
        """One tick.  Rows in the order of ``ids`` (intersection, lane, j ascending).  Returns the records added to
        the replay memory this tick as ``(row, state, action, r_target, next_state)``."""
        state_next = np.asarray(state_next, dtype=np.float64).reshape(-1, 7, 28)
        A = state_next.shape[0]
        plan = []
        for r in range(A):
            key = (int(env_of_row[r]), int(uid[r]))
            buf = self.buffers.setdefault(key, [])
            state_now = self.stored_state.get(key)
            if state_now is None:
                state_now = np.zeros((7, 28))                       # TIS:380
            buf.append([state_now, state_next[r][:, 2].copy(), float(reward[r]), state_next[r], bool(done[r])])  # main.py:244-246
            self.stored_state[key] = state_next[r]                  # TIS:288
            if done[r] or len(buf) > self.seq_max_step:             # main.py:247-248
                plan.append((r, key))
        need_q = [r for r, key in plan if not done[r]]
        q = dict(zip(need_q, bootstrap_q(self.target_actor, self.target_critic, state_next[need_q], self.dtype)))
        added = []
        for r, key in plan:
            buf = self.buffers[key]
            if done[r]:
                r_target = buf[-1][2]                               # main.py:250-251
            else:
                r_target = buf[-1][2] + gamma * q[r]                # main.py:256-260 (np.float32 Q, float64 sum)
            for cur in reversed(buf[:-1]):                          # main.py:261-262
                r_target = cur[2] + gamma * r_target
            first = buf[0]
            self.memory.add(np.array(first[0]), np.array(first[1]), r_target, np.array(first[3]), False)  # main.py:263-264
            added.append((r, first[0], first[1], float(r_target), first[3]))
            buf.pop(0)                                              # main.py:265-266
            if done[r]:
                del self.buffers[key]
                del self.stored_state[key]
        return added
