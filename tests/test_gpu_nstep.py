"""GPU tier of row N2: the CUDA critic, n-step folder and replay writer (csrc/nstep.cuh, through the C ABI)
against the numpy restatement (oracle/nstep_oracle.py) and the trace recorded from the reference's own
main.py:243-266 lines (tests/golden/nstep_mat1000.npz)."""
import os
import random

import numpy as np
import pytest
import torch

import nstep_common as K
import parity as P
from oracle import nstep_oracle
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights, BatchedActor
from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
from pve_mcc_for_unsignalized_intersection_b200.nstep import BatchedCritic, CriticWeights, NStepFolder

pytestmark = pytest.mark.gpu

# fp32 networks: same bound as the actor (tests/test_gpu_actor.py): conditioned at the 1e-4 level
Q_RTOL, Q_ATOL, Q_FRACTION, Q_WORST = 1e-5, 1e-5, 0.97, 5e-4
# n-step targets: float32 rewards (<= 1e-5 relative, the north-star bound) summed over <= 13 terms, plus
# gamma^n * Q' with the network bound above
T_RTOL, T_ATOL = 1e-5, 2e-4


def nets():
    a, c = K.load_nets()
    return ActorWeights(a), CriticWeights(c)


def realistic_obs(rng, n):
    obs = np.zeros((n, 7, 28), dtype=np.float32)
    for c in range(7):
        obs[:, :, 4 * c] = rng.uniform(-20, 180, (n, 7))
        obs[:, :, 4 * c + 1] = rng.uniform(0, 13, (n, 7))
        obs[:, :, 4 * c + 2] = rng.uniform(-3, 3, (n, 7))
        obs[:, :, 4 * c + 3] = rng.randint(0, 12, (n, 7))
    obs[rng.rand(n, 7) < 0.15] = 0                        # missing neighbours are all-zero rows
    return obs


@pytest.fixture(params=["tc5", "mma", "ffma"])
def critic_impl(request, monkeypatch):
    """The three device implementations of the critic: tcgen05 + tensor memory (bf16 x 3 split products, default), the
    same products on mma.sync, and fp32 FFMA."""
    monkeypatch.setenv("PVE_CRITIC_IMPL", request.param)
    return request.param


@pytest.mark.parametrize("n", [1, 33, 127, 128, 129, 5000, 70001])
def test_critic_kernel_against_oracle(n, critic_impl):
    _, cw = nets()
    critic = BatchedCritic(cw)
    rng = np.random.RandomState(n)
    obs = realistic_obs(rng, n)
    a7 = rng.uniform(-3, 3, (n, 7)).astype(np.float32)
    got = critic.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(a7).cuda()).cpu().numpy().astype(np.float64)
    want64 = nstep_oracle.critic_forward(cw, obs[:, 0].astype(np.float64), a7.astype(np.float64), np.float64)
    want32 = nstep_oracle.critic_forward(cw, obs[:, 0], a7, np.float32).astype(np.float64)
    err = np.abs(got - want64)
    ok = err <= Q_RTOL * np.abs(want64) + Q_ATOL
    # (for a handful of rows the fraction is too coarse: allow one row beyond the tight bound)
    assert (ok.mean() >= Q_FRACTION or (~ok).sum() <= 1) and err.max() <= Q_WORST, (ok.mean(), err.max())
    assert err.mean() <= 2.0 * np.abs(want32 - want64).mean() + 1e-7


def new_records(folder, n_before, n_after):
    idx = (torch.arange(n_before, n_after, device="cuda") % folder.capacity)
    return {k: t.index_select(0, idx).cpu().numpy() for k, t in folder.arrays().items()}


def test_recorded_reference_trace():
    """One intersection, the recorded actions: records per tick, their vehicles and targets, and the final deque
    (which wrapped: buffer_size 3000) against what the reference's own lines produced."""
    z = K.load_trace()
    aw, cw = nets()
    B = 3                                                  # three copies; rows of a tick are ordered by intersection
    scene = P.make_scene("cuda", B, vm=int(z["vm"]))
    scene.reset(z["arrive_time"], warmup=True)
    folder = NStepFolder(scene, BatchedActor(aw), BatchedCritic(cw), int(z["seq_max_step"]), buffer_size=3 * 14000)
    gamma = float(z["gamma"])
    ticks = z["n_added"].size
    n_prev, k = 0, 0
    for t in range(ticks):
        out = scene.step(torch.from_numpy(K.dense_actions(z, t, scene.veh_cap, B)).cuda())
        folder.push(out, gamma)
        c = folder.counters()
        want_n = int(z["n_added"][t])
        assert c["last_added"] == B * want_n and c["num_experiences"] == n_prev + B * want_n, (t, c)
        rec = new_records(folder, n_prev, c["num_experiences"])
        want_t = z["rec_target"][k:k + want_n]
        for b in range(B):
            got_t = rec["reward"][b * want_n:(b + 1) * want_n].astype(np.float64)
            assert np.all(np.abs(got_t - want_t) <= T_RTOL * np.abs(want_t) + T_ATOL), (t, b, np.abs(got_t - want_t).max())
            assert np.all(rec["done"] == 0)
            # action column of the stored next_state (TIS:290)
            assert np.array_equal(rec["action"], rec["next_state"][:, :, 2])
        n_prev, k = c["num_experiences"], k + want_n
    assert k == z["rec_target"].size and folder.counters()["slot_conflicts"] == 0
    assert int(scene.get_state()["id_seq"][0]) == int(z["id_seq"])


def test_folder_against_oracle_many_intersections_with_collisions():
    """48 intersections, random actions (collisions, so Done comes both ways), a small memory that wraps: the
    oracle is fed the device's own per-tick outputs, so states must be equal bit for bit."""
    aw, cw = nets()
    B, ticks, S = 48, 260, 12
    scene = P.make_scene("cuda", B, vm=6)
    scene.reset(synthetic_arrivals(B, 1000, 60.0, seed=11), warmup=False)
    folder = NStepFolder(scene, BatchedActor(aw), BatchedCritic(cw), S, buffer_size=scene.out_cap + 1)
    orc = nstep_oracle.NStepOracle(aw, cw, S, buffer_size=scene.out_cap + 1)
    rng = torch.Generator(device="cuda").manual_seed(5)
    gamma, n_prev, n_done = 0.83, 0, 0
    for t in range(ticks):
        acts = (torch.rand(B, scene.veh_cap, device="cuda", generator=rng) * 6 - 3) * scene.control_mask()
        out = scene.step(acts.contiguous())
        folder.push(out, gamma)
        o = P.outputs_to_numpy(out)
        done = (o["status"] & K.ST_DONE) != 0
        n_done += int(done.sum())
        added = orc.push(o["ids"][:, 0], o["ids"][:, 3], o["obs"], o["reward"], done, gamma)
        c = folder.counters()
        assert c["num_experiences"] == n_prev + len(added) == orc.memory.num_experiences, (t, c, len(added))
        rec = new_records(folder, n_prev, c["num_experiences"])
        for i, (row, state, action, target, nxt) in enumerate(added):
            assert np.array_equal(rec["state"][i], state.astype(np.float32)), (t, i)
            assert np.array_equal(rec["next_state"][i], nxt.astype(np.float32)), (t, i)
            assert np.array_equal(rec["action"][i], action.astype(np.float32)), (t, i)
            assert abs(rec["reward"][i] - target) <= T_RTOL * abs(target) + T_ATOL, (t, i, rec["reward"][i], target)
        n_prev = c["num_experiences"]
    assert n_done > 20 and n_prev > folder.capacity and folder.counters()["slot_conflicts"] == 0
    # the deque in the reference's order, and the sampler
    dq = folder.deque()
    assert len(folder) == len(orc.memory.buffer) == folder.capacity
    want_state = np.stack([e[0] for e in orc.memory.buffer]).astype(np.float32)
    want_target = np.array([e[2] for e in orc.memory.buffer])
    assert np.array_equal(dq["state"].cpu().numpy(), want_state)
    assert np.all(np.abs(dq["reward"].cpu().numpy() - want_target) <= T_RTOL * np.abs(want_target) + T_ATOL)
    batch = folder.get_batch(64, random.Random(3))
    pick = random.Random(3).sample(range(len(orc.memory.buffer)), 64)          # replay_buffer.py:20-22
    assert np.array_equal(batch["state"].cpu().numpy(), want_state[pick])


def test_seq_max_step_edges():
    """seq_max_step 0 (every tick emits its own transition) and 14 (largest ring) against the oracle."""
    aw, cw = nets()
    for S in (0, 14):
        B = 8
        scene = P.make_scene("cuda", B, vm=6)
        scene.reset(synthetic_arrivals(B, 1200, 30.0, seed=S), warmup=True)
        folder = NStepFolder(scene, BatchedActor(aw), BatchedCritic(cw), S, buffer_size=200000)
        orc = nstep_oracle.NStepOracle(aw, cw, S, buffer_size=200000)
        acts = torch.zeros(B, scene.veh_cap, device="cuda")
        n_prev = 0
        for t in range(60):
            out = scene.step(acts)
            folder.push(out, 0.5)
            o = P.outputs_to_numpy(out)
            added = orc.push(o["ids"][:, 0], o["ids"][:, 3], o["obs"], o["reward"], (o["status"] & 1) != 0, 0.5)
            c = folder.counters()
            assert c["num_experiences"] == n_prev + len(added), (S, t)
            rec = new_records(folder, n_prev, c["num_experiences"])
            for i, (row, state, action, target, nxt) in enumerate(added):
                assert np.array_equal(rec["state"][i], state.astype(np.float32)), (S, t, i)
                assert np.array_equal(rec["next_state"][i], nxt.astype(np.float32)), (S, t, i)
                assert abs(rec["reward"][i] - target) <= T_RTOL * abs(target) + T_ATOL
            n_prev = c["num_experiences"]
        assert n_prev > 0


def test_full_size_record_count():
    """BASELINE config 2 size (4 096 intersections): the number of records equals an independent count kept with
    torch (age per (intersection, uid)); no slot conflicts; targets finite and within the reward bounds."""
    aw, cw = nets()
    B, ticks, S = 4096, 80, 12
    scene = P.make_scene("cuda", B, vm=6)
    scene.reset(synthetic_arrivals(B, 1000, 40.0, seed=3), warmup=False)
    for _ in range(150):                                   # fill the intersections first
        scene.step((torch.rand(B, scene.veh_cap, device="cuda") * 6 - 3) * scene.control_mask())
    folder = NStepFolder(scene, BatchedActor(aw), BatchedCritic(cw), S, buffer_size=4_000_000)
    UID = 1024
    age = torch.zeros(B * UID, dtype=torch.int32, device="cuda")
    want = 0
    for t in range(ticks):
        out = scene.step((torch.rand(B, scene.veh_cap, device="cuda") * 6 - 3) * scene.control_mask())
        folder.push(out, 0.9)
        n = out.n_agents
        ids = out.ids[:n].long()
        assert int(ids[:, 3].max()) < UID
        key = ids[:, 0] * UID + ids[:, 3]
        age[key] += 1
        emit = out.done[:n] | (age[key] > S)
        age[key] -= emit.int()
        want += int(emit.sum())
    c = folder.counters()
    assert c["num_experiences"] == want and c["slot_conflicts"] == 0 and want > 1_000_000
    r = folder.arrays()["reward"][:want]
    assert bool(torch.isfinite(r).all()) and float(r.abs().max()) < 20.0 / (1 - 0.9) + 1


def test_folder_rejects_bad_arguments():
    aw, cw = nets()
    scene = P.make_scene("cuda", 2)
    actor, critic = BatchedActor(aw), BatchedCritic(cw)
    from pve_mcc_for_unsignalized_intersection_b200 import _native as N
    with pytest.raises(N.NativeError):
        NStepFolder(scene, actor, critic, seq_max_step=15)
    with pytest.raises(N.NativeError):
        NStepFolder(scene, actor, critic, uid_slots=100)
    with pytest.raises(N.NativeError):
        NStepFolder(scene, actor, critic, buffer_size=scene.out_cap)       # deque of buffer_size - 1 < one tick


def test_new_episode_keeps_the_memory_and_drops_the_buffers():
    """Two episodes back to back on the same tables: uids repeat, so without ``reset`` the second episode would
    continue the first one's histories."""
    aw, cw = nets()
    B, S = 6, 5
    tabs = synthetic_arrivals(B, 1000, 30.0, seed=21)
    scene = P.make_scene("cuda", B, vm=6)
    folder = NStepFolder(scene, BatchedActor(aw), BatchedCritic(cw), S, buffer_size=100000)
    orc = nstep_oracle.NStepOracle(aw, cw, S, buffer_size=100000)
    acts = torch.zeros(B, scene.veh_cap, device="cuda")
    n_prev = 0
    for episode in range(2):
        scene.reset(tabs, warmup=True)
        folder.reset()
        orc.reset()
        for t in range(70):
            out = scene.step(acts)
            folder.push(out, 0.7)
            o = P.outputs_to_numpy(out)
            added = orc.push(o["ids"][:, 0], o["ids"][:, 3], o["obs"], o["reward"], (o["status"] & 1) != 0, 0.7)
            c = folder.counters()
            assert c["num_experiences"] == n_prev + len(added), (episode, t)
            rec = new_records(folder, n_prev, c["num_experiences"])
            for i, (row, state, action, target, nxt) in enumerate(added):
                assert np.array_equal(rec["state"][i], state.astype(np.float32)), (episode, t, i)
                assert np.array_equal(rec["next_state"][i], nxt.astype(np.float32)), (episode, t, i)
                assert abs(rec["reward"][i] - target) <= T_RTOL * abs(target) + T_ATOL
            n_prev = c["num_experiences"]
    assert n_prev > 1000 and folder.counters()["slot_conflicts"] == 0


@pytest.mark.parametrize("impl", ["tc5", "mma", "ffma"])
def test_distinct_row_bootstrap_equals_the_seven_row_bootstrap(impl, monkeypatch):
    monkeypatch.setenv("PVE_ACTOR_IMPL", impl)             # both device implementations of the two networks
    monkeypatch.setenv("PVE_CRITIC_IMPL", impl)
    _distinct_row_bootstrap()


def _distinct_row_bootstrap():
    """pve_nstep_push_scene (target actor once per distinct row, actions gathered through nbr_src) against pve_nstep_push
    (target actor on all 7 rows of every observation): bootstrap values and replay memory bit for bit."""
    aw, cw = nets()
    B, S = 512, 12
    scene = P.make_scene("cuda", B, vm=6, neighbour_sources=True)
    scene.reset(synthetic_arrivals(B, 1000, 40.0, seed=8), warmup=False)
    actor, critic = BatchedActor(aw), BatchedCritic(cw)
    plain = NStepFolder(scene, actor, critic, S, buffer_size=3_000_000)
    fast = NStepFolder(scene, actor, critic, S, buffer_size=3_000_000)
    gen = torch.Generator(device="cuda").manual_seed(2)
    for t in range(220):
        acts = (torch.rand(B, scene.veh_cap, device="cuda", generator=gen) * 6 - 3) * scene.control_mask()
        out = scene.step(acts.contiguous())
        plain.push(out, 0.85, distinct_rows=False)
        fast.push(out, 0.85)
        if t % 20 == 19:
            n = out.n_agents
            assert torch.equal(plain.bootstrap_values()[:n], fast.bootstrap_values()[:n]), t
    c0, c1 = plain.counters(), fast.counters()
    assert c0["num_experiences"] == c1["num_experiences"] > 1_000_000 and c1["slot_conflicts"] == 0
    n = c0["num_experiences"]
    for k in ("state", "action", "reward", "next_state"):
        assert torch.equal(plain.arrays()[k][:n], fast.arrays()[k][:n]), k


def test_frame_log_in_place_equals_the_copied_observations():
    """pve_nstep_obs_slot: a step that writes its observations straight into the folder's frame log (bind_outputs) and
    a push that copies them in must leave the same replay memory, bit for bit; the log wraps many times."""
    aw, cw = nets()
    B, S = 256, 5
    scenes, folders = [], []
    for _ in range(2):
        scene = P.make_scene("cuda", B, vm=6, neighbour_sources=True)
        scene.reset(synthetic_arrivals(B, 1000, 40.0, seed=11), warmup=False)
        scenes.append(scene)
        folders.append(NStepFolder(scene, BatchedActor(aw), BatchedCritic(cw), S, buffer_size=1_000_000))
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(150):
        acts = ((torch.rand(B, scenes[0].veh_cap, device="cuda", generator=gen) * 6 - 3) * scenes[0].control_mask()).contiguous()
        if t == 70:
            for f in folders:
                f.reset()                                            # a new episode: buffers dropped, memory kept
        folders[0].bind_outputs()
        out0 = scenes[0].step(acts)
        out1 = scenes[1].step(acts)
        n = out0.n_agents
        assert n == out1.n_agents and torch.equal(out0.obs[:n], out1.obs[:n])
        folders[0].push(out0, 0.9)
        folders[1].push(out1, 0.9)
    c0, c1 = folders[0].counters(), folders[1].counters()
    assert c0["num_experiences"] == c1["num_experiences"] > 100_000 and c0["slot_conflicts"] == 0
    n = c0["num_experiences"]
    for k in ("state", "action", "reward", "next_state"):
        assert torch.equal(folders[0].arrays()[k][:n], folders[1].arrays()[k][:n]), k
