"""GPU tier for the actor row (N1): the CUDA actor (csrc/actor.cuh, through the C ABI) against the numpy
oracle, and the closed loop actor + environment step against config 1 of the reference."""
import os

import numpy as np
import pytest
import torch

import parity as P
from oracle import actor_oracle
from pve_mcc_for_unsignalized_intersection_b200 import _native as N
from pve_mcc_for_unsignalized_intersection_b200.actor import ActorWeights, BatchedActor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# The network is fp32 in the reference (tf.float32 placeholders, NET:15).  It is ill-conditioned at the 1e-4
# level: rounding only the *inputs* (positions ~160 m) to fp32 moves some actions by 9e-5, and numpy's own
# fp32 evaluation deviates from the float64 one by up to 1.2e-4 on the recorded rows.  So the kernel is held
# to: 1e-5 relative (+1e-5 absolute near 0) on at least 97 % of the rows, never more than 5e-4 absolute,
# and on average no farther from the float64 evaluation than numpy-fp32 is.
RTOL, ATOL, FRACTION, WORST = 1e-5, 1e-5, 0.97, 5e-4


def check_actions(got, rows32, w, what):
    want32 = actor_oracle.actor_forward(w, rows32, np.float32)
    want64 = actor_oracle.actor_forward(w, rows32.astype(np.float64), np.float64)
    err = np.abs(got.astype(np.float64) - want64)
    ok = err <= RTOL * np.abs(want64) + ATOL
    assert ok.mean() >= FRACTION, "%s: only %.2f%% of the actions within tolerance" % (what, 100 * ok.mean())
    assert err.max() <= WORST, "%s: worst action error %g" % (what, err.max())
    err_np = np.abs(want32.astype(np.float64) - want64)
    assert err.mean() <= 2.0 * err_np.mean() + 1e-7, (what, err.mean(), err_np.mean())
    assert np.all(np.abs(got) <= 3.0)


@pytest.fixture(params=["tc5", "mma", "ffma"])
def impl(request, monkeypatch):
    """The three device implementations of the actor: tcgen05 + tensor memory (bf16 x 3 split products, default), the same
    products on mma.sync, and fp32 FFMA."""
    monkeypatch.setenv("PVE_ACTOR_IMPL", request.param)
    return request.param


def test_actor_kernel_on_recorded_rows(impl):
    z = np.load(os.path.join(GOLD, "actor_rollout_mat1000.npz"))
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    actor = BatchedActor(w)
    rows32 = z["rows"].astype(np.float32)
    got = actor.forward(torch.from_numpy(rows32).cuda()).cpu().numpy()
    check_actions(got, rows32, w, "recorded rows")


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 70001])
def test_actor_kernel_shapes_and_edge_rows(n, impl):
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    actor = BatchedActor(w)
    rng = np.random.RandomState(n)
    rows = np.zeros((n, 28), dtype=np.float32)
    for c in range(7):                                   # (p, v, a, lane) blocks like a real observation row
        rows[:, 4 * c] = rng.uniform(-20, 180, n)
        rows[:, 4 * c + 1] = rng.uniform(5, 13, n)
        rows[:, 4 * c + 2] = rng.uniform(-3, 3, n)
        rows[:, 4 * c + 3] = rng.randint(0, 12, n)
    rows[::7, 4:] = 0                                    # vehicles with no neighbours (TIS:1334)
    rows[0] = 0                                          # an all-zero row: variance 0 in the first LayerNorm
    out = torch.full((n + 5,), 7.0, device="cuda")
    got = actor.forward(torch.from_numpy(rows).cuda(), out=out[:n]).cpu().numpy()
    assert torch.all(out[n:] == 7.0)                     # nothing written past the end
    check_actions(got, rows, w, "n=%d" % n)


def test_random_initialised_actor(impl):
    w = ActorWeights.random(5)
    actor = BatchedActor(w)
    rows = (np.random.RandomState(1).randn(500, 28) * 30).astype(np.float32)
    got = actor.forward(torch.from_numpy(rows).cuda()).cpu().numpy()
    check_actions(got, rows, w, "random init")


def test_act_on_scene_matches_forward_and_masks(impl):
    """pve_act: the policy on the stored row 0 of controlled vehicles, 0 elsewhere (main.py:398-404)."""
    from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    B = 24
    scene = P.make_scene("cuda", B, vm=5)
    actor = BatchedActor(w)
    scene.reset(synthetic_arrivals(B, 1000, 60.0, seed=4, rows=40), warmup=True)
    rng = np.random.RandomState(0)
    for t in range(150):
        acts = actor.act(scene) if t % 2 else (torch.rand(B, scene.veh_cap, device="cuda") * 6 - 3)
        scene.step(acts)
    mask = scene.control_mask()
    rows = scene.row0()
    acts = actor.act(scene)
    assert torch.all(acts[~mask] == 0)
    dense = actor.forward(rows[mask].contiguous())
    assert torch.equal(acts[mask], dense)                # same kernel arithmetic on both entry points
    check_actions(dense.cpu().numpy(), rows[mask].cpu().numpy(), w, "scene rows")
    assert int(mask.sum()) > 400
    noise = torch.randn(B, scene.veh_cap, device="cuda")
    noisy = actor.act(scene, noise=noise, noise_scale=0.25)
    torch.testing.assert_close(noisy[mask], acts[mask] + 0.25 * noise[mask], rtol=0, atol=1e-6)
    assert torch.all(noisy[~mask] == 0)


@pytest.mark.parametrize("caps,B", [((192, 128), 300), ((384, 320), 160), ((576, 416), 120)])
def test_act_in_the_large_capacity_classes(caps, B, monkeypatch):
    """pve_act of the tcgen05 kernel where an intersection spans 6 ... 18 chunks of 32 vehicle slots (several producer
    passes per ticket) and tickets do not divide the batch: controlled slots get the dense kernel's action, the others 0."""
    from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
    monkeypatch.setenv("PVE_ACTOR_IMPL", "tc5")
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    scene = P.make_scene("cuda", B, vm=5, veh_cap=caps[0], agent_cap=caps[1])
    actor = BatchedActor(w)
    scene.reset(synthetic_arrivals(B, 1800, 60.0, seed=caps[0], rows=60), warmup=True)
    acts = torch.empty(B, scene.veh_cap, device="cuda")
    for t in range(200):
        actor.act(scene, out=acts)
        scene.step(torch.where(acts > 0, acts * 0 - 3.0, acts))      # brake hard: queues fill the class
    mask = scene.control_mask()
    rows = scene.row0()
    acts = actor.act(scene)
    assert torch.all(acts[~mask] == 0) and scene.stats()["overflow"] == 0
    dense = actor.forward(rows[mask].contiguous())
    assert torch.equal(acts[mask], dense)
    # jammed queues are rows on which the fp32 network itself is less accurate (numpy-fp32 misses the 1e-5 band on more
    # than 3 % of them): hold the kernel to numpy-fp32's own distance from the float64 evaluation
    got, rows32 = dense.cpu().numpy(), rows[mask].cpu().numpy()
    want64 = actor_oracle.actor_forward(w, rows32.astype(np.float64), np.float64)
    err = np.abs(got.astype(np.float64) - want64)
    err_np = np.abs(actor_oracle.actor_forward(w, rows32, np.float32).astype(np.float64) - want64)
    band = lambda e: (e <= RTOL * np.abs(want64) + ATOL).mean()
    assert err.mean() <= 2.0 * err_np.mean() + 1e-7 and err.max() <= WORST and band(err) >= band(err_np) - 0.02, \
        (caps, err.mean(), err_np.mean(), err.max(), band(err), band(err_np))
    assert int(mask.sum()) > 40 * B


def test_rollout_entry_point_equals_the_python_loop():
    """pve_rollout (n ticks of act + step enqueued by the library) against the same ticks driven from Python: state,
    counters and the last tick's outputs bit for bit."""
    from pve_mcc_for_unsignalized_intersection_b200.arrivals import synthetic_arrivals
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    B, T = 48, 333
    scenes = [P.make_scene("cuda", B, vm=5) for _ in range(2)]
    actor = BatchedActor(w)
    for sc in scenes:
        sc.reset(synthetic_arrivals(B, 1000, 60.0, seed=21), warmup=True)
    acts = torch.empty(B, scenes[0].veh_cap, device="cuda")
    for _ in range(T):
        scenes[0].step(actor.act(scenes[0], out=acts))
    out1 = actor.rollout(scenes[1], T - 100)
    out1 = actor.rollout(scenes[1], 100)
    st0, st1 = scenes[0].get_state(), scenes[1].get_state()
    for k in st0:
        assert np.array_equal(np.asarray(st0[k]), np.asarray(st1[k])), k
    n = scenes[0].out.n_agents
    assert n == out1.n_agents and n > 0
    assert torch.equal(scenes[0].out.obs[:n], out1.obs[:n]) and torch.equal(scenes[0].out.reward[:n], out1.reward[:n])
    assert scenes[0].stats() == scenes[1].stats()


def test_closed_loop_reproduces_config1_of_the_reference():
    """BASELINE.json config 1: pretrained actor, arvTimeNewVeh_new_1000_12.mat, 1000 ticks -- here with the
    CUDA actor and the CUDA scene in a loop, against the trace recorded from the unmodified reference scene
    (tests/golden/make_actor_golden.py).  64 copies of the intersection must all agree."""
    z = np.load(os.path.join(GOLD, "actor_rollout_mat1000.npz"))
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    B = 64
    scene = P.make_scene("cuda", B, vm=5)
    actor = BatchedActor(w)
    scene.reset(z["arrive_time"], warmup=True)
    acts = torch.empty(B, scene.veh_cap, device="cuda")
    coll = torch.zeros(B, dtype=torch.int64, device="cuda")
    lock = torch.zeros(B, dtype=torch.int64, device="cuda")
    trace = z["trace"]
    for t in range(trace.shape[0]):
        actor.act(scene, out=acts)
        out = scene.step(acts)
        n = out.n_agents
        off = out.agent_offset.long()
        hit = (out.cpv[:n] > 0).long()
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), hit.cumsum(0)])
        coll += csum[off[1:]] - csum[off[:-1]]                                   # main.py:410-412
        lock += out.env_lock.long()
        if t % 50 == 49 or t == trace.shape[0] - 1:
            st = scene.get_state()
            per_env = (off[1:] - off[:-1]).cpu().numpy()
            assert np.all(per_env == trace[t, 0]), "tick %d agents %s vs %d" % (t, per_env[:4], trace[t, 0])
            assert np.all(st["id_seq"] == trace[t, 1]) and np.all(st["passed_veh"] == trace[t, 2])
            assert np.all(st["passed_step_total"] == trace[t, 5])
            assert torch.all(lock == int(trace[t, 3])) and torch.all(coll == int(trace[t, 4]))
            rs = out.reward[:n].double().sum().item() / B
            assert abs(rs - z["reward_sum"][t]) <= 1e-3 * max(1.0, abs(z["reward_sum"][t]))
    st = scene.get_state()
    ptm = st["passed_step_total"][0] / (st["passed_veh"][0] + 0.0001) * 0.1         # main.py:525
    assert [int(st["id_seq"][0]), int(coll[0]), int(st["passed_veh"][0]), int(lock[0])] == z["outcome"].tolist()
    assert abs(ptm - float(z["ptm"])) < 1e-9


def test_evaluation_report_matches_the_reference_run():
    """Row N4: evaluate_tables = the reference's test driver (main.py:394-441 / 553-581); tables of different
    length side by side; the report line equals the one the reference prints for the recorded run."""
    from pve_mcc_for_unsignalized_intersection_b200 import evaluate
    z = np.load(os.path.join(GOLD, "actor_rollout_mat1000.npz"))
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    arr = z["arrive_time"]
    res = evaluate.evaluate_tables([arr, arr[:20], arr], w, ticks=1000)
    v, c, p, l = z["outcome"].tolist()
    for r in (res[0], res[2]):
        assert [r["vehicles"], r["collisions"], r["passed"], r["lock_total"]] == [v, c, p, l]
        assert r["passed_step_total"] == int(z["trace"][-1, 5])
        assert abs(r["jerk_total"] - float(z["jerk_total"])) <= 1e-5 * float(z["jerk_total"])
        head, tail = r["report"].split(" jerks ")
        assert head == "vehicle number 323  collisions occurred number 0 collisions rate 0.0 pT-m 12.2943 s"
        assert tail.endswith(" lock_num 548") and abs(float(tail.split()[0]) - 208.79941400944244) < 2e-3
    assert res[1]["vehicles"] < v                     # the truncated table runs dry earlier


@pytest.mark.parametrize("density", [400, 1200])
def test_long_evaluation_matches_the_reference_run(density):
    """Row N4 at the length of a real evaluation: 6000 ticks of the closed loop (CUDA actor + CUDA scene) on two of the
    density files of main.py:543 ``batch_test`` -- 400 veh/h, where the reference's own logging loop raises,
    and 1200 veh/h -- against the run of the unmodified reference scene recorded by tests/golden/make_eval_golden.py;
    the tallies come from the per-intersection device counters (no per-tick host work).

    At 400 veh/h the closed loop reproduces the recorded run exactly.  At 1200 veh/h it is chaotic over 6000 ticks: the
    policy is fp32 and its GPU evaluation differs from the numpy one in the last bits (the reference's TensorFlow graph
    would too), one flipped near-collision changes who is removed, so only what the arrival table fixes (vehicles) is
    exact there and the report quantities are held to the run-to-run spread of such perturbations."""
    from pve_mcc_for_unsignalized_intersection_b200 import evaluate
    z = np.load(os.path.join(GOLD, "eval_mat%d_6000.npz" % density))
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    res = evaluate.evaluate_tables([z["arrive_time"]] * 2, w, ticks=6000)
    v, c, p, l, pst = z["outcome"].tolist()
    assert res[0] == res[1]                               # the device itself is deterministic
    for r in res:
        if density == 400:
            assert [r["vehicles"], r["collisions"], r["passed"], r["lock_total"], r["passed_step_total"]] == [v, c, p, l, pst], (r, z["outcome"])
            assert abs(r["jerk_total"] - float(z["jerk_total"])) <= 1e-5 * float(z["jerk_total"])
            assert r["report"] == evaluate.format_report(v, c, p, pst, r["jerk_total"], l)
        else:
            assert r["vehicles"] == v
            assert abs(r["passed"] - p) <= 0.005 * p and abs(r["collisions"] - c) <= 8
            assert abs(r["lock_total"] - l) <= 0.03 * l and abs(r["passed_step_total"] / r["passed"] - pst / p) <= 0.005 * pst / p
            assert abs(r["jerk_total"] / r["passed"] - float(z["jerk_total"]) / p) <= 0.02 * float(z["jerk_total"]) / p
