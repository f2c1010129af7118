"""CPU tier for the actor row (N1): fixtures, the numpy oracle, the checkpoint reader."""
import os

import numpy as np
import pytest

import parity  # noqa: F401  (sys.path)
from oracle import actor_oracle
from pve_mcc_for_unsignalized_intersection_b200.actor import ACTOR_FLOATS, PARAM_SPECS, ActorWeights
from pve_mcc_for_unsignalized_intersection_b200.checkpoint import read_bundle_index

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_CKPT = "/root/reference/model_data/baseline"


def test_weight_fixture_layout():
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    assert ACTOR_FLOATS == 6393 and w.flat().shape == (6393,)
    flat = w.flat()
    # order of the flat vector = order of include/pve_mcc.h
    assert np.array_equal(flat[:28], w.tensors["LayerNorm/gamma"])
    assert np.array_equal(flat[56:56 + 28 * 64].reshape(28, 64), w.tensors["dense/kernel"])
    assert flat[-1] == w.tensors["dense_2/bias"][0]
    with pytest.raises(ValueError):
        ActorWeights({n: np.zeros((3,)) for n, _ in PARAM_SPECS})


def test_recorded_rollout_is_config1_of_the_reference():
    """pretrained actor on arvTimeNewVeh_new_1000_12.mat, 1000 ticks."""
    z = np.load(os.path.join(GOLD, "actor_rollout_mat1000.npz"))
    assert z["outcome"].tolist() == [323, 0, 281, 548]
    assert abs(float(z["ptm"]) - 12.294) < 5e-4
    assert z["trace"].shape == (1000, 6) and z["trace"][-1, 1] == 323


def test_oracle_reproduces_recorded_actions():
    z = np.load(os.path.join(GOLD, "actor_rollout_mat1000.npz"))
    w = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    a32 = actor_oracle.actor_forward(w, z["rows"], np.float32)
    a64 = actor_oracle.actor_forward(w, z["rows"], np.float64)
    np.testing.assert_allclose(a64, z["actions_f64"], rtol=0, atol=1e-12)
    # batched vs single-row BLAS summation order in fp32
    np.testing.assert_allclose(a32, z["actions_f32"], rtol=1e-5, atol=1e-5)
    # fp32 against float64: the network is ill-conditioned at the 1e-4 level (inputs ~160 m rounded to fp32
    # alone move some actions by 9e-5), see tests/test_gpu_actor.py for the bound the kernel is held to
    err = np.abs(a32 - a64)
    assert (err <= 1e-5 * np.abs(a64) + 1e-5).mean() >= 0.97 and err.max() <= 5e-4
    assert np.all(np.abs(a64) <= 3.0)


def test_policy_actions_zero_for_uncontrolled():
    w = ActorWeights.random(3)
    rows = np.random.RandomState(0).randn(10, 28)
    ctrl = np.array([1, 0, 1, 1, 0, 0, 1, 0, 0, 1], dtype=bool)
    a = actor_oracle.policy_actions(w, rows, ctrl)
    assert np.all(a[~ctrl] == 0) and np.all(a[ctrl] != 0)


@pytest.mark.skipif(not os.path.isdir(REF_CKPT), reason="reference checkpoint only exists in the build container")
def test_checkpoint_reader_against_the_shipped_bundle():
    index = read_bundle_index(os.path.join(REF_CKPT, "66.cptk.index"))
    assert len(index) == 152
    assert sum(int(np.prod(e["shape"])) for e in index.values()) == 79412
    assert index["agent1actor/dense/kernel"] == {"dtype": 1, "shape": (28, 64), "shard": 0, "offset": 245412,
                                                 "size": 7168}
    w = ActorWeights.from_checkpoint(REF_CKPT)
    f = ActorWeights.from_npz(os.path.join(GOLD, "actor_agent1.npz"))
    for n, _ in PARAM_SPECS:
        assert np.array_equal(w.tensors[n], f.tensors[n]), n


def test_actor_needs_a_gpu_and_says_so():
    import torch
    from pve_mcc_for_unsignalized_intersection_b200 import _native
    from pve_mcc_for_unsignalized_intersection_b200.actor import BatchedActor
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_native.NativeError):
        BatchedActor(ActorWeights.random(0))
    with pytest.raises(_native.NativeError):
        BatchedActor(ActorWeights.random(0), device="cpu")


def test_report_line_and_mat_ingest(tmp_path):
    """Row N4: the batch_test result line (main.py:576-581) and scipy .mat ingest (main.py:388-389)."""
    import scipy.io as scio
    from pve_mcc_for_unsignalized_intersection_b200 import evaluate
    z = np.load(os.path.join(GOLD, "actor_rollout_mat1000.npz"))
    v, c, p, l = z["outcome"].tolist()
    line = evaluate.format_report(v, c, p, int(z["trace"][-1, 5]), float(z["jerk_total"]), l)
    assert line == ("vehicle number 323  collisions occurred number 0 collisions rate 0.0 pT-m 12.2943 s "
                    "jerks 208.79941400944244 lock_num 548")
    path = str(tmp_path / "arvTimeNewVeh_new_1000_12.mat")
    scio.savemat(path, {"arvTimeNewVeh": z["arrive_time"]})
    arr = evaluate.load_arrivals(path)
    assert arr.dtype == np.float64 and np.array_equal(arr, z["arrive_time"])
    stacked = evaluate.stack_tables([arr, arr[:10]])
    assert stacked.shape == (2,) + arr.shape and np.all(stacked[1, 10:] == 0) and np.array_equal(stacked[0], arr)
    scio.savemat(path, {"arvTimeNewVeh": np.zeros((5, 4))})
    with pytest.raises(ValueError):
        evaluate.load_arrivals(path)
