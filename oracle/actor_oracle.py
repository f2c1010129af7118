"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs) -- never imported by
the product path.

numpy restatement of the reference's actor network, model_agent_maddpg.py:23-49 (NET):

    x = layer_norm(state)            NET:26-27
    x = dense(x, 64)                 NET:28-29
    x = relu(layer_norm(x))          NET:30-32
    x = dense(x, 64)                 NET:34-35
    x = relu(layer_norm(x))          NET:36-38
    x = dense(x, 1)                  NET:40-41
    a = 3 * tanh(x)                  NET:46-47

``tf.contrib.layers.layer_norm`` and ``tf.layers.dense`` belong to tensorflow==1.12.0 (README.md:16),
which is neither vendored in the reference nor installed here.  Their published algorithm
(tensorflow/contrib/layers/python/layers/layers.py ``layer_norm``; nn_impl.py ``moments``,
``batch_normalization``) is restated:

    mean = reduce_mean(x, axis=1); var = reduce_mean((x - mean)^2, axis=1)      begin_norm_axis = 1
    inv  = rsqrt(var + 1e-12) * gamma
    y    = x * inv + (beta - mean * inv)

PARITY PINNING: TensorFlow cannot be run in this image, so there is no golden vector produced by the
reference's own graph: at the level of single activations this oracle is "parity unpinned".  What pins it
is the closed loop: driven by this actor with the shipped checkpoint, the *unmodified reference scene*
reproduces the config-1 outcome recorded in BASELINE.md section 2 (323 vehicles, 0 collisions, 281 passed,
pT-m 12.294 s, 548 lock events after 1000 ticks of arvTimeNewVeh_new_1000_12.mat) -- see
tests/golden/make_actor_golden.py and tests/test_actor_oracle.py.
"""
import numpy as np

EPS = 1e-12


def layer_norm(x, gamma, beta, dtype=np.float32):
    x = x.astype(dtype)
    mean = x.mean(axis=1, keepdims=True, dtype=dtype)
    var = np.mean((x - mean) ** 2, axis=1, keepdims=True, dtype=dtype)
    inv = (dtype(1.0) / np.sqrt(var + dtype(EPS))) * gamma.astype(dtype)
    return x * inv + (beta.astype(dtype) - mean * inv)


def actor_forward(weights, rows, dtype=np.float32):
    """``weights``: mapping with the names of actor.PARAM_SPECS; ``rows`` [n, 28] -> actions [n]."""
    t = weights.tensors if hasattr(weights, "tensors") else weights
    x = np.asarray(rows, dtype=dtype).reshape(-1, 28)
    x = layer_norm(x, t["LayerNorm/gamma"], t["LayerNorm/beta"], dtype)
    x = x @ t["dense/kernel"].astype(dtype) + t["dense/bias"].astype(dtype)
    x = np.maximum(layer_norm(x, t["LayerNorm_1/gamma"], t["LayerNorm_1/beta"], dtype), 0)
    x = x @ t["dense_1/kernel"].astype(dtype) + t["dense_1/bias"].astype(dtype)
    x = np.maximum(layer_norm(x, t["LayerNorm_2/gamma"], t["LayerNorm_2/beta"], dtype), 0)
    x = x @ t["dense_2/kernel"].astype(dtype) + t["dense_2/bias"].astype(dtype)
    return (dtype(3.0) * np.tanh(x)).reshape(-1)


def policy_actions(weights, row0, control, dtype=np.float32):
    """main.py:398-404 for one intersection: action of every vehicle slot, 0 where not controlled."""
    row0 = np.asarray(row0, dtype=np.float64)
    out = np.zeros(row0.shape[0], dtype=np.float64)
    idx = np.flatnonzero(control)
    if idx.size:
        out[idx] = actor_forward(weights, row0[idx], dtype)
    return out
