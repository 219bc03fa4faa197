"""Properties of the oracle itself (CPU, float64): the two loss restatements agree with each other on random crop layouts,
and the hand-derived backward formulas agree with central finite differences of the forward.  The golden vectors pin the
oracle to the reference on four fixed cases (tests/test_oracle_golden.py); these properties cover the shapes in between
(any C, any G <= C, ragged K, warm-up temperatures, 1-3 MLP layers, trainable gain)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import np_oracle as O


def _case(seed, C, G, B, K):
    r = np.random.default_rng(seed)
    s = r.normal(0, 1.5, (C * B, K))
    t = r.normal(0, 1.0, (G * B, K))
    c = r.normal(0, 0.3, (1, K))
    return s, t, c


layout = st.tuples(st.integers(0, 2 ** 31 - 1), st.integers(2, 9), st.integers(1, 3), st.integers(1, 4), st.integers(3, 37),
                   st.sampled_from([0.04, 0.055, 0.07]), st.sampled_from([0.1, 0.2]))


@settings(max_examples=40, deadline=None, derandomize=True)
@given(layout)
def test_closed_form_equals_pair_loop(p):
    seed, C, G, B, K, temp, st_temp = p
    G = min(G, C)
    s, t, c = _case(seed, C, G, B, K)
    a = O.dino_loss_loop(s, t, c, temp, C, G, st_temp)
    b = O.dino_loss_closed(s, t, c, temp, C, G, st_temp)
    assert abs(a - b) <= 1e-12 * max(1.0, abs(a))


@settings(max_examples=25, deadline=None, derandomize=True)
@given(layout)
def test_loss_gradient_equals_finite_differences(p):
    seed, C, G, B, K, temp, st_temp = p
    G = min(G, C)
    s, t, c = _case(seed, C, G, B, K)
    g = O.dino_loss_grad(s, t, c, temp, C, G, st_temp)
    r = np.random.default_rng(seed + 1)
    for _ in range(3):                                      # directional derivatives along random directions
        d = r.normal(0, 1, s.shape)
        h = 1e-5
        fd = (O.dino_loss_loop(s + h * d, t, c, temp, C, G, st_temp) - O.dino_loss_loop(s - h * d, t, c, temp, C, G, st_temp)) / (2 * h)
        an = float((g * d).sum())
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (fd, an)
    # every sample's gradient sums to zero over the classes for crops all teachers see: n_v p_v - sum q has total mass 0
    assert np.abs(g.sum(axis=1)).max() < 1e-12


def _head(seed, nlayers, in_dim, hidden, bott, K):
    r = np.random.default_rng(seed)
    sd = {}
    if nlayers == 1:
        sd["mlp.weight"], sd["mlp.bias"] = r.normal(0, 0.3, (bott, in_dim)), r.normal(0, 0.1, bott)
    else:
        dims = [in_dim] + [hidden] * (nlayers - 1) + [bott]
        idx = 0
        for li in range(nlayers):
            sd[f"mlp.{idx}.weight"], sd[f"mlp.{idx}.bias"] = r.normal(0, 0.3, (dims[li + 1], dims[li])), r.normal(0, 0.1, dims[li + 1])
            idx += 2
    sd["last_layer.weight_g"] = 1.0 + 0.2 * r.normal(0, 1, (K, 1))
    sd["last_layer.weight_v"] = r.normal(0, 0.5, (K, bott))
    return sd


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.integers(2, 6), st.integers(3, 9))
def test_head_backward_equals_finite_differences(seed, nlayers, N, K):
    in_dim, hidden, bott = 5, 7, 4
    sd = _head(seed, nlayers, in_dim, hidden, bott, K)
    r = np.random.default_rng(seed + 7)
    x = r.normal(0, 1, (N, in_dim))
    up = r.normal(0, 1, (N, K))                             # upstream gradient: L = sum(up * logits)
    logits, cache = O.head_forward(x, sd, return_cache=True)
    grads = O.head_backward(up, cache)

    def L(sd_, x_):
        return float((up * O.head_forward(x_, sd_)).sum())

    h = 1e-6
    for name in list(sd) + ["x"]:
        base = x if name == "x" else sd[name]
        d = r.normal(0, 1, base.shape)
        if name == "x":
            fd = (L(sd, x + h * d) - L(sd, x - h * d)) / (2 * h)
        else:
            fd = (L({**sd, name: base + h * d}, x) - L({**sd, name: base - h * d}, x)) / (2 * h)
        an = float((np.asarray(grads[name]).reshape(base.shape) * d).sum())
        assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)), (name, fd, an)
    # weight-norm geometry: dv is orthogonal to v row by row; logits of unit rows are bounded by |g|
    v = sd["last_layer.weight_v"]
    assert np.abs((grads["last_layer.weight_v"] * v).sum(1)).max() < 1e-10 * max(1.0, np.abs(grads["last_layer.weight_v"]).max())
    assert (np.abs(logits) <= np.abs(sd["last_layer.weight_g"]).T + 1e-12).all()


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.floats(0.9, 1.0))
def test_ema_and_center_identities(seed, m):
    r = np.random.default_rng(seed)
    pk = [r.normal(0, 1, (3, 5)).astype(np.float32), r.normal(0, 1, 7).astype(np.float32)]
    pq = [r.normal(0, 1, (3, 5)).astype(np.float32), r.normal(0, 1, 7).astype(np.float32)]
    same = O.ema_update_fp32(pk, pq, 1.0)
    assert all(np.array_equal(a, b) for a, b in zip(same, pk))                     # m = 1: exact identity (StepGraph warm-up relies on it)
    out = O.ema_update_fp32(pk, pq, m)
    for o, a, b in zip(out, pk, pq):
        assert o.dtype == np.float32
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        assert (o >= lo - 1e-6).all() and (o <= hi + 1e-6).all()                   # a convex combination, up to fp32 rounding
    # center: splitting the teacher rows over two ranks and all-reducing equals the single-rank update on all rows
    t = r.normal(0, 1, (8, 11))
    c0 = r.normal(0, 0.3, (1, 11))
    one = O.update_center(c0, t, 0.9)
    two = O.update_center(c0, t[:4], 0.9, world_size=2, all_rank_outputs=[t[:4], t[4:]])
    assert np.abs(one - two).max() < 1e-14
