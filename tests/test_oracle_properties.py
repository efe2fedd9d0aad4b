"""Closed-form and property checks of the oracle: the oracle is what every GPU
parity test is judged against, so besides the reference-generated golden vectors (tests/test_oracle_golden.py) it is
held to properties that do not depend on any implementation: streaming == offline for the carried-state
convolutions whatever the chunking, the ring-buffer attention == a plain windowed attention over the full history
with the write-before-attend visibility rule spelled out, RoPE scores depend on position differences only, the
LSD/Euler integrator on a linear field has a closed form, batch-1 prefill in pieces == in one go."""

import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import rel_l2


@pytest.fixture(scope="module")
def orc64(cfg, weights):
    from oracle.ptts_oracle import Oracle
    return Oracle(weights, cfg, dtype=np.float64, eos_threshold=1e30)


def _splits(draw, total, max_parts=5):
    """Random composition of `total` into 1..max_parts positive chunk sizes."""
    n = draw(st.integers(1, min(max_parts, total)))
    cuts = sorted(draw(st.lists(st.integers(1, total - 1), min_size=n - 1, max_size=n - 1, unique=True))) if total > 1 else []
    edges = [0] + cuts + [total]
    return [b - a for a, b in zip(edges[:-1], edges[1:]) if b > a]


@st.composite
def chunked(draw, lo=3, hi=24):
    total = draw(st.integers(lo, hi))
    return total, _splits(draw, total), draw(st.integers(0, 2 ** 31 - 1))


CONV = "mimi.decoder.model.0.conv"             # k = 7, 512 -> 512 (modules/seanet.py:136-141)
RES_CONV = "mimi.decoder.model.3.block.1.conv"  # k = 3 residual-block conv
CONVTR = "mimi.decoder.model.2.convtr"          # first transposed conv, stride = ratios[0], k = 2 * stride


@settings(max_examples=25, deadline=None)
@given(chunked())
def test_streaming_conv_is_chunk_invariant_and_causal(orc64, case):
    """modules/conv.py:113-150: feeding T rows in any chunking gives the rows of ONE causal convolution of the whole
    input with k-1 zero rows in front (the initial `previous`)."""
    from oracle.ptts_oracle import MimiState
    total, parts, seed = case
    for name in (CONV, RES_CONV):
        w = orc64.w[name + ".weight"]
        k, c_in = w.shape[2], w.shape[1]
        x = np.random.default_rng(seed).standard_normal((total, c_in))
        st_a = orc64.new_mimi_state()
        y_stream = np.concatenate([orc64._conv(name, x[a:a + n], st_a) for a, n in zip(np.cumsum([0] + parts[:-1]), parts)])
        xp = np.concatenate([np.zeros((k - 1, c_in)), x])
        y_off = orc64.w[name + ".bias"] + sum(xp[j:j + total] @ w[:, :, j].T for j in range(k))
        assert rel_l2(y_stream, y_off) < 1e-12
        assert isinstance(st_a, MimiState) and st_a.conv_prev[name].shape == (k - 1, c_in)
        np.testing.assert_array_equal(st_a.conv_prev[name], xp[-(k - 1):])


@settings(max_examples=25, deadline=None)
@given(chunked(lo=2, hi=12))
def test_streaming_convtr_is_chunk_invariant(orc64, case):
    """modules/conv.py:182-200: overlap-add with a carried partial == the first T*s rows of the offline transposed
    convolution of the whole input (bias added once per output row)."""
    total, parts, seed = case
    w = orc64.w[CONVTR + ".weight"]                     # [in, out, k]
    stride = int(orc64.ratios[0])
    k = w.shape[2]
    x = np.random.default_rng(seed).standard_normal((total, w.shape[0]))
    st_a = orc64.new_mimi_state()
    y_stream = np.concatenate([orc64._convtr(CONVTR, x[a:a + n], st_a, stride)
                               for a, n in zip(np.cumsum([0] + parts[:-1]), parts)])
    full = np.zeros(((total - 1) * stride + k, w.shape[1]))
    for t in range(total):
        for j in range(k):
            full[t * stride + j] += x[t] @ w[:, :, j]
    y_off = full[: total * stride] + orc64.w[CONVTR + ".bias"]
    assert y_stream.shape == y_off.shape
    assert rel_l2(y_stream, y_off) < 1e-12


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 4000), st.integers(0, 4000), st.integers(0, 2000), st.integers(0, 2 ** 31 - 1))
def test_rope_scores_depend_on_relative_position_only(p_q, p_k, shift, seed):
    """modules/rope.py:9-42: <rope(q, a), rope(k, b)> == <rope(q, a + s), rope(k, b + s)>, and rope is an isometry."""
    from oracle.ptts_oracle import rope_rotate
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((1, 2, 64))
    k = rng.standard_normal((1, 2, 64))
    a = np.einsum("thd,thd->h", rope_rotate(q, [p_q]), rope_rotate(k, [p_k]))
    b = np.einsum("thd,thd->h", rope_rotate(q, [p_q + shift]), rope_rotate(k, [p_k + shift]))
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(rope_rotate(q, [p_q]), axis=-1), np.linalg.norm(q, axis=-1), rtol=1e-12)
    np.testing.assert_allclose(rope_rotate(q, [0]), q, atol=0)


def test_ring_attention_equals_windowed_attention_over_the_history(orc64):
    """modules/attention.py:67-105, 220-264.  The 250-slot ring is written BEFORE the chunk attends, so a query at
    position p of the chunk [e, e+T) sees key position j iff  j <= p,  p - j < 250  and  j > (e+T-1) - 250  (the
    chunk's own writes have already evicted the oldest T-1 keys that a sliding window would still show to the chunk's
    first rows).  Stated on the full history without any ring; 20 chunks of 16 cross the wrap at 250."""
    from oracle.ptts_oracle import rope_rotate, softmax_rows
    rng = np.random.default_rng(3)
    w = orc64.w
    i = 0
    p = f"mimi.decoder_transformer.transformer.layers.{i}.self_attn"
    mh, md, cap = orc64.mh, orc64.md, orc64.context
    dh = md // mh
    stt = orc64.new_mimi_state()
    hist_k, hist_v = [], []
    for chunk in range(20):
        t = 16
        x = rng.standard_normal((t, md))
        e = stt.offset
        got = orc64._mimi_attention(i, x, stt)
        stt.offset += t
        stt.end_offset += t
        qkv = (x @ w[p + ".in_proj.weight"].T).reshape(t, 3, mh, dh)
        pos = e + np.arange(t)
        q = rope_rotate(qkv[:, 0], pos, orc64.m_max_period)
        hist_k.append(rope_rotate(qkv[:, 1], pos, orc64.m_max_period))
        hist_v.append(qkv[:, 2])
        K, V = np.concatenate(hist_k), np.concatenate(hist_v)
        j = np.arange(K.shape[0])
        last = e + t - 1
        vis = (j[None, :] <= pos[:, None]) & (pos[:, None] - j[None, :] < cap) & (j[None, :] > last - cap)
        s = np.einsum("thd,lhd->htl", q, K) / math.sqrt(dh)
        s = np.where(vis[None], s, -np.inf)
        want = np.einsum("htl,lhd->thd", softmax_rows(s), V).reshape(t, md) @ w[p + ".out_proj.weight"].T
        assert rel_l2(got, want) < 1e-10, chunk
    assert stt.offset == 320 > cap


@pytest.mark.parametrize("n_steps", [1, 2, 5])
def test_lsd_euler_on_a_linear_field_has_the_closed_form(cfg, weights, n_steps):
    """models/flow_lm.py:18-28: x <- x + v(s, t, x) / n from x0 = sqrt(temp) * z.  For v = a x + b (whatever s, t):
    x_n = (1 + a/n)^n x0 + b ((1 + a/n)^n - 1) / a; the time arguments are (i/n, (i+1)/n)."""
    from oracle.ptts_oracle import Oracle
    orc = Oracle(weights, cfg, dtype=np.float64, eos_threshold=1e30, temp=0.49, lsd_decode_steps=n_steps)
    a, b = -0.7, 0.3
    seen = []

    def field(c, s, t, x):
        seen.append((s, t))
        return a * x + b

    orc.flow_velocity = field
    z = np.random.default_rng(1).standard_normal(orc.ldim)
    got = orc.sample_latent(np.zeros(orc.d if hasattr(orc, "d") else 1024), z)
    g = (1 + a / n_steps) ** n_steps
    want = g * (0.7 * z) + b * (g - 1) / a
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert seen == [(i / n_steps, (i + 1) / n_steps) for i in range(n_steps)]


@settings(max_examples=8, deadline=None)
@given(st.integers(2, 14), st.integers(0, 2 ** 31 - 1))
def test_flow_prefill_in_pieces_equals_prefill_in_one_go(orc64, n_tok, seed):
    """modules/attention.py:150-182 with the KV cache of stateful_module.py: text prefilled as [a | b] in two calls
    leaves the same keys / values as one call, and the next decode step sees the same hidden state -- the property the
    paged KV cache, the chunked prefill and continuous batching all rely on."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, 4000, size=n_tok)
    cut = int(rng.integers(1, n_tok))
    s1 = orc64.new_flow_state()
    orc64.prefill_text(s1, ids)
    s2 = orc64.new_flow_state()
    orc64.prefill_text(s2, ids[:cut])
    orc64.prefill_text(s2, ids[cut:])
    assert s1.length == s2.length == n_tok
    for l in range(len(s1.k)):
        assert rel_l2(s2.k[l], s1.k[l]) < 1e-10
        assert rel_l2(s2.v[l], s1.v[l]) < 1e-10
    clone = s1.clone()
    z = rng.standard_normal(orc64.ldim)
    lat = rng.standard_normal(orc64.ldim)
    out1 = orc64.step(s1, lat, z)
    out2 = orc64.step(s2, lat, z)
    assert rel_l2(out2[0], out1[0]) < 1e-9
    assert clone.length == n_tok and s1.length == n_tok + 1          # clone() is a deep copy (tts_model.py:372-373)
