"""CPU, only where the reference was compiled (oracle/_ref exists — the build container and, because
oracle/_ref travels with gpurun, the GPU box): the C oracle against the UNMODIFIED reference on
fresh seeded inputs, larger than the committed fixtures."""
import numpy as np
import pytest

from refutil import (F16, F32, I64, Q8_0, Q8_K, WEIGHT_TYPES, have_ref, nmse, oracle, orc_mul_mat, orc_quantize_act, ptr, rand_blocks,
                     ref, ref_quantize_weights, row_bytes, run_ref_op)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")


@pytest.mark.parametrize("k", [256, 4096, 14336])
def test_quantisers_bit_exact(k):
    _, cpu = ref()
    rng = np.random.default_rng(k)
    x = (rng.standard_normal(k) * rng.uniform(0.01, 50)).astype(np.float32)
    for t, orc_fn, ref_fn in ((Q8_0, oracle().orc_quantize_row_q8_0, cpu.quantize_row_q8_0), (Q8_K, oracle().orc_quantize_row_q8_K, cpu.quantize_row_q8_K)):
        a = np.zeros(row_bytes(t, k), np.uint8); b = a.copy()
        orc_fn(ptr(x), ptr(a), k); ref_fn(ptr(x), ptr(b), k)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("t", WEIGHT_TYPES)
def test_mul_mat_through_cpu_backend(t):
    rng = np.random.default_rng(t)
    m, k, n = 32, 4096, 2
    for W in (ref_quantize_weights(t, (rng.standard_normal((m, k)) * 0.02).astype(np.float32)), rand_blocks(rng, t, m, k)):
        X = rng.standard_normal((n, k)).astype(np.float32)
        _, _, out = run_ref_op("mul_mat", [("w", t, [k, m], W), ("x", F32, [k, n], X)])
        want = np.frombuffer(out, np.float32).reshape(n, m)
        got = orc_mul_mat(t, W, X, m, n, k)
        assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("kvt", [F16, Q8_0])
def test_flash_attn_through_cpu_backend(kvt):
    rng = np.random.default_rng(kvt)
    dk, nh, nhkv, nt, nkv = 64, 8, 1, 1, 512
    q = rng.standard_normal((nt, nh, dk)).astype(np.float32)
    kf = rng.standard_normal((nkv, nhkv * dk)).astype(np.float32); vf = rng.standard_normal((nkv, nhkv * dk)).astype(np.float32)
    rb_row, rb_head = row_bytes(kvt, nhkv * dk), row_bytes(kvt, dk)
    kc = np.zeros((nkv, rb_row), np.uint8); vc = kc.copy(); ids = np.arange(nkv, dtype=np.int64)
    oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), kvt, nhkv * dk, nkv, rb_row)
    oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), kvt, nhkv * dk, nkv, rb_row)
    mask = np.full((64, nkv), -np.inf, np.float32); mask[0, :333] = 0
    m16 = mask.astype(np.float16)
    _, _, out = run_ref_op("flash_attn", [("q", F32, [dk, nh, nt], q), ("k", kvt, [nhkv * dk, nkv], kc), ("v", kvt, [nhkv * dk, nkv], vc), ("mask", F16, [nkv, 64], m16)],
                           dict(dk=dk, dv=dk, n_head_kv=nhkv, n_kv=nkv, scale=0.125))
    want = np.frombuffer(out, np.float32).reshape(nt, nh, dk)
    y = np.zeros_like(want)
    oracle().orc_flash_attn_ext(ptr(q), nh * dk * 4, dk * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), ptr(y),
                                kvt, dk, dk, nh, nhkv, nt, nkv, 0.125, 0.0, 0.0)
    assert np.abs(y - want).max() <= (2e-3 if kvt == F16 else 2e-6) * np.abs(want).max()


def test_reference_parity_harness_runs():
    """the reference's own test-backend-ops (tests/test-backend-ops.cpp, unmodified) loads and lists the CPU device"""
    import os
    import subprocess
    from refutil import REF_DIR
    r = subprocess.run([os.path.join(REF_DIR, "test-backend-ops"), "support", "-o", "MUL_MAT"], capture_output=True, text=True,
                       env=dict(os.environ, LD_LIBRARY_PATH=REF_DIR))
    assert r.returncode == 0 and "CPU" in (r.stdout + r.stderr)


# ---- the wide kernels' formats, MUL_MAT_ID, GET_ROWS on quantised tables (oracle/oracle_ext.c) ----------
from refutil import EXT_TYPES, ACT_TYPE, I32, Q8_1, TYPE_NAME, orc_dequant  # noqa: E402


def test_q8_1_quantiser_bit_exact():
    _, cpu = ref()
    for k in (32, 4096, 29568):
        rng = np.random.default_rng(k)
        x = (rng.standard_normal(k) * rng.uniform(0.01, 50)).astype(np.float32)
        x[:32] = 0
        a = np.zeros(row_bytes(Q8_1, k), np.uint8); b = a.copy()
        oracle().orc_quantize_row_q8_1(ptr(x), ptr(a), k); cpu.quantize_row_q8_1(ptr(x), ptr(b), k)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("t", EXT_TYPES)
def test_ext_dequant_bit_exact_and_vec_dot(t):
    base, cpu = ref()
    rng = np.random.default_rng(100 + t)
    m, k = 8, 2048
    for W in (ref_quantize_weights(t, (rng.standard_normal((m, k)) * 0.05).astype(np.float32)), rand_blocks(rng, t, m, k)):
        deq = np.zeros((m, k), np.float32)
        for i in range(m):
            getattr(base, "dequantize_row_" + TYPE_NAME[t])(ptr(W[i]), ptr(deq[i]), k)
        assert np.array_equal(orc_dequant(t, W, m, k), deq)
        x = rng.standard_normal((1, k)).astype(np.float32)
        aq = orc_quantize_act(t, x)
        fn = getattr(cpu, f"ggml_vec_dot_{TYPE_NAME[t]}_{TYPE_NAME[ACT_TYPE[t]]}")
        for i in range(m):
            want = np.zeros(1, np.float32)
            fn(k, ptr(want), 0, ptr(W[i]), 0, ptr(aq[0]), 0, 1)
            got = oracle().orc_vec_dot(t, k, ptr(W[i]), ptr(aq[0]))
            assert abs(got - want[0]) <= 3e-6 * max(abs(want[0]), np.abs(deq[i]).max() * 30)


@pytest.mark.parametrize("t", EXT_TYPES)
def test_ext_mul_mat_through_cpu_backend(t):
    rng = np.random.default_rng(200 + t)
    m, k, n = 32, 4096, 2
    for W in (ref_quantize_weights(t, (rng.standard_normal((m, k)) * 0.02).astype(np.float32)), rand_blocks(rng, t, m, k)):
        X = rng.standard_normal((n, k)).astype(np.float32)
        _, _, out = run_ref_op("mul_mat", [("w", t, [k, m], W), ("x", F32, [k, n], X)])
        want = np.frombuffer(out, np.float32).reshape(n, m)
        got = orc_mul_mat(t, W, X, m, n, k)
        assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("t", [2, 12, 14, 3, 11, 23])
@pytest.mark.parametrize("n_b1", [1, 0])
def test_mul_mat_id_through_cpu_backend(t, n_b1):
    """MUL_MAT_ID as build_moe_ffn emits it (llama-graph.cpp): shared activation (up / gate) and one activation per used expert (down)"""
    rng = np.random.default_rng(300 + t)
    m, k, n_expert, n_used, n_tok = 24, 512, 6, 3, 4
    n_b1 = n_b1 or n_used
    W = ref_quantize_weights(t, (rng.standard_normal((n_expert * m, k)) * 0.05).astype(np.float32))
    b = rng.standard_normal((n_tok, n_b1, k)).astype(np.float32)
    ids = np.stack([rng.permutation(n_expert)[:n_used] for _ in range(n_tok)]).astype(np.int32)
    _, ne, out = run_ref_op("mul_mat_id", [("w", t, [k, m, n_expert], W), ("x", F32, [k, n_b1, n_tok], b), ("ids", I32, [n_used, n_tok], ids)])
    assert list(ne[:3]) == [m, n_used, n_tok]
    want = np.frombuffer(out, np.float32).reshape(n_tok, n_used, m)
    got = np.zeros_like(want)
    oracle().orc_mul_mat_id(t, ptr(W), ptr(b), ptr(ids), ptr(got), m, k, n_expert, n_used, n_tok, n_b1, n_used)
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("t", WEIGHT_TYPES + EXT_TYPES)
def test_get_rows_quantised_through_cpu_backend(t):
    rng = np.random.default_rng(400 + t)
    nrows, k = 40, 1024
    W = ref_quantize_weights(t, (rng.standard_normal((nrows, k)) * 0.05).astype(np.float32))
    ids = np.array([3, 39, 0, 3, 17], np.int32)
    _, _, out = run_ref_op("get_rows", [("src", t, [k, nrows], W), ("ids", I32, [ids.size], ids)])
    want = np.frombuffer(out, np.float32).reshape(ids.size, k)
    got = np.zeros_like(want)
    oracle().orc_get_rows_q(t, ptr(W), ptr(ids), ptr(got), k, ids.size)
    assert np.array_equal(got, want)


def test_q4_0_kv_cache_set_rows_and_flash_attn_through_cpu_backend():
    """KV cache type q4_0 (llama-box -ctk q4_0 -ctv q4_0): SET_ROWS bytes identical, FLASH_ATTN_EXT within the Q8_0 tolerance"""
    from refutil import Q4_0
    rng = np.random.default_rng(40)
    dk, nh, nhkv, nt, nkv = 128, 8, 2, 2, 512
    kf = rng.standard_normal((nkv, nhkv * dk)).astype(np.float32); vf = rng.standard_normal((nkv, nhkv * dk)).astype(np.float32)
    kf[3, :32] = 0; kf[5, 7] = -kf[5, :32].max() if False else kf[5, 7]
    rb_row, rb_head = row_bytes(Q4_0, nhkv * dk), row_bytes(Q4_0, dk)
    z = np.zeros((nkv, rb_row), np.uint8); ids = np.arange(nkv, dtype=np.int64)
    _, _, kc_ref = run_ref_op("set_rows", [("cache", Q4_0, [nhkv * dk, nkv], z), ("src", F32, [nhkv * dk, nkv], kf), ("ids", I64, [nkv], ids)])
    _, _, vc_ref = run_ref_op("set_rows", [("cache", Q4_0, [nhkv * dk, nkv], z), ("src", F32, [nhkv * dk, nkv], vf), ("ids", I64, [nkv], ids)])
    kc = z.copy(); vc = z.copy()
    oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), Q4_0, nhkv * dk, nkv, rb_row)
    oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), Q4_0, nhkv * dk, nkv, rb_row)
    assert np.array_equal(kc.reshape(-1), np.frombuffer(kc_ref, np.uint8)) and np.array_equal(vc.reshape(-1), np.frombuffer(vc_ref, np.uint8))
    q = rng.standard_normal((nt, nh, dk)).astype(np.float32)
    mask = np.full((64, nkv), -np.inf, np.float32); mask[0, :333] = 0; mask[1, :334] = 0
    m16 = mask.astype(np.float16)
    _, _, out = run_ref_op("flash_attn", [("q", F32, [dk, nh, nt], q), ("k", Q4_0, [nhkv * dk, nkv], kc), ("v", Q4_0, [nhkv * dk, nkv], vc), ("mask", F16, [nkv, 64], m16)],
                           dict(dk=dk, dv=dk, n_head_kv=nhkv, n_kv=nkv, scale=float(1 / np.sqrt(dk))))
    want = np.frombuffer(out, np.float32).reshape(nt, nh, dk)
    y = np.zeros_like(want)
    oracle().orc_flash_attn_ext(ptr(q), nh * dk * 4, dk * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), ptr(y),
                                Q4_0, dk, dk, nh, nhkv, nt, nkv, float(1 / np.sqrt(dk)), 0.0, 0.0)
    assert np.abs(y - want).max() <= 2e-6 * np.abs(want).max()


def test_moe_router_glue_through_cpu_backend():
    """SOFT_MAX (no mask), ARGSORT incl. ties, SUM_ROWS, DIV with broadcasting, the batched GET_ROWS and the broadcast MUL of build_moe_ffn"""
    rng = np.random.default_rng(77)
    ne, nt, nu = 8, 5, 2
    x = (rng.standard_normal((nt, ne)) * 3).astype(np.float32)
    _, _, out = run_ref_op("soft_max", [("x", F32, [ne, nt], x)], dict(scale=1.0))
    want = np.frombuffer(out, np.float32).reshape(nt, ne); y = np.zeros_like(x)
    oracle().orc_soft_max_rows(ptr(x), ptr(y), ne, nt, 1.0)
    assert np.abs(y - want).max() <= 1e-6
    xs = x.copy(); xs[1, 3] = xs[1, 5]; xs[2, :] = 0.25                      # ties: the exchange sort's order is part of the contract
    for desc in (1, 0):
        _, _, out = run_ref_op("argsort", [("x", F32, [ne, nt], xs)], dict(desc=desc))
        idx = np.zeros((nt, ne), np.int32); oracle().orc_argsort_rows(ptr(xs), ptr(idx), ne, nt, desc)
        assert np.array_equal(idx, np.frombuffer(out, np.int32).reshape(nt, ne))
    w = rng.uniform(0.01, 1, (nt, nu)).astype(np.float32)
    _, _, out = run_ref_op("sum_rows", [("x", F32, [nu, nt], w)])
    s = np.zeros(nt, np.float32); oracle().orc_sum_rows(ptr(w), ptr(s), nu, nt)
    assert np.array_equal(s, np.frombuffer(out, np.float32))
    _, _, out = run_ref_op("div", [("a", F32, [nu, nt], w), ("b", F32, [1, nt], s)])
    assert np.array_equal(np.frombuffer(out, np.float32).reshape(nt, nu), w / s[:, None])
    probs = y.reshape(nt, ne, 1); ids = np.stack([rng.permutation(ne)[:nu] for _ in range(nt)]).astype(np.int32)
    _, _, out = run_ref_op("get_rows", [("src", F32, [1, ne, nt], probs), ("ids", I32, [nu, nt], ids)])
    assert np.array_equal(np.frombuffer(out, np.float32).reshape(nt, nu), np.take_along_axis(y, ids, axis=1))
    ex = rng.standard_normal((nt, nu, 64)).astype(np.float32)
    _, _, out = run_ref_op("mul", [("a", F32, [64, nu, nt], ex), ("b", F32, [1, nu, nt], w.reshape(nt, nu, 1))])
    assert np.array_equal(np.frombuffer(out, np.float32).reshape(nt, nu, 64), ex * w[:, :, None])


def test_non_flash_attention_ops_through_cpu_backend():
    """attention without -fa: the batched f16 MUL_MATs (KQ with GQA broadcast, KQV) and SOFT_MAX with mask + ALiBi"""
    rng = np.random.default_rng(55)
    hd, nkv, nt, nh, nhk = 64, 96, 3, 4, 2
    Kc = rng.standard_normal((nhk, nkv, hd)).astype(np.float16); Q = rng.standard_normal((nh, nt, hd)).astype(np.float32)
    _, _, out = run_ref_op("mul_mat", [("w", F16, [hd, nkv, nhk], Kc), ("x", F32, [hd, nt, nh], Q)])
    want = np.frombuffer(out, np.float32).reshape(nh, nt, nkv)
    kq = np.zeros_like(want)
    oracle().orc_mul_mat_f16(ptr(Kc), hd * 2, nkv * hd * 2, nhk, ptr(Q), hd * 4, nt * hd * 4, ptr(kq), nkv * 4, nt * nkv * 4, nkv, nt, nh, hd)
    assert np.abs(kq - want).max() <= 2e-6 * np.abs(want).max()
    for mask_t, max_bias in ((F32, 0.0), (F16, 8.0)):
        mask = np.full((64, nkv), -np.inf, np.float32)
        for t in range(nt):
            mask[t, :nkv - nt + t + 1] = rng.uniform(-2, 0, nkv - nt + t + 1) if max_bias > 0 else 0
        m = mask.astype(np.float16) if mask_t == F16 else mask
        _, _, out = run_ref_op("soft_max", [("x", F32, [nkv, nt, nh], want), ("mask", mask_t, [nkv, 64], m)], dict(scale=0.125, max_bias=max_bias))
        wp = np.frombuffer(out, np.float32).reshape(nh, nt, nkv); y = np.zeros_like(wp)
        oracle().orc_soft_max_mask(ptr(want), ptr(y), ptr(m), int(mask_t == F16), nkv, nkv, nt, nh, 0.125, max_bias)
        assert np.abs(y - wp).max() <= 2e-6


def test_unary_ops_through_cpu_backend():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((3, 40)) * 4).astype(np.float32)
    for name, op, prm, tol in (("scale", 0, dict(s=0.37, b=0.0), 0.0), ("scale", 0, dict(s=2.5, b=-0.75), 1e-7), ("silu", 1, {}, 1e-6), ("sigmoid", 2, {}, 1e-7)):
        _, _, out = run_ref_op(name, [("x", F32, [40, 3], x)], prm)
        want = np.frombuffer(out, np.float32).reshape(3, 40); y = np.zeros_like(x)
        oracle().orc_unary(op, ptr(x), ptr(y), x.size, float(prm.get("s", 1.0)), float(prm.get("b", 0.0)))
        assert np.abs(y - want).max() <= tol * max(1.0, np.abs(want).max()), name
