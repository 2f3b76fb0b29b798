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
