"""CPU: llama-box_b200/csrc/extfmt.cuh — the per-sub-block arithmetic of the wide matvec / MUL_MAT_ID / quantised GET_ROWS kernels
 — compiled with g++ (tests/hostsim/extsim.cpp) and checked against the C oracle, which tests/test_oracle_vs_ref.py
pins to the unmodified reference.  The same functions run inside the CUDA kernels; this verifies their bit manipulation without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from refutil import (ACT_TYPE, BLOCK_ELEMS, EXT_TYPES, Q8_0, Q8_1, Q8_K, ROOT, WEIGHT_TYPES, oracle, orc_dequant, orc_mul_mat, orc_quantize_act, ptr,
                     rand_blocks, repack_rows_np, row_bytes)

SIM_DIR = os.path.join(ROOT, "tests", "hostsim")


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(SIM_DIR, "libextsim.so")
    src = os.path.join(SIM_DIR, "extsim.cpp"); hdr = os.path.join(ROOT, "llama-box_b200", "csrc", "extfmt.cuh"); hdr2 = os.path.join(ROOT, "llama-box_b200", "csrc", "repack_layout.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    vp, i64 = C.c_void_p, C.c_int64
    L.sim_row_dot.argtypes = [C.c_int, vp, i64, i64, vp, vp, vp, vp, vp]
    L.sim_row_dequant.argtypes = [C.c_int, vp, i64, i64, vp]
    L.sim_h2f.restype = C.c_float; L.sim_h2f.argtypes = [C.c_uint16]
    L.sim_repack_row.argtypes = [C.c_int, vp, vp, i64, i64]
    L.sim_q4_0_quantize_row.argtypes = [vp, vp, i64]
    L.sim_q4_0n_row_dot.restype = C.c_float; L.sim_q4_0n_row_dot.argtypes = [vp, i64, vp, vp, vp]
    L.sim_q4_0n_row_dequant.argtypes = [vp, i64, vp]
    return L


def split_act(t, aq, k):
    """oracle activation blocks (q8_0 / q8_1 / q8_K) -> the SoA column the kernels keep in shared memory: qs, d, s, bs"""
    at = ACT_TYPE[t]
    if at == Q8_K:
        b = aq.reshape(k // 256, 292)
        d = b[:, 0:4].copy().view(np.float32).reshape(-1)
        qs = b[:, 4:260].copy().view(np.int8).reshape(-1)
        bs = b[:, 260:292].copy().view(np.int16).reshape(-1)
        return qs, d, np.zeros(1, np.float32), bs
    hdr = 2 if at == Q8_0 else 4
    b = aq.reshape(k // 32, hdr + 32)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(-1)
    s = b[:, 2:4].copy().view(np.float16).astype(np.float32).reshape(-1) if at == Q8_1 else np.zeros(k // 32, np.float32)
    qs = b[:, hdr:].copy().view(np.int8).reshape(-1)
    bs = qs.reshape(-1, 32).astype(np.int32).sum(axis=1).astype(np.int16)
    return qs, d, s, bs


def test_half_conversion(sim):
    hs = np.arange(0, 65536, dtype=np.uint32).astype(np.uint16)
    want = hs.view(np.float16).astype(np.float32)
    got = np.array([sim.sim_h2f(int(h)) for h in hs], np.float32)
    fin = np.isfinite(want)
    assert np.array_equal(got[fin], want[fin]) and np.all(np.isnan(got[~fin]) == np.isnan(want[~fin]))


@pytest.mark.parametrize("t", WEIGHT_TYPES + EXT_TYPES)
@pytest.mark.parametrize("k", [256, 4096])
def test_sub_block_dot_and_dequant(sim, t, k):
    rng = np.random.default_rng(1000 * t + k)
    m = 12
    W = rand_blocks(rng, t, m, k)                                   # every bit pattern of quants / scales
    x = (rng.standard_normal((1, k)) * 2).astype(np.float32)
    qs, d, s, bs = split_act(t, orc_quantize_act(t, x)[0], k)
    Wl = repack_rows_np(t, W, k)                                    # library layout where the kernels use it
    nb = k // BLOCK_ELEMS[t]
    want = orc_mul_mat(t, W, x, m, 1, k)[0]
    deq_want = orc_dequant(t, W, m, k)
    scale = np.abs(want).max()
    for i in range(m):
        got = np.zeros(1, np.float32)
        assert sim.sim_row_dot(t, ptr(Wl[i]), k, nb, ptr(qs), ptr(d), ptr(s), ptr(bs), ptr(got)) == 1
        assert abs(got[0] - want[i]) <= 2e-5 * scale, (t, i, got[0], want[i])
        y = np.zeros(k, np.float32)
        assert sim.sim_row_dequant(t, ptr(Wl[i]), k, nb, ptr(y)) == 1
        assert np.array_equal(y, deq_want[i]), (t, i)


@pytest.mark.parametrize("t", [3, 7, 20, 39])
def test_rows_that_are_not_a_multiple_of_256(sim, t):
    """32-element block formats: any k % 32 == 0 (the wide kernels do not need 256-element units)"""
    rng = np.random.default_rng(t)
    k, m = 29568 // 4, 4                                            # 7392 = 28.875 x 256
    W = rand_blocks(rng, t, m, k); x = rng.standard_normal((1, k)).astype(np.float32)
    qs, d, s, bs = split_act(t, orc_quantize_act(t, x)[0], k)
    Wl = repack_rows_np(t, W, k); want = orc_mul_mat(t, W, x, m, 1, k)[0]
    for i in range(m):
        got = np.zeros(1, np.float32)
        sim.sim_row_dot(t, ptr(Wl[i]), k, k // 32, ptr(qs), ptr(d), ptr(s), ptr(bs), ptr(got))
        assert abs(got[0] - want[i]) <= 2e-5 * np.abs(want).max()


def test_q4_0_kv_cache_block_functions(sim):
    """KV cache type q4_0: the quantiser SET_ROWS runs (bytes identical to ggml's from_float, incl. f16 rounding of d over 12 decades, ties, zero
    blocks), the K-row dot against the q8_0 form of a query and the V de-quantisation of the wide attention kernel"""
    from refutil import Q4_0
    rng = np.random.default_rng(5)
    k = 32 * 4096
    x = (rng.standard_normal(k) * 10.0 ** rng.uniform(-7, 5, k // 32).repeat(32)).astype(np.float32)
    x[:32] = 0; x[32:64] = 1.0; x[64:96] = -1.0; x[96] = 3.0; x[97] = -3.0; x[128:160] = 1e-9
    want = np.zeros(row_bytes(Q4_0, k), np.uint8); got = want.copy()
    oracle().orc_quantize_row_q4_0(ptr(x), ptr(want), k); sim.sim_q4_0_quantize_row(ptr(x), ptr(got), k)
    assert np.array_equal(got, want)
    kk = 128
    K = rand_blocks(rng, Q4_0, 16, kk); q = rng.standard_normal((1, kk)).astype(np.float32)
    qs, d, s, bs = split_act(Q4_0, orc_quantize_act(Q4_0, q)[0], kk)
    wantd = orc_mul_mat(Q4_0, K, q, 16, 1, kk)[0]; deq = orc_dequant(Q4_0, K, 16, kk)
    for i in range(16):
        assert abs(sim.sim_q4_0n_row_dot(ptr(K[i]), kk, ptr(qs), ptr(d), ptr(bs)) - wantd[i]) <= 2e-6 * np.abs(wantd).max()
        y = np.zeros(kk, np.float32); sim.sim_q4_0n_row_dequant(ptr(K[i]), kk, ptr(y))
        assert np.array_equal(y, deq[i])


@pytest.mark.parametrize("t", [2, 6, 8, 14])
def test_numpy_model_of_the_library_layout_is_the_repack_kernels_permutation(sim, t):
    """refutil.repack_rows_np — the layout model every CPU test of the wide kernels is built on — against repacked_off (llama-box_b200/csrc/repack_layout.cuh),
    the function the hardware-verified repack kernel applies to every 2-byte unit of a row"""
    rng = np.random.default_rng(t)
    k = 2048
    W = rand_blocks(rng, t, 5, k)
    want = repack_rows_np(t, W, k)
    rb = row_bytes(t, k)
    for i in range(5):
        out = np.zeros(rb, np.uint8)
        sim.sim_repack_row(t, ptr(W[i]), ptr(out), k // BLOCK_ELEMS[t], rb)
        assert np.array_equal(out, want[i])
