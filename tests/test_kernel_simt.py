"""CPU: the SOURCE of the wide path's CUDA kernels (llama-box_b200/csrc/mmvq_ext_kernels.cuh, fattn_ext_kernels.cuh, with the warp quantisers
of actquant.cuh / actquant_ext.cuh) executed under a small SIMT emulation (tests/hostsim/simt.h: one OS thread per CUDA thread, barriers for
__syncthreads and warp shuffles) and compared with the C oracle.  This exercises what tests/test_extfmt_hostsim.py cannot: the shared-memory
layout, the in-kernel activation quantisers, row / sub-block loop bounds and tails, warp reductions, MUL_MAT_ID indexing, the softmax merge of the
q4_0 attention kernel.  What remains for the GPU: real memory-model / alignment behaviour and launch plumbing (tests/test_gpu_zz_wide.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from refutil import (BLOCK_ELEMS, EXT_TYPES, Q4_0, Q6_K, ROOT, WEIGHT_TYPES, oracle, orc_mul_mat, ptr, rand_blocks, repack_rows_np, row_bytes)

SIM_DIR = os.path.join(ROOT, "tests", "hostsim")
CSRC = os.path.join(ROOT, "llama-box_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not installed")


@pytest.fixture(scope="module")
def K():
    so = os.path.join(SIM_DIR, "libkernsim.so")
    deps = [os.path.join(SIM_DIR, f) for f in ("kernsim.cpp", "simt.h")] + [os.path.join(CSRC, f) for f in
            ("mmvq_ext_kernels.cuh", "fattn_ext_kernels.cuh", "glue_ext_kernels.cuh", "extfmt.cuh", "actquant.cuh", "actquant_ext.cuh", "common.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w", "-I" + CUDA_INC, "-x", "c++",
                               os.path.join(SIM_DIR, "kernsim.cpp"), "-o", so])
    L = C.CDLL(so)
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    L.sim_mul_mat_vec_wide.argtypes = [i32, vp, vp, i64, vp, i64, vp, vp, i64, i64, i64, i32]
    L.sim_mul_mat_id.argtypes = [i32, vp, i64, vp, i64, i64, i64, vp, i64, vp, i64, i64, i64, i64, i64, i64, i64, i32]
    L.sim_get_rows_q.argtypes = [i32, vp, i64, i64, vp, vp, i64, i64, i64]
    L.sim_set_rows_q4_0.argtypes = [vp, i64, vp, vp, i64, i64, i64]
    L.sim_misaligned.restype = C.c_long
    L.sim_unary.argtypes = [i32, vp, vp, i64, f32, f32]
    L.sim_flash_attn_any.argtypes = [i32, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, vp, i64, i64, i64, i64, i64, f32, f32, f32]
    L.sim_binary_strided.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sim_soft_max_rows.argtypes = [vp, i64, vp, i64, i64, i64, f32]
    L.sim_argsort_rows.argtypes = [vp, i64, vp, i64, i64, i64, i32]
    L.sim_sum_rows.argtypes = [vp, i64, vp, i64, i64]
    L.sim_get_rows_f32_batched.argtypes = [vp, i64, i64, i64, vp, i64, vp, i64, i64, i64, i64, i64]
    L.sim_mul_mat_f32.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i64]
    L.sim_soft_max_mask.argtypes = [vp, vp, vp, i32, i64, i64, i64, i64, f32, f32]
    L.sim_mul_mat_f16.argtypes = [vp, i64, i64, i64, vp, i64, i64, vp, i64, i64, i64, i64, i64, i64]
    L.sim_scatter_rows1.argtypes = [vp, vp, vp, i32, i64, i64]
    L.sim_flash_attn_q4_0.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, vp, i64, i64, i64, i64, i64, f32, f32, f32]
    return L


@pytest.fixture(autouse=True)
def no_misaligned_loads(K):
    """weight loads the kernels issue as 2- / 4-byte accesses must be naturally aligned (a GPU traps on them; x86 would not)"""
    before = K.sim_misaligned()
    yield
    assert K.sim_misaligned() == before


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("t", WEIGHT_TYPES + EXT_TYPES)
def test_wide_matvec_kernel(K, t):
    rng = np.random.default_rng(t)
    m, k, n = 20, 512, 2                                         # 20 rows over 2 row groups of 8 warps: a second, partial pass of the row loop
    W = rand_blocks(rng, t, m, k); X = (rng.standard_normal((n, k)) * 3).astype(np.float32)
    X[1, 256:288] = 0                                            # an all-zero activation block
    bias = rng.standard_normal(m).astype(np.float32); res = rng.standard_normal((n, m)).astype(np.float32)
    want = orc_mul_mat(t, W, X, m, n, k)
    Wl = repack_rows_np(t, W, k)
    dst = np.full((n, m), 7.0, np.float32)
    assert K.sim_mul_mat_vec_wide(t, ptr(Wl), ptr(X), k, ptr(dst), m, None, None, m, k, n, 2) == 1
    assert rel(dst, want) <= 2e-5
    assert K.sim_mul_mat_vec_wide(t, ptr(Wl), ptr(X), k, ptr(dst), m, ptr(bias), ptr(res), m, k, n, 2) == 1
    assert rel(dst, want + bias[None, :] + res) <= 2e-5


@pytest.mark.parametrize("t", [12, 3, 11])
def test_wide_matvec_kernel_column_groups(K, t):
    """11 columns = one CTA group of 8 and one of 3: the columns of a group share a pass over the weights"""
    rng = np.random.default_rng(t)
    m, k, n = 10, 512, 11
    W = rand_blocks(rng, t, m, k); X = rng.standard_normal((n, k)).astype(np.float32)
    res = rng.standard_normal((n, m)).astype(np.float32)
    want = orc_mul_mat(t, W, X, m, n, k) + res
    dst = np.full((n, m), 7.0, np.float32)
    K.sim_mul_mat_vec_wide(t, ptr(repack_rows_np(t, W, k)), ptr(X), k, ptr(dst), m, None, ptr(res), m, k, n, 1)
    assert rel(dst, want) <= 2e-5


@pytest.mark.parametrize("t", [3, 7, 20, 39])
def test_wide_matvec_kernel_rows_not_a_multiple_of_256(K, t):
    rng = np.random.default_rng(t)
    m, k = 9, 29568 // 4                                         # 7392 = 231 blocks: the last 256-chunk of the prologue is partial
    W = rand_blocks(rng, t, m, k); X = rng.standard_normal((1, k)).astype(np.float32)
    want = orc_mul_mat(t, W, X, m, 1, k)
    dst = np.zeros((1, m), np.float32)
    K.sim_mul_mat_vec_wide(t, ptr(W), ptr(X), k, ptr(dst), m, None, None, m, k, 1, 1)
    assert rel(dst, want) <= 2e-5


@pytest.mark.parametrize("t", [2, 8, 12, 14, 3, 11, 23])
@pytest.mark.parametrize("shared", [True, False])
def test_mul_mat_id_kernel(K, t, shared):
    rng = np.random.default_rng(10 * t + shared)
    m, k, n_expert, n_used, n_tok = 12, 512, 4, 2, 3
    n_b1 = 1 if shared else n_used
    W = rand_blocks(rng, t, n_expert * m, k)
    b = rng.standard_normal((n_tok, n_b1, k)).astype(np.float32)
    ids_stride = n_expert                                        # ids as llama's top-k VIEW of the argsort result: row stride = n_expert
    ids = np.full((n_tok, ids_stride), -5, np.int32)
    for tk in range(n_tok):
        ids[tk, :n_used] = rng.permutation(n_expert)[:n_used]
    ids[2, 1] = n_expert + 3                                     # out of range: that output block must stay untouched
    want = np.full((n_tok, n_used, m), 7.0, np.float32)
    oracle().orc_mul_mat_id(t, ptr(W), ptr(b), ptr(ids), ptr(want), m, k, n_expert, n_used, n_tok, n_b1, ids_stride)
    Wl = repack_rows_np(t, W, k)
    dst = np.full((n_tok, n_used, m), 7.0, np.float32)
    assert K.sim_mul_mat_id(t, ptr(Wl), m * row_bytes(t, k), ptr(b), n_b1 * k, k, n_b1, ptr(ids), ids_stride, ptr(dst), n_used * m, m, m, k, n_expert, n_used, n_tok, 1) == 1
    assert np.all(dst[2, 1] == 7.0)
    assert rel(dst, want) <= 2e-5


@pytest.mark.parametrize("t,ncols", [(t, 512) for t in WEIGHT_TYPES + EXT_TYPES] + [(12, 8192), (3, 8192 + 32)])
def test_get_rows_kernel_bit_exact(K, t, ncols):
    rng = np.random.default_rng(t + ncols)
    if BLOCK_ELEMS[t] == 256:
        ncols = ncols // 256 * 256
    nrows = 9
    W = rand_blocks(rng, t, nrows, ncols)
    ids = np.array([8, 0, 3, 8, 100, -1], np.int32)              # the last two are out of range: zeros
    want = np.zeros((ids.size, ncols), np.float32)
    oracle().orc_get_rows_q(t, ptr(W), ptr(ids[:4]), ptr(want), ncols, 4)
    dst = np.full((ids.size, ncols), 7.0, np.float32)
    assert K.sim_get_rows_q(t, ptr(repack_rows_np(t, W, ncols)), row_bytes(t, ncols), nrows, ptr(ids), ptr(dst), ncols, ncols, ids.size) == 1
    assert np.array_equal(dst, want)


@pytest.mark.parametrize("d,n_head,n_head_kv,n_tok,n_kv,max_bias,softcap", [(128, 4, 2, 2, 80, 0.0, 0.0), (64, 4, 4, 1, 37, 8.0, 0.0), (128, 2, 1, 1, 5, 0.0, 30.0)])
def test_q4_0_kv_cache_kernels(K, d, n_head, n_head_kv, n_tok, n_kv, max_bias, softcap):
    rng = np.random.default_rng(d + n_kv)
    rb_row, rb_head = row_bytes(Q4_0, n_head_kv * d), row_bytes(Q4_0, d)
    kf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32); vf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32)
    kf[1, :32] = 0
    ids = rng.permutation(n_kv).astype(np.int64)
    kc = np.zeros((n_kv, rb_row), np.uint8); vc = kc.copy(); kc2 = kc.copy(); vc2 = kc.copy()
    oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), Q4_0, n_head_kv * d, n_kv, rb_row)
    oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), Q4_0, n_head_kv * d, n_kv, rb_row)
    K.sim_set_rows_q4_0(ptr(kf), n_head_kv * d, ptr(ids), ptr(kc2), rb_row, n_head_kv * d, n_kv)
    K.sim_set_rows_q4_0(ptr(vf), n_head_kv * d, ptr(ids), ptr(vc2), rb_row, n_head_kv * d, n_kv)
    assert np.array_equal(kc2, kc) and np.array_equal(vc2, vc)
    q = rng.standard_normal((n_tok, n_head, d)).astype(np.float32)
    mask = np.full((64, n_kv), -np.inf, np.float32)
    for t in range(n_tok):
        mask[t, :max(1, n_kv - n_tok + t + 1 - 3)] = rng.uniform(-1, 0, max(1, n_kv - n_tok + t + 1 - 3)) if max_bias > 0 else 0
    m16 = mask.astype(np.float16)
    scale = float(1 / np.sqrt(d))
    want = np.zeros((n_tok, n_head, d), np.float32)
    oracle().orc_flash_attn_ext(ptr(q), n_head * d * 4, d * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), ptr(want),
                                Q4_0, d, d, n_head, n_head_kv, n_tok, n_kv, scale, max_bias, softcap)
    dst = np.zeros_like(want)
    K.sim_flash_attn_q4_0(ptr(q), n_head * d, d, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), n_kv, ptr(dst),
                          d, n_head, n_head_kv, n_tok, n_kv, scale, max_bias, softcap)
    assert rel(dst, want) <= 2e-5


def i64a(*v):
    return np.array(v, np.int64)


def test_moe_router_glue_kernels(K):
    """the router of build_moe_ffn, op by op: f32 router matmul, SOFT_MAX, ARGSORT (ties!), batched GET_ROWS of the selected probabilities, SUM_ROWS,
    DIV, the broadcast MUL by the expert weights and the ADD over strided expert slices"""
    rng = np.random.default_rng(9)
    E, NE, NU, NT = 96, 8, 2, 5
    Wg = rng.standard_normal((NE, E)).astype(np.float32); x = rng.standard_normal((NT, E)).astype(np.float32)
    logits = np.zeros((NT, NE), np.float32)
    K.sim_mul_mat_f32(ptr(Wg), E, ptr(x), E, ptr(logits), NE, NE, E, NT)
    want = (x.astype(np.float64) @ Wg.astype(np.float64).T)
    assert np.abs(logits - want).max() <= 1e-5 * np.abs(want).max()
    logits[2, 3] = logits[2, 6]                                  # a tie
    probs = np.zeros_like(logits); wantp = np.zeros_like(logits)
    K.sim_soft_max_rows(ptr(logits), NE, ptr(probs), NE, NE, NT, 1.0); oracle().orc_soft_max_rows(ptr(logits), ptr(wantp), NE, NT, 1.0)
    assert np.abs(probs - wantp).max() <= 1e-6
    for desc in (1, 0):
        idx = np.zeros((NT, NE), np.int32); wanti = np.zeros_like(idx)
        K.sim_argsort_rows(ptr(wantp), NE, ptr(idx), NE, NE, NT, desc); oracle().orc_argsort_rows(ptr(wantp), ptr(wanti), NE, NT, desc)
        assert np.array_equal(idx, wanti)
    K.sim_argsort_rows(ptr(wantp), NE, ptr(idx), NE, NE, NT, 1)
    # selected experts = the first NU columns of the argsort result, as a VIEW (row stride NE): ids_batch_stride = NE
    w = np.full((NT, NU), 7.0, np.float32)
    K.sim_get_rows_f32_batched(ptr(wantp), 1, NE, NE, ptr(idx), NE, ptr(w), 1, NU, 1, NU, NT)
    assert np.array_equal(w, np.take_along_axis(wantp, idx[:, :NU], axis=1))
    s = np.zeros(NT, np.float32); wants = np.zeros(NT, np.float32)
    K.sim_sum_rows(ptr(w), NU, ptr(s), NU, NT); oracle().orc_sum_rows(ptr(w), ptr(wants), NU, NT)
    assert np.array_equal(s, wants)
    wn = np.zeros_like(w)                                        # DIV: [NU, NT] / [1, NT]
    K.sim_binary_strided(2, ptr(w), ptr(i64a(4, 4 * NU, 4 * NU * NT, 4 * NU * NT)), ptr(s), ptr(i64a(1, NT, 1, 1)), ptr(i64a(4, 4, 4 * NT, 4 * NT)), ptr(wn),
                         ptr(i64a(NU, NT, 1, 1)), ptr(i64a(4, 4 * NU, 4 * NU * NT, 4 * NU * NT)))
    assert np.array_equal(wn, w / s[:, None])
    ex = rng.standard_normal((NT, NU, E)).astype(np.float32)      # experts [E, NU, NT] * weights [1, NU, NT]
    exw = np.zeros_like(ex)
    nb3 = i64a(4, 4 * E, 4 * E * NU, 4 * E * NU * NT)
    K.sim_binary_strided(1, ptr(ex), ptr(nb3), ptr(wn), ptr(i64a(1, NU, NT, 1)), ptr(i64a(4, 4, 4 * NU, 4 * NU * NT)), ptr(exw), ptr(i64a(E, NU, NT, 1)), ptr(nb3))
    assert np.array_equal(exw, ex * wn[:, :, None])
    out = np.zeros((NT, E), np.float32)                          # ADD of the two expert slices: views [E, NT] with row stride E * NU
    sl = i64a(4, 4 * E * NU, 4 * E * NU * NT, 4 * E * NU * NT)
    K.sim_binary_strided(0, ptr(exw), ptr(sl), C.c_void_p(exw.ctypes.data + 4 * E), ptr(i64a(E, NT, 1, 1)), ptr(sl), ptr(out), ptr(i64a(E, NT, 1, 1)), ptr(i64a(4, 4 * E, 4 * E * NT, 4 * E * NT)))
    assert np.array_equal(out, exw[:, 0, :] + exw[:, 1, :])


def test_non_flash_attention_kernels(K):
    """attention without -fa, as build_attn_mha emits it: V stored transposed by an element-wise SET_ROWS, KQ = K^T Q over a permuted f16 view of the K cache
    with GQA broadcast, SOFT_MAX with mask (+ ALiBi), KQV over the transposed V cache, CONT of the permuted result"""
    rng = np.random.default_rng(21)
    hd, n_kv, kv_size, nt, nh, nhk = 64, 40, 48, 3, 4, 2
    E = nhk * hd
    # K cache rows [kv_size][nhk*hd] f16; V cache TRANSPOSED: element (cell, d, head) at (head*hd + d) * kv_size + cell
    Kc = rng.standard_normal((kv_size, E)).astype(np.float16)
    vcur = rng.standard_normal((n_kv, E)).astype(np.float32)
    Vt = np.zeros(E * kv_size, np.float16)
    cells = rng.permutation(kv_size)[:n_kv]
    ids = (np.arange(E, dtype=np.int64)[None, :] * kv_size + cells[:, None]).reshape(-1)          # v_idxs of llama-kv-cache-unified.cpp for v_trans
    K.sim_scatter_rows1(ptr(vcur), ptr(ids), ptr(Vt), 1, ids.size, Vt.size)
    want_vt = np.zeros(E * kv_size, np.float16); want_vt[ids] = vcur.reshape(-1).astype(np.float16)
    assert np.array_equal(Vt.view(np.uint16), want_vt.view(np.uint16))
    Q = rng.standard_normal((nt, nh, hd)).astype(np.float32)                                     # [hd, n_head, n_tok]; viewed permuted as [hd, n_tok, n_head]
    kq = np.zeros((nh, nt, kv_size), np.float32); wkq = np.zeros_like(kq)
    args = (ptr(Kc), E * 2, hd * 2, nhk, ptr(Q), nh * hd * 4, hd * 4)
    K.sim_mul_mat_f16(*args, ptr(kq), kv_size * 4, nt * kv_size * 4, kv_size, nt, nh, hd)
    oracle().orc_mul_mat_f16(*args, ptr(wkq), kv_size * 4, nt * kv_size * 4, kv_size, nt, nh, hd)
    assert rel(kq, wkq) <= 2e-6
    for is_f16, max_bias in ((0, 0.0), (1, 8.0)):
        mask = np.full((64, kv_size), -np.inf, np.float32)
        for t in range(nt):
            mask[t, cells[:n_kv - nt + t + 1]] = rng.uniform(-1, 0) if max_bias > 0 else 0
        m = mask.astype(np.float16) if is_f16 else mask
        p = np.zeros_like(kq); wp = np.zeros_like(kq)
        K.sim_soft_max_mask(ptr(wkq), ptr(p), ptr(m), is_f16, kv_size, kv_size, nt, nh, 0.125, max_bias)
        oracle().orc_soft_max_mask(ptr(wkq), ptr(wp), ptr(m), is_f16, kv_size, kv_size, nt, nh, 0.125, max_bias)
        assert np.abs(p - wp).max() <= 1e-6
    kqv = np.zeros((nh, nt, hd), np.float32); wkqv = np.zeros_like(kqv)                           # V view [kv_size, hd, nhk]: nb1 = kv_size*2, nb2 = kv_size*hd*2
    args = (ptr(want_vt), kv_size * 2, kv_size * hd * 2, nhk, ptr(wp), kv_size * 4, nt * kv_size * 4)
    K.sim_mul_mat_f16(*args, ptr(kqv), hd * 4, nt * hd * 4, hd, nt, nh, kv_size)
    oracle().orc_mul_mat_f16(*args, ptr(wkqv), hd * 4, nt * hd * 4, hd, nt, nh, kv_size)
    assert rel(kqv, wkqv) <= 2e-6
    out = np.zeros((nt, nh, hd), np.float32)                                                      # CONT of permute(kqv, 0, 2, 1, 3): [hd, n_head, n_tok]
    src_nb = i64a(4, nt * hd * 4, hd * 4, nh * nt * hd * 4)
    K.sim_binary_strided(3, ptr(wkqv), ptr(src_nb), ptr(wkqv), ptr(i64a(1, 1, 1, 1)), ptr(i64a(4, 4, 4, 4)), ptr(out), ptr(i64a(hd, nh, nt, 1)), ptr(i64a(4, hd * 4, nh * hd * 4, nt * nh * hd * 4)))
    assert np.array_equal(out, wkqv.transpose(1, 0, 2))


@pytest.mark.parametrize("kvt,d", [(1, 256), (1, 96), (8, 256), (8, 80 + 16), (2, 192), (8, 32), (1, 64)])
def test_flash_attn_any_head_size_kernel(K, kvt, d):
    """head sizes the tuned kernels do not carry (Gemma 256, Phi 96 ...) over F16 / Q8_0 / Q4_0 caches, GQA, mask, soft-cap"""
    from refutil import F16, Q8_0
    rng = np.random.default_rng(kvt * 1000 + d)
    n_head, n_head_kv, n_tok, n_kv = 4, 2, 2, 45
    rb_row, rb_head = row_bytes(kvt, n_head_kv * d), row_bytes(kvt, d)
    kf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32); vf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32)
    ids = np.arange(n_kv, dtype=np.int64)
    kc = np.zeros((n_kv, rb_row), np.uint8); vc = kc.copy()
    oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), kvt, n_head_kv * d, n_kv, rb_row)
    oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), kvt, n_head_kv * d, n_kv, rb_row)
    q = rng.standard_normal((n_tok, n_head, d)).astype(np.float32)
    mask = np.full((64, n_kv), -np.inf, np.float32); mask[0, :40] = 0; mask[1, :41] = 0
    m16 = mask.astype(np.float16)
    scale, softcap = float(1 / np.sqrt(d)), (20.0 if d == 256 else 0.0)
    want = np.zeros((n_tok, n_head, d), np.float32)
    oracle().orc_flash_attn_ext(ptr(q), n_head * d * 4, d * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), ptr(want),
                                kvt, d, d, n_head, n_head_kv, n_tok, n_kv, scale, 0.0, softcap)
    dst = np.zeros_like(want)
    K.sim_flash_attn_any(kvt, ptr(q), n_head * d, d, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), n_kv, ptr(dst),
                         d, n_head, n_head_kv, n_tok, n_kv, scale, 0.0, softcap)
    if kvt == 1:
        # F16 V: the oracle accumulates in fp16 (ops.cpp:8278-8340: ~2e-3 of noise), the kernel in f32 — compare with the f64 value of the same f16 inputs
        Kd = kc.view(np.float16).reshape(n_kv, n_head_kv, d).astype(np.float64); Vd = vc.view(np.float16).reshape(n_kv, n_head_kv, d).astype(np.float64)
        exact = np.zeros((n_tok, n_head, d))
        for t in range(n_tok):
            for h in range(n_head):
                sc = Kd[:, h // (n_head // n_head_kv), :] @ q[t, h].astype(np.float16).astype(np.float64) * scale
                if softcap:
                    sc = softcap * np.tanh(sc / softcap)
                sc = sc + mask[t].astype(np.float64)
                w = np.exp(sc - sc.max()); w /= w.sum()
                exact[t, h] = w @ Vd[:, h // (n_head // n_head_kv), :]
        assert rel(dst, exact) <= 2e-5 and rel(want, exact) <= 4e-3
    else:
        assert rel(dst, want) <= 2e-5


def test_unary_kernels(K):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(1000) * 5).astype(np.float32)
    for op, sc, b, tol in ((0, 0.37, 0.0, 0.0), (0, 2.5, -0.75, 0.0), (1, 0, 0, 1e-6), (2, 0, 0, 1e-7)):
        y = np.zeros_like(x); want = np.zeros_like(x)
        K.sim_unary(op, ptr(x), ptr(y), x.size, sc, b); oracle().orc_unary(op, ptr(x), ptr(want), x.size, sc, b)
        assert np.abs(y - want).max() <= tol * max(1.0, np.abs(want).max()), op


def test_wide_matvec_kernel_random_shapes(K):
    """seeded sweep over shapes the fixed cases do not hit: one row, fewer rows than warps, odd row counts over several row groups, 1..9 columns
    (column-group tails), k from one block up to several 256-chunks with a partial last chunk"""
    rng = np.random.default_rng(2026)
    for it in range(24):
        t = int(rng.choice(WEIGHT_TYPES + EXT_TYPES))
        be = BLOCK_ELEMS[t]
        if t in EXT_TYPES and be == 32:
            k = 32 * int(rng.integers(1, 40))
        elif t == Q6_K:
            k = 512 * int(rng.integers(1, 3))
        else:
            k = 256 * int(rng.integers(1, 4))
        m = int(rng.choice([1, 3, 7, 8, 9, 17, 33])); n = int(rng.integers(1, 10)); gx = int(rng.integers(1, 4))
        W = rand_blocks(rng, t, m, k); X = rng.standard_normal((n, k)).astype(np.float32)
        bias = rng.standard_normal(m).astype(np.float32) if it % 2 else None
        want = orc_mul_mat(t, W, X, m, n, k) + (bias[None, :] if bias is not None else 0)
        dst = np.full((n, m), 7.0, np.float32)
        K.sim_mul_mat_vec_wide(t, ptr(repack_rows_np(t, W, k)), ptr(X), k, ptr(dst), m, ptr(bias), None, m, k, n, gx)
        assert rel(dst, want) <= 2e-5, (t, m, k, n, gx)


def test_mul_mat_id_and_q4_0_attention_random_shapes(K):
    """seeded sweep: MUL_MAT_ID with 1..5 tokens, 1..4 used experts, activation sharing patterns n_b1 | n_used; q4_0 attention with GQA ratios 1 / 2 / 4,
    1..3 tokens, KV lengths that are not multiples of the warp count"""
    rng = np.random.default_rng(77)
    for it in range(10):
        t = int(rng.choice([2, 12, 13, 3, 7, 10, 23, 39]))
        k = 256 * int(rng.integers(1, 3)); m = int(rng.choice([4, 12, 20]))
        n_expert = int(rng.integers(2, 7)); n_used = int(rng.integers(1, min(4, n_expert) + 1)); n_tok = int(rng.integers(1, 6))
        n_b1 = int(rng.choice([d for d in range(1, n_used + 1) if n_used % d == 0]))
        W = rand_blocks(rng, t, n_expert * m, k)
        b = rng.standard_normal((n_tok, n_b1, k)).astype(np.float32)
        ids = np.stack([rng.permutation(n_expert)[:n_used] for _ in range(n_tok)]).astype(np.int32)
        want = np.zeros((n_tok, n_used, m), np.float32)
        oracle().orc_mul_mat_id(t, ptr(W), ptr(b), ptr(ids), ptr(want), m, k, n_expert, n_used, n_tok, n_b1, n_used)
        dst = np.zeros_like(want)
        K.sim_mul_mat_id(t, ptr(repack_rows_np(t, W, k)), m * row_bytes(t, k), ptr(b), n_b1 * k, k, n_b1, ptr(ids), n_used, ptr(dst), n_used * m, m, m, k, n_expert, n_used, n_tok, 1)
        assert rel(dst, want) <= 2e-5, (t, m, k, n_expert, n_used, n_tok, n_b1)
    for it in range(6):
        d = int(rng.choice([64, 128])); n_head_kv = int(rng.choice([1, 2])); n_head = n_head_kv * int(rng.choice([1, 2, 4])); n_tok = int(rng.integers(1, 4)); n_kv = int(rng.integers(3, 70))
        rb_row, rb_head = row_bytes(Q4_0, n_head_kv * d), row_bytes(Q4_0, d)
        kf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32); vf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32)
        ids = np.arange(n_kv, dtype=np.int64)
        kc = np.zeros((n_kv, rb_row), np.uint8); vc = kc.copy()
        oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), Q4_0, n_head_kv * d, n_kv, rb_row)
        oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), Q4_0, n_head_kv * d, n_kv, rb_row)
        q = rng.standard_normal((n_tok, n_head, d)).astype(np.float32)
        use_mask = it % 2 == 0
        mask = np.full((64, n_kv), -np.inf, np.float32)
        for tt in range(n_tok):
            mask[tt, :max(1, n_kv - n_tok + tt + 1)] = 0
        m16 = mask.astype(np.float16)
        want = np.zeros((n_tok, n_head, d), np.float32)
        oracle().orc_flash_attn_ext(ptr(q), n_head * d * 4, d * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)) if use_mask else None, ptr(want),
                                    Q4_0, d, d, n_head, n_head_kv, n_tok, n_kv, 0.1, 0.0, 0.0)
        dst = np.zeros_like(want)
        K.sim_flash_attn_q4_0(ptr(q), n_head * d, d, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)) if use_mask else None, n_kv, ptr(dst),
                              d, n_head, n_head_kv, n_tok, n_kv, 0.1, 0.0, 0.0)
        assert rel(dst, want) <= 2e-5, (d, n_head, n_head_kv, n_tok, n_kv, use_mask)
