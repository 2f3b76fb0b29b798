#!/usr/bin/env python3
"""Generate tests/golden/*.bin from the UNMODIFIED reference build (oracle/_ref, see oracle/Makefile).

Run in the build container (needs /root/reference to have been compiled into oracle/_ref):
    python oracle/make_golden.py
Each fixture is a bundle (tests/refutil.py format) holding the seeded inputs AND the output the
reference's ggml-cpu backend produced for them (through oracle/_ref/ref_ops, i.e. a real ggml graph
on the real CPU backend) or the reference function's direct result (quantisers, dequantisers,
vec_dot).  tests/test_oracle_golden.py replays them against oracle/liboracle.so everywhere;
tests/test_gpu_ops_golden.py replays them against the CUDA kernels on the GPU box.
Deterministic data: tests/test-quantize-fns.cpp:31-35 style cosines + seeded normals.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from refutil import (F16, F32, GOLDEN_DIR, I32, I64, Q4_0, Q8_0, Q8_K, ACT_TYPE, TYPE_NAME, WEIGHT_TYPES, cos_data, ptr,  # noqa: E402
                     ref, ref_quantize_weights, row_bytes, run_ref_op, write_bundle)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    base, cpu = ref()
    rng = np.random.default_rng(1234)

    # ---- activation quantisers + fp16 conversion (function-level KATs)
    k = 1024
    x = np.stack([cos_data(k, 0.0), (rng.standard_normal(k) * 3).astype(np.float32), np.zeros(k, np.float32)])
    x[1, 300] = -x[1].max() if False else x[1, 300]
    q80 = np.zeros((3, row_bytes(Q8_0, k)), np.uint8); q8k = np.zeros((3, row_bytes(Q8_K, k)), np.uint8)
    for i in range(3):
        cpu.quantize_row_q8_0(ptr(x[i]), ptr(q80[i]), k)
        cpu.quantize_row_q8_K(ptr(x[i]), ptr(q8k[i]), k)
    q8k[2] = 0  # all-zero rows: the reference leaves bsums untouched; pin them to zero
    xs = (rng.standard_normal(4096) * 10 ** rng.uniform(-8, 5, 4096)).astype(np.float32)
    h = np.zeros(xs.size, np.uint16); cpu.ggml_cpu_fp32_to_fp16(ptr(xs), ptr(h), xs.size)
    write_bundle(os.path.join(GOLDEN_DIR, "quantize_act.bin"), [("x", F32, [k, 3], x), ("q8_0", Q8_0, [k, 3], q80), ("q8_K", Q8_K, [k, 3], q8k),
                                                               ("f32", F32, [xs.size], xs), ("f16", F16, [xs.size], h)])

    # ---- weights: reference quantiser output, dequantised values, vec_dot and MUL_MAT through the CPU backend
    m, k, n = 24, 2048, 3
    w = (rng.standard_normal((m, k)) * 0.05).astype(np.float32)
    X = rng.standard_normal((n, k)).astype(np.float32)
    names = {2: "q4_0", 6: "q5_0", 8: "q8_0", 12: "q4_K", 13: "q5_K", 14: "q6_K"}
    for t in WEIGHT_TYPES:
        Wq = ref_quantize_weights(t, w)
        deq = np.zeros((m, k), np.float32)
        for i in range(m):
            getattr(base, "dequantize_row_" + names[t])(ptr(Wq[i]), ptr(deq[i]), k)
        _, _, out = run_ref_op("mul_mat", [("w", t, [k, m], Wq), ("x", F32, [k, n], X)])
        write_bundle(os.path.join(GOLDEN_DIR, f"mul_mat_{TYPE_NAME[t]}.bin"),
                     [("w", t, [k, m], Wq), ("x", F32, [k, n], X), ("deq", F32, [k, m], deq), ("dst", F32, [m, n], out)])

    # ---- rms_norm (+mul)
    xr = (rng.standard_normal((3, 2048)) * 2).astype(np.float32); wr = (1 + 0.1 * rng.standard_normal(2048)).astype(np.float32)
    _, _, out = run_ref_op("rms_norm", [("x", F32, [2048, 3], xr), ("w", F32, [2048], wr)], {"eps": 1e-5})
    write_bundle(os.path.join(GOLDEN_DIR, "rms_norm.bin"), [("x", F32, [2048, 3], xr), ("w", F32, [2048], wr), ("dst", F32, [2048, 3], out)])

    # ---- rope: llama-3 NORM with freq factors, qwen2 NEOX, YaRN
    for tag, prm, use_ff in (("norm_ff", dict(n_dims=128, mode=0, n_ctx_orig=8192, freq_base=500000.0, freq_scale=1.0, ext_factor=0.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0), True),
                             ("neox", dict(n_dims=128, mode=2, n_ctx_orig=32768, freq_base=1000000.0, freq_scale=1.0, ext_factor=0.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0), False),
                             ("yarn", dict(n_dims=64, mode=2, n_ctx_orig=4096, freq_base=10000.0, freq_scale=0.25, ext_factor=1.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0), False)):
        hd = prm["n_dims"]; nh, nt = 4, 5
        xx = rng.standard_normal((nt, nh, hd)).astype(np.float32); pos = np.array([0, 1, 63, 4095, 20000], np.int32)
        ff = rng.uniform(1, 8, hd // 2).astype(np.float32)
        T = [("x", F32, [hd, nh, nt], xx), ("pos", I32, [nt], pos)] + ([("ff", F32, [hd // 2], ff)] if use_ff else [])
        _, _, out = run_ref_op("rope", T, prm)
        pv = np.array([prm[k2] for k2 in ("n_dims", "mode", "n_ctx_orig", "freq_base", "freq_scale", "ext_factor", "attn_factor", "beta_fast", "beta_slow")], np.float32)
        write_bundle(os.path.join(GOLDEN_DIR, f"rope_{tag}.bin"), T + [("params", F32, [9], pv), ("dst", F32, [hd, nh, nt], out)])

    # ---- set_rows into F16 / Q8_0 caches
    for dt in (F16, Q8_0):
        nc, nr, tot = 512, 3, 8
        src = (rng.standard_normal((nr, nc)) * 2).astype(np.float32); ids = np.array([6, 1, 3], np.int64)
        cache = np.zeros((tot, row_bytes(dt, nc)), np.uint8)
        _, _, out = run_ref_op("set_rows", [("cache", dt, [nc, tot], cache), ("src", F32, [nc, nr], src), ("ids", I64, [nr], ids)])
        write_bundle(os.path.join(GOLDEN_DIR, f"set_rows_{TYPE_NAME[dt]}.bin"), [("src", F32, [nc, nr], src), ("ids", I64, [nr], ids), ("dst", dt, [nc, tot], out)])

    # ---- flash_attn_ext over F16 / Q8_0 caches (GQA 4:1, causal-ish mask, 2 tokens)
    for kvt in (F16, Q8_0):
        dk, nh, nhkv, nt, nkv = 128, 8, 2, 2, 256
        q = rng.standard_normal((nt, nh, dk)).astype(np.float32)
        kf = rng.standard_normal((nkv, nhkv * dk)).astype(np.float32); vf = rng.standard_normal((nkv, nhkv * dk)).astype(np.float32)
        z = np.zeros((nkv, row_bytes(kvt, nhkv * dk)), np.uint8); allid = np.arange(nkv, dtype=np.int64)
        _, _, kc = run_ref_op("set_rows", [("cache", kvt, [nhkv * dk, nkv], z), ("src", F32, [nhkv * dk, nkv], kf), ("ids", I64, [nkv], allid)])
        _, _, vc = run_ref_op("set_rows", [("cache", kvt, [nhkv * dk, nkv], z), ("src", F32, [nhkv * dk, nkv], vf), ("ids", I64, [nkv], allid)])
        mask = np.full((64, nkv), -np.inf, np.float32); mask[0, :200] = 0; mask[1, :201] = 0
        m16 = mask.astype(np.float16)
        prm = dict(dk=dk, dv=dk, n_head_kv=nhkv, n_kv=nkv, scale=float(1 / np.sqrt(dk)))
        T = [("q", F32, [dk, nh, nt], q), ("k", kvt, [nhkv * dk, nkv], kc), ("v", kvt, [nhkv * dk, nkv], vc), ("mask", F16, [nkv, 64], m16)]
        _, _, out = run_ref_op("flash_attn", T, prm)
        write_bundle(os.path.join(GOLDEN_DIR, f"flash_attn_{TYPE_NAME[kvt]}.bin"), T + [("dst", F32, [dk, nh, nt], out)])

    # ---- swiglu
    g = (rng.standard_normal((2, 512)) * 3).astype(np.float32); u = rng.standard_normal((2, 512)).astype(np.float32)
    _, _, out = run_ref_op("swiglu", [("gate", F32, [512, 2], g), ("up", F32, [512, 2], u)])
    write_bundle(os.path.join(GOLDEN_DIR, "swiglu.bin"), [("gate", F32, [512, 2], g), ("up", F32, [512, 2], u), ("dst", F32, [512, 2], out)])
    # ---- the wide path's formats: reference quantiser output, de-quantised values, MUL_MAT through the CPU backend
    from refutil import EXT_TYPES, Q4_K, Q8_1
    rng = np.random.default_rng(4321)                # its own stream: the fixtures above stay byte-identical
    m2, k2, n2 = 8, 1024, 2
    w2 = (rng.standard_normal((m2, k2)) * 0.05).astype(np.float32); X2 = rng.standard_normal((n2, k2)).astype(np.float32)
    for t in EXT_TYPES:
        Wq = ref_quantize_weights(t, w2)
        deq = np.zeros((m2, k2), np.float32)
        for i in range(m2):
            getattr(base, "dequantize_row_" + TYPE_NAME[t])(ptr(Wq[i]), ptr(deq[i]), k2)
        _, _, out = run_ref_op("mul_mat", [("w", t, [k2, m2], Wq), ("x", F32, [k2, n2], X2)])
        write_bundle(os.path.join(GOLDEN_DIR, f"ext_mul_mat_{TYPE_NAME[t]}.bin"),
                     [("w", t, [k2, m2], Wq), ("x", F32, [k2, n2], X2), ("deq", F32, [k2, m2], deq), ("dst", F32, [m2, n2], out)])
    kx = x.shape[1]                                   # the activation rows of quantize_act.bin
    q81 = np.zeros((3, row_bytes(Q8_1, kx)), np.uint8)
    for i in range(3):
        cpu.quantize_row_q8_1(ptr(x[i]), ptr(q81[i]), kx)
    # ---- MUL_MAT_ID (f2) as build_moe_ffn emits it: shared activation and one activation per used expert; GET_ROWS on a quantised table (f4)
    me, ke, n_expert, n_used, n_tok = 16, 512, 5, 2, 3
    We = ref_quantize_weights(Q4_K, (rng.standard_normal((n_expert * me, ke)) * 0.05).astype(np.float32))
    ids = np.stack([rng.permutation(n_expert)[:n_used] for _ in range(n_tok)]).astype(np.int32)
    T = [("w", Q4_K, [ke, me, n_expert], We), ("ids", I32, [n_used, n_tok], ids), ("q8_1", Q8_1, [kx, 3], q81), ("x81", F32, [kx, 3], x)]
    for tag, n_b1 in (("shared", 1), ("per_expert", n_used)):
        b = rng.standard_normal((n_tok, n_b1, ke)).astype(np.float32)
        _, _, out = run_ref_op("mul_mat_id", [("w", Q4_K, [ke, me, n_expert], We), ("x", F32, [ke, n_b1, n_tok], b), ("ids", I32, [n_used, n_tok], ids)])
        T += [("b_" + tag, F32, [ke, n_b1, n_tok], b), ("dst_" + tag, F32, [me, n_used, n_tok], out)]
    gid = np.array([7, 0, 79, 7], np.int32)
    _, _, out = run_ref_op("get_rows", [("src", Q4_K, [ke, n_expert * me], We), ("ids", I32, [gid.size], gid)])
    T += [("gr_ids", I32, [gid.size], gid), ("gr_dst", F32, [ke, gid.size], out)]
    write_bundle(os.path.join(GOLDEN_DIR, "ext_moe_get_rows.bin"), T)

    print("golden fixtures:", sorted(os.listdir(GOLDEN_DIR)))


if __name__ == "__main__":
    main()
