"""Shared test helpers (TEST INFRASTRUCTURE): bundle I/O, ctypes handles for the C oracle
(oracle/liboracle.so) and — when it was built in this container — the unmodified reference
(oracle/_ref/*.so), and generators for valid GGUF quant blocks.

Block layouts follow /root/reference/llama.cpp/ggml/src/ggml-common.h:170-175,219-224,295-344.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# enum ggml_type values (ggml/include/ggml.h:377-418)
F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K = 0, 1, 2, 6, 8, 12, 13, 14, 15
Q4_1, Q5_1, Q8_1, Q2_K, Q3_K, IQ4_NL, IQ4_XS, MXFP4 = 3, 7, 9, 10, 11, 20, 23, 39      # the wide kernels' formats
I32, I64 = 26, 27
TYPE_NAME = {F32: "f32", F16: "f16", Q4_0: "q4_0", Q5_0: "q5_0", Q8_0: "q8_0", Q4_K: "q4_K", Q5_K: "q5_K", Q6_K: "q6_K", Q8_K: "q8_K",
             Q4_1: "q4_1", Q5_1: "q5_1", Q8_1: "q8_1", Q2_K: "q2_K", Q3_K: "q3_K", IQ4_NL: "iq4_nl", IQ4_XS: "iq4_xs", MXFP4: "mxfp4"}
BLOCK_ELEMS = {F32: 1, F16: 1, Q4_0: 32, Q5_0: 32, Q8_0: 32, Q4_K: 256, Q5_K: 256, Q6_K: 256, Q8_K: 256, I32: 1, I64: 1,
               Q4_1: 32, Q5_1: 32, Q8_1: 32, Q2_K: 256, Q3_K: 256, IQ4_NL: 32, IQ4_XS: 256, MXFP4: 32}
BLOCK_BYTES = {F32: 4, F16: 2, Q4_0: 18, Q5_0: 22, Q8_0: 34, Q4_K: 144, Q5_K: 176, Q6_K: 210, Q8_K: 292, I32: 4, I64: 8,
               Q4_1: 20, Q5_1: 24, Q8_1: 36, Q2_K: 84, Q3_K: 110, IQ4_NL: 18, IQ4_XS: 136, MXFP4: 17}
WEIGHT_TYPES = [Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K]
EXT_TYPES = [Q4_1, Q5_1, Q2_K, Q3_K, IQ4_NL, IQ4_XS, MXFP4]
ACT_TYPE = {Q4_0: Q8_0, Q5_0: Q8_0, Q8_0: Q8_0, Q4_K: Q8_K, Q5_K: Q8_K, Q6_K: Q8_K,
            Q4_1: Q8_1, Q5_1: Q8_1, Q2_K: Q8_K, Q3_K: Q8_K, IQ4_NL: Q8_0, IQ4_XS: Q8_K, MXFP4: Q8_0}


def row_bytes(t, k):
    assert k % BLOCK_ELEMS[t] == 0
    return k // BLOCK_ELEMS[t] * BLOCK_BYTES[t]


# ----------------------------------------------------------------------------- bundles
def write_bundle(path, tensors):
    """tensors: list of (name, ggml_type, ne(list<=4), bytes-like)"""
    with open(path, "wb") as f:
        f.write(struct.pack("<II", 0x42543242, len(tensors)))
        for name, t, ne, data in tensors:
            ne = list(ne) + [1] * (4 - len(ne))
            b = np.ascontiguousarray(data).tobytes() if not isinstance(data, (bytes, bytearray)) else bytes(data)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb)
            f.write(struct.pack("<i4qQ", t, *ne, len(b)))
            f.write(b)


def read_bundle(path):
    out = {}
    with open(path, "rb") as f:
        magic, n = struct.unpack("<II", f.read(8))
        assert magic == 0x42543242
        for _ in range(n):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            t, ne0, ne1, ne2, ne3, nb = struct.unpack("<i4qQ", f.read(4 + 32 + 8))
            out[name] = (t, [ne0, ne1, ne2, ne3], f.read(nb))
    return out


# ----------------------------------------------------------------------------- oracle (our C port)
_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
        L.orc_fp16_to_fp32.restype = f32; L.orc_fp16_to_fp32.argtypes = [C.c_uint16]
        L.orc_fp32_to_fp16.restype = C.c_uint16; L.orc_fp32_to_fp16.argtypes = [f32]
        L.orc_quantize_row_q8_0.argtypes = [vp, vp, i64]
        L.orc_quantize_row_q8_K.argtypes = [vp, vp, i64]
        L.orc_dequantize_row.argtypes = [i32, vp, vp, i64]
        L.orc_vec_dot.restype = f32; L.orc_vec_dot.argtypes = [i32, i64, vp, vp]
        L.orc_mul_mat.argtypes = [i32, vp, vp, vp, i64, i64, i64]
        L.orc_rms_norm.argtypes = [vp, vp, vp, i64, i64, f32]
        L.orc_rope.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, f32, f32, f32, f32, f32, f32]
        L.orc_set_rows.argtypes = [vp, vp, vp, i32, i64, i64, i64]
        L.orc_flash_attn_ext.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, vp, i32, i64, i64, i64, i64, i64, i64, f32, f32, f32]
        L.orc_swiglu.argtypes = [vp, vp, vp, i64]
        L.orc_add.argtypes = [vp, vp, vp, i64, i64, i64]
        L.orc_mul.argtypes = [vp, vp, vp, i64, i64, i64]
        L.orc_get_rows_f32.argtypes = [vp, vp, vp, i64, i64]
        L.orc_cpy_f32_f16.argtypes = [vp, vp, i64]
        L.orc_quantize_row_q8_1.argtypes = [vp, vp, i64]
        L.orc_mul_mat_id.argtypes = [i32, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64]
        L.orc_get_rows_q.argtypes = [i32, vp, vp, vp, i64, i64]
        L.orc_soft_max_rows.argtypes = [vp, vp, i64, i64, f32]
        L.orc_argsort_rows.argtypes = [vp, vp, i64, i64, i32]
        L.orc_sum_rows.argtypes = [vp, vp, i64, i64]
        L.orc_unary.argtypes = [i32, vp, vp, i64, f32, f32]
        L.orc_mul_mat_f16.argtypes = [vp, i64, i64, i64, vp, i64, i64, vp, i64, i64, i64, i64, i64, i64]
        L.orc_soft_max_mask.argtypes = [vp, vp, vp, i32, i64, i64, i64, i64, f32, f32]
        _oracle = L
    return _oracle


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def orc_quantize_act(wtype, x):
    """x: f32 [n, k] -> uint8 [n, row_bytes(act)] in the weight type's vec_dot_type"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, k = x.shape
    at = ACT_TYPE[wtype]
    out = np.zeros((n, row_bytes(at, k)), dtype=np.uint8)
    fn = {Q8_0: oracle().orc_quantize_row_q8_0, Q8_1: oracle().orc_quantize_row_q8_1, Q8_K: oracle().orc_quantize_row_q8_K}[at]
    for i in range(n):
        fn(ptr(x[i]), ptr(out[i]), k)
    return out


def orc_mul_mat(wtype, W, X, m, n, k):
    X = np.ascontiguousarray(X, dtype=np.float32)
    dst = np.zeros((n, m), dtype=np.float32)
    oracle().orc_mul_mat(wtype, ptr(W), ptr(X), ptr(dst), m, n, k)
    return dst


def orc_dequant(t, W, nrows, k):
    out = np.zeros((nrows, k), dtype=np.float32)
    W = np.ascontiguousarray(W).reshape(nrows, -1)
    for i in range(nrows):
        oracle().orc_dequantize_row(t, ptr(W[i]), ptr(out[i]), k)
    return out


# ----------------------------------------------------------------------------- reference build (optional)
_ref = None


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libggml-cpu.so")) and os.path.exists(os.path.join(REF_DIR, "ref_ops"))


def ref():
    """ctypes handles to the unmodified reference: (libggml-base, libggml-cpu)"""
    global _ref
    if _ref is None:
        base = C.CDLL(os.path.join(REF_DIR, "libggml-base.so"), mode=C.RTLD_GLOBAL)
        cpu = C.CDLL(os.path.join(REF_DIR, "libggml-cpu.so"), mode=C.RTLD_GLOBAL)
        vp, i64 = C.c_void_p, C.c_int64
        base.ggml_quantize_chunk.restype = C.c_size_t
        base.ggml_quantize_chunk.argtypes = [C.c_int, vp, vp, i64, i64, i64, vp]
        for nm in ("q4_0", "q5_0", "q8_0", "q4_K", "q5_K", "q6_K", "q4_1", "q5_1", "q2_K", "q3_K", "iq4_nl", "iq4_xs", "mxfp4"):
            getattr(base, "dequantize_row_" + nm).argtypes = [vp, vp, i64]
        cpu.quantize_row_q8_0.argtypes = [vp, vp, i64]
        cpu.quantize_row_q8_1.argtypes = [vp, vp, i64]
        cpu.quantize_row_q8_K.argtypes = [vp, vp, i64]
        cpu.ggml_cpu_fp32_to_fp16.argtypes = [vp, vp, i64]
        for nm in ("q4_0_q8_0", "q5_0_q8_0", "q8_0_q8_0", "q4_K_q8_K", "q5_K_q8_K", "q6_K_q8_K",
                   "q4_1_q8_1", "q5_1_q8_1", "q2_K_q8_K", "q3_K_q8_K", "iq4_nl_q8_0", "iq4_xs_q8_K", "mxfp4_q8_0"):
            getattr(cpu, "ggml_vec_dot_" + nm).argtypes = [C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t, C.c_int]
        cpu.ggml_cpu_init()
        _ref = (base, cpu)
    return _ref


def ref_quantize_weights(t, w):
    """w f32 [m, k] -> uint8 [m, row_bytes] using the reference's ggml_quantize_chunk"""
    base, _ = ref()
    w = np.ascontiguousarray(w, dtype=np.float32)
    m, k = w.shape
    out = np.zeros((m, row_bytes(t, k)), dtype=np.uint8)
    if m * k < (1 << 22):
        base.ggml_quantize_chunk(t, ptr(w), ptr(out), 0, m, k, None)
        return out
    # big tensors (BASELINE-shaped vocab / ffn matrices): row chunks on a few threads (ctypes releases the GIL)
    from concurrent.futures import ThreadPoolExecutor
    nthr = max(1, min(16, len(os.sched_getaffinity(0))))
    step = (m + nthr - 1) // nthr

    def work(r0):
        r1 = min(m, r0 + step)
        base.ggml_quantize_chunk(t, ptr(w[r0:r1]), ptr(out[r0:r1]), 0, r1 - r0, k, None)
    with ThreadPoolExecutor(nthr) as tp:
        list(tp.map(work, range(0, m, step)))
    return out


def run_ref_op(op, tensors, params=None, backend="CPU", plugin=None, threads=1, tmpdir="/tmp"):
    """Run one ggml op through oracle/_ref/ref_ops; returns (type, ne, bytes) of dst."""
    import tempfile
    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        write_bundle(fin, tensors)
        cmd = [os.path.join(REF_DIR, "ref_ops"), "--op", op, "--in", fin, "--out", fout, "--backend", backend, "--threads", str(threads)]
        if plugin:
            cmd += ["--plugin", plugin]
        for k, v in (params or {}).items():
            cmd.append(f"{k}={v}")
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"ref_ops failed ({r.returncode}): {r.stderr[-2000:]}")
        return read_bundle(fout)["dst"]


# ----------------------------------------------------------------------------- data generators
def rand_f16_scale(rng, n, lo=1e-3, hi=2e-2, signed=False):
    v = rng.uniform(lo, hi, size=n).astype(np.float32)
    if signed:
        v *= rng.choice([-1.0, 1.0], size=n).astype(np.float32)
    return v.astype(np.float16).view(np.uint16)


def rand_blocks(rng, t, nrows, k, scale_mul=1.0):
    """Random but valid quant blocks (every bit pattern of qs/scales is legal; d finite)."""
    _rs = rand_f16_scale
    rand_f16_scale_l = lambda rng, n, lo=1e-3, hi=2e-2, signed=False: _rs(rng, n, lo * scale_mul, hi * scale_mul, signed)  # noqa: E731
    nb = nrows * (k // BLOCK_ELEMS[t])
    raw = rng.integers(0, 256, size=(nb, BLOCK_BYTES[t]), dtype=np.uint8)
    if t in (Q4_0, Q5_0, Q8_0):
        raw[:, 0:2] = rand_f16_scale_l(rng, nb, signed=True).view(np.uint8).reshape(nb, 2)
    elif t in (Q4_K, Q5_K):
        raw[:, 0:2] = rand_f16_scale_l(rng, nb, 1e-4, 2e-3).view(np.uint8).reshape(nb, 2)
        raw[:, 2:4] = rand_f16_scale_l(rng, nb, 1e-4, 2e-3).view(np.uint8).reshape(nb, 2)
    elif t == Q6_K:
        raw[:, 208:210] = rand_f16_scale_l(rng, nb, 1e-5, 2e-4, signed=True).view(np.uint8).reshape(nb, 2)
    elif t in (Q4_1, Q5_1):                                  # {d, m, ...}
        raw[:, 0:2] = rand_f16_scale_l(rng, nb).view(np.uint8).reshape(nb, 2)
        raw[:, 2:4] = rand_f16_scale_l(rng, nb, 1e-3, 1e-1, signed=True).view(np.uint8).reshape(nb, 2)
    elif t == IQ4_NL:
        raw[:, 0:2] = rand_f16_scale_l(rng, nb, 1e-4, 2e-3, signed=True).view(np.uint8).reshape(nb, 2)
    elif t == IQ4_XS:
        raw[:, 0:2] = rand_f16_scale_l(rng, nb, 1e-5, 2e-4, signed=True).view(np.uint8).reshape(nb, 2)
    elif t == MXFP4:                                         # E8M0 exponent byte: 2^-17 .. 2^-7 (halved)
        raw[:, 0] = rng.integers(112, 122, size=nb, dtype=np.uint8)
    elif t == Q2_K:                                          # {scales[16], qs[64], d, dmin}
        raw[:, 80:82] = rand_f16_scale_l(rng, nb, 1e-4, 2e-3).view(np.uint8).reshape(nb, 2)
        raw[:, 82:84] = rand_f16_scale_l(rng, nb, 1e-4, 2e-3).view(np.uint8).reshape(nb, 2)
    elif t == Q3_K:                                          # {hmask[32], qs[64], scales[12], d}
        raw[:, 108:110] = rand_f16_scale_l(rng, nb, 1e-4, 2e-3, signed=True).view(np.uint8).reshape(nb, 2)
    return raw.reshape(nrows, -1)


def repack_rows_np(t, W, k):
    """numpy model of b200_repack_rows (llama-box_b200/csrc/repack.cu): ggml rows -> the library's row layout."""
    W = np.ascontiguousarray(W).reshape(-1, row_bytes(t, k))
    nb = k // BLOCK_ELEMS[t]
    blk = W.reshape(W.shape[0], nb, BLOCK_BYTES[t])
    if t == Q4_0:
        parts = [blk[:, :, 2:18], blk[:, :, 0:2]]
    elif t == Q5_0:
        parts = [blk[:, :, 6:22], blk[:, :, 2:6], blk[:, :, 0:2]]
    elif t == Q8_0:
        parts = [blk[:, :, 2:34], blk[:, :, 0:2]]
    elif t == Q6_K:
        parts = [blk[:, :, 0:128], blk[:, :, 128:192], blk[:, :, 192:208], blk[:, :, 208:210]]
    else:
        return W.copy()
    return np.concatenate([p.reshape(W.shape[0], -1) for p in parts], axis=1)


def cos_data(n, off=0.0):
    """deterministic data of tests/test-quantize-fns.cpp:31-35: 0.1 + 2*cos(i + off)"""
    i = np.arange(n, dtype=np.float32)
    return (0.1 + 2.0 * np.cos(i + np.float32(off))).astype(np.float32)


def nmse(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    d = ((a - b) ** 2).sum(); s = (b ** 2).sum()
    return d / s if s > 0 else d
