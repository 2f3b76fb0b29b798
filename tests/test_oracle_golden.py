"""CPU: the C oracle (oracle/oracle.c) against the committed golden vectors produced by the
UNMODIFIED reference (tests/golden/, oracle/make_golden.py).  This is what pins the oracle on
machines where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest

from golden_cases import f32, load, raw2d
from refutil import (F16, Q8_0, Q8_K, TYPE_NAME, WEIGHT_TYPES, nmse, oracle, orc_dequant, orc_mul_mat, ptr, row_bytes)


def test_activation_quantisers_bit_exact():
    g = load("quantize_act.bin")
    x = f32(g["x"]); k = x.shape[1]
    for i in range(x.shape[0]):
        a = np.zeros(row_bytes(Q8_0, k), np.uint8); oracle().orc_quantize_row_q8_0(ptr(x[i]), ptr(a), k)
        assert np.array_equal(a, raw2d(g["q8_0"], 3)[i])
        b = np.zeros(row_bytes(Q8_K, k), np.uint8); oracle().orc_quantize_row_q8_K(ptr(x[i]), ptr(b), k)
        assert np.array_equal(b, raw2d(g["q8_K"], 3)[i])


def test_fp16_conversion_bit_exact():
    g = load("quantize_act.bin")
    xs = f32(g["f32"]); want = g["f16"][2].view(np.uint16)
    got = np.zeros(xs.size, np.uint16); oracle().orc_cpy_f32_f16(ptr(xs), ptr(got), xs.size)
    assert np.array_equal(got, want)
    back = np.array([oracle().orc_fp16_to_fp32(int(h)) for h in want[:512]], np.float32)
    assert np.array_equal(back, want[:512].view(np.float16).astype(np.float32))


@pytest.mark.parametrize("t", WEIGHT_TYPES)
def test_dequant_and_mul_mat(t):
    g = load(f"mul_mat_{TYPE_NAME[t]}.bin")
    _, ne, _ = g["w"]; k, m = ne[0], ne[1]
    W = raw2d(g["w"], m); X = f32(g["x"]); n = X.shape[0]
    assert np.array_equal(orc_dequant(t, W, m, k), f32(g["deq"]))          # bit-exact dequantisation
    want = f32(g["dst"]); got = orc_mul_mat(t, W, X, m, n, k)
    # integer block sums are identical; only the f32 summation order differs (SIMD lanes vs scalar)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    assert nmse(got, want) < 1e-12


def test_rms_norm_bit_exact():
    g = load("rms_norm.bin")
    x, w, want = f32(g["x"]), f32(g["w"]), f32(g["dst"])
    y = np.zeros_like(x); oracle().orc_rms_norm(ptr(x), ptr(w), ptr(y), x.shape[1], x.shape[0], 1e-5)
    assert np.array_equal(y, want)


@pytest.mark.parametrize("tag", ["norm_ff", "neox", "yarn"])
def test_rope(tag):
    g = load(f"rope_{tag}.bin")
    x = f32(g["x"]); pos = g["pos"][2].view(np.int32); pv = f32(g["params"])
    ff = f32(g["ff"]) if "ff" in g else None
    nt, nh, hd = x.shape
    y = np.zeros_like(x)
    oracle().orc_rope(ptr(x), ptr(y), ptr(pos), ptr(ff), hd, nh, nt, int(pv[0]), int(pv[1]), int(pv[2]), pv[3], pv[4], pv[5], pv[6], pv[7], pv[8])
    # NORM/NEOX: theta is the same f32 product sequence -> ~1e-7.  YaRN blends two thetas of magnitude
    # ~pos; the reference build lets gcc contract that blend into an FMA (default -ffp-contract=fast), so a
    # 1-ulp theta difference at pos=20000 (ulp 2e-3 rad) is inherent to the reference, not to the port.
    assert np.abs(y - f32(g["dst"])).max() <= (1e-4 if tag == "yarn" else 1e-6)


@pytest.mark.parametrize("dt", [F16, Q8_0])
def test_set_rows_bit_exact(dt):
    g = load(f"set_rows_{TYPE_NAME[dt]}.bin")
    src = f32(g["src"]); ids = g["ids"][2].view(np.int64); _, ne, want = g["dst"]
    nc, tot = ne[0], ne[1]; stride = row_bytes(dt, nc)
    cache = np.zeros((tot, stride), np.uint8)
    oracle().orc_set_rows(ptr(src), ptr(ids), ptr(cache), dt, nc, src.shape[0], stride)
    assert np.array_equal(cache.reshape(-1), want)


@pytest.mark.parametrize("kvt", [F16, Q8_0])
def test_flash_attn(kvt):
    g = load(f"flash_attn_{TYPE_NAME[kvt]}.bin")
    q = f32(g["q"]); nt, nh, dk = q.shape
    _, kne, kc = g["k"]; nkv = kne[1]; nhkv = kne[0] // dk
    vc = g["v"][2]; mask = g["mask"][2].view(np.uint16)
    rb_row, rb_head = row_bytes(kvt, nhkv * dk), row_bytes(kvt, dk)
    y = np.zeros((nt, nh, dk), np.float32)
    oracle().orc_flash_attn_ext(ptr(q), nh * dk * 4, dk * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(mask), ptr(y),
                                kvt, dk, dk, nh, nhkv, nt, nkv, float(1 / np.sqrt(dk)), 0.0, 0.0)
    want = f32(g["dst"])
    tol = 2e-3 if kvt == F16 else 2e-6      # F16 V: fp16 accumulator, SIMD association differs
    assert np.abs(y - want).max() <= tol * np.abs(want).max()


def test_swiglu():
    g = load("swiglu.bin")
    a, b, want = f32(g["gate"]), f32(g["up"]), f32(g["dst"])
    y = np.zeros_like(a); oracle().orc_swiglu(ptr(a), ptr(b), ptr(y), a.size)
    assert np.abs(y - want).max() <= 1e-6 * np.abs(want).max()   # the reference uses a vectorised expf (ggml_v_expf)


# ---- the wide path: oracle_ext.c against fixtures the unmodified reference produced (oracle/make_golden.py) ----------
from refutil import EXT_TYPES, Q4_K, Q8_1  # noqa: E402


@pytest.mark.parametrize("t", EXT_TYPES)
def test_ext_dequant_and_mul_mat(t):
    g = load(f"ext_mul_mat_{TYPE_NAME[t]}.bin")
    _, ne, _ = g["w"]; k, m = ne[0], ne[1]
    W = raw2d(g["w"], m); X = f32(g["x"]); n = X.shape[0]
    assert np.array_equal(orc_dequant(t, W, m, k), f32(g["deq"]))
    want = f32(g["dst"]); got = orc_mul_mat(t, W, X, m, n, k)
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


def test_q8_1_mul_mat_id_get_rows():
    g = load("ext_moe_get_rows.bin")
    x = f32(g["x81"]); k = x.shape[1]
    for i in range(x.shape[0]):
        a = np.zeros(row_bytes(Q8_1, k), np.uint8); oracle().orc_quantize_row_q8_1(ptr(x[i]), ptr(a), k)
        assert np.array_equal(a, raw2d(g["q8_1"], 3)[i])
    _, ne, W = g["w"]; ke, me, n_expert = ne[0], ne[1], ne[2]
    ids = g["ids"][2].view(np.int32); n_used, n_tok = g["ids"][1][0], g["ids"][1][1]
    for tag in ("shared", "per_expert"):
        b = g["b_" + tag][2].view(np.float32); n_b1 = g["b_" + tag][1][1]
        want = g["dst_" + tag][2].view(np.float32)
        got = np.zeros_like(want)
        oracle().orc_mul_mat_id(Q4_K, ptr(W), ptr(b), ptr(ids), ptr(got), me, ke, n_expert, n_used, n_tok, n_b1, n_used)
        assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()
    gid = g["gr_ids"][2].view(np.int32); want = g["gr_dst"][2].view(np.float32)
    got = np.zeros_like(want)
    oracle().orc_get_rows_q(Q4_K, ptr(W), ptr(gid), ptr(got), ke, gid.size)
    assert np.array_equal(got, want)
