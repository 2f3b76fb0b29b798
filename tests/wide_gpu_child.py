"""Child process of tests/test_gpu_zz_wide.py: runs ONE wide-path check on cuda:0 and prints a JSON line.  A separate process so that a
faulting kernel cannot poison the CUDA context of the pytest process that runs the rest of the GPU suite."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_pkg  # noqa: E402
from refutil import (BLOCK_ELEMS, F32, I32, Q4_K, Q6_K, have_ref, oracle, orc_dequant, orc_mul_mat, ptr, rand_blocks, ref_quantize_weights,  # noqa: E402
                     repack_rows_np, row_bytes)

import torch  # noqa: E402

pkg = load_pkg()
ops = pkg.ops
L = ops.lib


def dev(a):
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.view(np.uint8).reshape(-1) if a.dtype not in (np.float32, np.int32) else a).cuda()
    return t


def weights(rng, t, rows, k, kind):
    if kind == "quantised" and have_ref():
        return ref_quantize_weights(t, (rng.standard_normal((rows, k)) * 0.05).astype(np.float32))
    return rand_blocks(rng, t, rows, k)


def rel(got, want):
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def do_mul_mat(t, m, k, ncols, kind):
    rng = np.random.default_rng(1000 * t + k + ncols)
    W = weights(rng, t, m, k, kind)
    X = rng.standard_normal((ncols, k)).astype(np.float32)
    bias = rng.standard_normal(m).astype(np.float32); res = rng.standard_normal((ncols, m)).astype(np.float32)
    want = orc_mul_mat(t, W, X, m, ncols, k)
    Wd = dev(np.concatenate([repack_rows_np(t, W, k).reshape(-1), np.zeros(64, np.uint8)])); Xd = dev(X)
    out = {}
    dst = torch.zeros((ncols, m), dtype=torch.float32, device="cuda")
    ops.check(L.b200_mul_mat_vec_wide(t, ops.p(Wd), ops.p(Xd), k, ops.p(dst), m, None, None, m, k, ncols, ops.stream()))
    torch.cuda.synchronize()
    out["plain"] = rel(dst.cpu().numpy(), want)
    bd, rd = dev(bias), dev(res)
    ops.check(L.b200_mul_mat_vec_wide(t, ops.p(Wd), ops.p(Xd), k, ops.p(dst), m, ops.p(bd), ops.p(rd), m, k, ncols, ops.stream()))
    torch.cuda.synchronize()
    out["bias_residual"] = rel(dst.cpu().numpy(), want + bias[None, :] + res)
    return out


def do_mul_mat_id(t, m, k, n_expert, n_used, n_tok, shared):
    rng = np.random.default_rng(7 * t + k + n_tok)
    n_b1 = 1 if shared else n_used
    W = weights(rng, t, n_expert * m, k, "quantised")
    b = rng.standard_normal((n_tok, n_b1, k)).astype(np.float32)
    ids = np.stack([rng.permutation(n_expert)[:n_used] for _ in range(n_tok)]).astype(np.int32)
    want = np.zeros((n_tok, n_used, m), np.float32)
    oracle().orc_mul_mat_id(t, ptr(W), ptr(b), ptr(ids), ptr(want), m, k, n_expert, n_used, n_tok, n_b1, n_used)
    Wd = dev(np.concatenate([repack_rows_np(t, W, k).reshape(-1), np.zeros(64, np.uint8)])); bd = dev(b); idd = dev(ids)
    dst = torch.full((n_tok, n_used, m), 7.0, dtype=torch.float32, device="cuda")
    ops.check(L.b200_mul_mat_id(t, ops.p(Wd), m * row_bytes(t, k), ops.p(bd), n_b1 * k, k, n_b1, ops.p(idd), n_used, ops.p(dst), n_used * m, m,
                                m, k, n_expert, n_used, n_tok, ops.stream()))
    torch.cuda.synchronize()
    return {"err": rel(dst.cpu().numpy(), want)}


def do_get_rows(t, nrows, k):
    rng = np.random.default_rng(13 * t + k)
    W = weights(rng, t, nrows, k, "quantised")
    ids = np.array([3, nrows - 1, 0, 3, 17], np.int32)
    want = np.zeros((ids.size, k), np.float32)
    oracle().orc_get_rows_q(t, ptr(W), ptr(ids), ptr(want), k, ids.size)
    Wd = dev(np.concatenate([repack_rows_np(t, W, k).reshape(-1), np.zeros(64, np.uint8)])); idd = dev(ids)
    dst = torch.zeros((ids.size, k), dtype=torch.float32, device="cuda")
    ops.check(L.b200_get_rows_q(t, ops.p(Wd), row_bytes(t, k), nrows, ops.p(idd), ops.p(dst), k, k, ids.size, ops.stream()))
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    return {"bit_exact": bool(np.array_equal(got, want)), "err": rel(got, want)}


def do_repack_model(t, k):
    """the numpy model of the library layout (refutil.repack_rows_np) against the real b200_repack_rows"""
    rng = np.random.default_rng(t)
    W = rand_blocks(rng, t, 16, k)
    Wd = dev(np.concatenate([W.reshape(-1), np.zeros(64, np.uint8)]))
    ops.check(L.b200_repack_rows(t, ops.p(Wd), 16, k, ops.stream()))
    torch.cuda.synchronize()
    return {"equal": bool(np.array_equal(Wd.cpu().numpy()[:W.size], repack_rows_np(t, W, k).reshape(-1)))}


def do_executor(wtype, n_tok):
    """token embedding (GET_ROWS on a Q4_K table) -> RMS_NORM * w -> MUL_MAT on a wide-only format -> MoE FFN (3 x MUL_MAT_ID + SwiGLU) through the graph
    executor, with CUDA graphs + fusion, against the oracle run op by op"""
    assert os.environ.get("GGML_B200_WIDE") == "1"
    G = importlib.import_module("llama_box_b200.graph")
    rng = np.random.default_rng(wtype + n_tok)
    E, FF, NE, NU, VOC = 1024, 2048, 6, 2, 512
    table = ref_quantize_weights(Q4_K, (rng.standard_normal((VOC, E)) * 0.05).astype(np.float32)) if have_ref() else rand_blocks(rng, Q4_K, VOC, E)
    Wp = weights(rng, wtype, E, E, "quantised")
    Wup = weights(rng, Q4_K, NE * FF, E, "quantised"); Wgate = weights(rng, Q4_K, NE * FF, E, "quantised"); Wdown = weights(rng, Q6_K, NE * E, FF, "quantised")
    nw = (1 + 0.1 * rng.standard_normal(E)).astype(np.float32)
    tok = rng.integers(0, VOC, n_tok).astype(np.int32)
    ids = np.stack([rng.permutation(NE)[:NU] for _ in range(n_tok)]).astype(np.int32)
    # ---- oracle
    O = oracle()
    emb = np.zeros((n_tok, E), np.float32); O.orc_get_rows_q(Q4_K, ptr(table), ptr(tok), ptr(emb), E, n_tok)
    nrm = np.zeros_like(emb); O.orc_rms_norm(ptr(emb), ptr(nw), ptr(nrm), E, n_tok, 1e-5)
    proj = orc_mul_mat(wtype, Wp, nrm, E, n_tok, E)
    up = np.zeros((n_tok, NU, FF), np.float32); gate = np.zeros_like(up)
    O.orc_mul_mat_id(Q4_K, ptr(Wup), ptr(proj), ptr(ids), ptr(up), FF, E, NE, NU, n_tok, 1, NU)
    O.orc_mul_mat_id(Q4_K, ptr(Wgate), ptr(proj), ptr(ids), ptr(gate), FF, E, NE, NU, n_tok, 1, NU)
    act = np.zeros_like(up); O.orc_swiglu(ptr(gate), ptr(up), ptr(act), act.size)
    want = np.zeros((n_tok, NU, E), np.float32)
    O.orc_mul_mat_id(Q6_K, ptr(Wdown), ptr(act), ptr(ids), ptr(want), E, FF, NE, NU, n_tok, NU, NU)
    # ---- executor
    keep = []

    def up_w(t, W, k):
        d = dev(np.concatenate([repack_rows_np(t, W, k).reshape(-1), np.zeros(64, np.uint8)])); keep.append(d); return d

    def f32t(ne):
        d = torch.zeros(int(np.prod(ne)), dtype=torch.float32, device="cuda"); keep.append(d); return G.T(d.data_ptr(), G.F32, ne), d
    nl = G.NodeList()
    tokd = dev(tok); idd = dev(ids); nwd = dev(nw); keep += [tokd, idd, nwd]
    tab = G.T(up_w(Q4_K, table, E).data_ptr(), G.Q4_K, [E, VOC])
    e_t, _ = f32t([E, n_tok]); nl.add(G.OP_GET_ROWS, e_t, [tab, G.T(tokd.data_ptr(), G.I32, [n_tok])])
    n_t, _ = f32t([E, n_tok]); nl.add(G.OP_RMS_NORM, n_t, [e_t], [G.f32_bits(1e-5)])
    c_t = G.T(n_t.ptr, G.F32, [E, n_tok]); nl.add(G.OP_MUL, c_t, [n_t, G.T(nwd.data_ptr(), G.F32, [E])])
    p_t, _ = f32t([E, n_tok]); nl.add(G.OP_MUL_MAT, p_t, [G.T(up_w(wtype, Wp, E).data_ptr(), wtype, [E, E]), c_t])
    x3 = nl.view_op(G.T(p_t.ptr, G.F32, [E, 1, n_tok], [4, 4 * E, 4 * E, 4 * E * n_tok]), p_t)
    idt = G.T(idd.data_ptr(), G.I32, [NU, n_tok])
    u_t, _ = f32t([FF, NU, n_tok]); nl.add(G.OP_MUL_MAT_ID, u_t, [G.T(up_w(Q4_K, Wup, E).data_ptr(), G.Q4_K, [E, FF, NE]), x3, idt])
    g_t, _ = f32t([FF, NU, n_tok]); nl.add(G.OP_MUL_MAT_ID, g_t, [G.T(up_w(Q4_K, Wgate, E).data_ptr(), G.Q4_K, [E, FF, NE]), x3, idt])
    a_t, _ = f32t([FF, NU, n_tok]); nl.add(G.OP_GLU_SWIGLU, a_t, [g_t, u_t], [2, 0])
    d_t, dd = f32t([E, NU, n_tok]); nl.add(G.OP_MUL_MAT_ID, d_t, [G.T(up_w(Q6_K, Wdown, FF).data_ptr(), G.Q6_K, [FF, E, NE]), a_t, idt])
    nodes = nl.build()
    ex = G.Executor(0)
    sup = [bool(ex.supports(nodes[i])) for i in range(len(nodes))]
    errs = []
    for rep in range(3):                                   # eager, capture, replay
        dd.zero_()
        ex.compute(nodes)
        torch.cuda.synchronize()
        errs.append(rel(dd.cpu().numpy().reshape(n_tok, NU, E), want))
    return {"supports": sup, "errs": errs, "captures": int(ex.captures), "replays": int(ex.replays)}


def do_kv_q4_0(d, n_head, n_head_kv, n_tok, n_kv):
    """KV cache type q4_0: SET_ROWS bytes vs the oracle's from_float, then FLASH_ATTN_EXT over the cache vs the oracle (causal-style mask, GQA)"""
    from refutil import Q4_0
    rng = np.random.default_rng(d + n_kv + n_tok)
    rb_row, rb_head = row_bytes(Q4_0, n_head_kv * d), row_bytes(Q4_0, d)
    kf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32); vf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32)
    kf[1, :32] = 0
    ids = rng.permutation(n_kv).astype(np.int64)
    kc = np.zeros((n_kv, rb_row), np.uint8); vc = kc.copy()
    oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), Q4_0, n_head_kv * d, n_kv, rb_row)
    oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), Q4_0, n_head_kv * d, n_kv, rb_row)
    kfd, vfd, idd = dev(kf), dev(vf), torch.from_numpy(ids).cuda()
    kcd = torch.zeros(n_kv * rb_row + 64, dtype=torch.uint8, device="cuda"); vcd = torch.zeros_like(kcd)
    for lo in range(0, n_kv, 4096):                         # a few calls like a sequence of ubatches
        n = min(4096, n_kv - lo)
        ops.check(L.b200_set_rows_q4_0(C.c_void_p(kfd.data_ptr() + lo * n_head_kv * d * 4), n_head_kv * d, C.c_void_p(idd.data_ptr() + lo * 8), ops.p(kcd), rb_row, n_head_kv * d, n, ops.stream()))
        ops.check(L.b200_set_rows_q4_0(C.c_void_p(vfd.data_ptr() + lo * n_head_kv * d * 4), n_head_kv * d, C.c_void_p(idd.data_ptr() + lo * 8), ops.p(vcd), rb_row, n_head_kv * d, n, ops.stream()))
    torch.cuda.synchronize()
    out = {"set_rows_equal": bool(np.array_equal(kcd.cpu().numpy()[:kc.size], kc.reshape(-1)) and np.array_equal(vcd.cpu().numpy()[:vc.size], vc.reshape(-1)))}
    q = rng.standard_normal((n_tok, n_head, d)).astype(np.float32)
    npad = (n_tok + 63) // 64 * 64
    mask = np.full((npad, n_kv), -np.inf, np.float32)
    for t in range(n_tok):
        mask[t, :max(1, n_kv - n_tok + t + 1 - 7)] = 0
    m16 = mask.astype(np.float16)
    scale = float(1 / np.sqrt(d))
    want = np.zeros((n_tok, n_head, d), np.float32)
    oracle().orc_flash_attn_ext(ptr(q), n_head * d * 4, d * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), ptr(want),
                                Q4_0, d, d, n_head, n_head_kv, n_tok, n_kv, scale, 0.0, 0.0)
    qd = dev(q); md = torch.from_numpy(m16.view(np.int16)).cuda()
    dst = torch.zeros((n_tok, n_head, d), dtype=torch.float32, device="cuda")
    kc2 = dev(np.concatenate([kc.reshape(-1), np.zeros(64, np.uint8)])); vc2 = dev(np.concatenate([vc.reshape(-1), np.zeros(64, np.uint8)]))
    ops.check(L.b200_flash_attn_q4_0(ops.p(qd), n_head * d, d, ops.p(kc2), rb_row, rb_head, ops.p(vc2), rb_row, rb_head, ops.p(md), n_kv, ops.p(dst),
                                     d, n_head, n_head_kv, n_tok, n_kv, scale, 0.0, 0.0, ops.stream()))
    torch.cuda.synchronize()
    out["attn_err"] = rel(dst.cpu().numpy(), want)
    return out


def do_type_suite(t):
    """every C-ABI case of one format in ONE process (a fresh interpreter + torch import per case would dominate the run time)"""
    out = {}

    def run(name, fn, *a):
        try:
            out[name] = fn(*a)
        except Exception as e:                              # a CUDA fault makes every later case of this child fail too: they are reported as such
            out[name] = {"error": repr(e)[:300]}
    for (m, k, n) in [(64, 256, 1), (512, 4096, 1), (96, 2048, 3), (200, 14336, 8)]:
        if t == Q6_K and k % 512:
            k = 512
        for kind in ("random", "quantised"):
            run(f"mul_mat {m}x{k}x{n} {kind}", do_mul_mat, t, m, k, n, kind)
    if BLOCK_ELEMS[t] == 32 and t not in (2, 6, 8):
        run("mul_mat k=29568", do_mul_mat, t, 48, 29568, 2, "random")          # Qwen2-72B's n_ff: not a multiple of 256
    for shared in (1, 0):
        for n_tok in (1, 5):
            run(f"mul_mat_id shared={shared} n_tok={n_tok}", do_mul_mat_id, t, 96, 2048, 6, 3, n_tok, shared)
    run("get_rows", do_get_rows, t, 40, 2048)
    return out


def main():
    what = sys.argv[1]; a = [int(v) if v.lstrip("-").isdigit() else v for v in sys.argv[2:]]
    torch.cuda.set_device(0)
    fn = {"kv_q4_0": do_kv_q4_0, "type_suite": do_type_suite, "mul_mat": do_mul_mat, "mul_mat_id": do_mul_mat_id, "get_rows": do_get_rows, "repack_model": do_repack_model, "executor": do_executor}[what]
    print("RESULT " + json.dumps(fn(*a)))


if __name__ == "__main__":
    main()
