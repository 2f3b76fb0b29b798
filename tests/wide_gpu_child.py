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
    """one Mixtral-style block through the graph executor (CUDA graphs + fusion) against the oracle run op by op: token embedding (GET_ROWS on a Q4_K
    table) -> RMS_NORM * w -> MUL_MAT on a wide-only format -> build_moe_ffn: f32 router matmul, SOFT_MAX, ARGSORT / top-k view, batched GET_ROWS,
    SUM_ROWS, DIV, 3 x MUL_MAT_ID + SwiGLU, weighted MUL, ADD over the expert slices (tests/moe_graph.py)"""
    assert os.environ.get("GGML_B200_WIDE") == "1"
    import moe_graph
    G = importlib.import_module("llama_box_b200.graph")
    E, FF, NE, NU, VOC = 1024, 2048, 6, 2, 512
    O = oracle()
    for attempt in range(8):                                # a seed whose top-k is not a near-tie (the f32 router sums associate differently)
        rng = np.random.default_rng(1000 * attempt + 10 * wtype + n_tok)
        table = ref_quantize_weights(Q4_K, (rng.standard_normal((VOC, E)) * 0.05).astype(np.float32)) if have_ref() else rand_blocks(rng, Q4_K, VOC, E)
        Wp = weights(rng, wtype, E, E, "quantised")
        Wg = rng.standard_normal((NE, E)).astype(np.float32)
        nw = (1 + 0.1 * rng.standard_normal(E)).astype(np.float32)
        tok = rng.integers(0, VOC, n_tok).astype(np.int32)
        emb = np.zeros((n_tok, E), np.float32); O.orc_get_rows_q(Q4_K, ptr(table), ptr(tok), ptr(emb), E, n_tok)
        nrm = np.zeros_like(emb); O.orc_rms_norm(ptr(emb), ptr(nw), ptr(nrm), E, n_tok, 1e-5)
        proj = orc_mul_mat(wtype, Wp, nrm, E, n_tok, E)
        logits = (proj.astype(np.float64) @ Wg.astype(np.float64).T).astype(np.float32)
        probs = np.zeros_like(logits); O.orc_soft_max_rows(ptr(logits), ptr(probs), NE, n_tok, 1.0)
        srt = np.sort(probs, axis=1)[:, ::-1]
        if np.all(srt[:, NU - 1] - srt[:, NU] > 1e-3 * srt[:, NU - 1]):
            break
    Wup = weights(rng, Q4_K, NE * FF, E, "quantised"); Wgate = weights(rng, Q4_K, NE * FF, E, "quantised"); Wdown = weights(rng, Q6_K, NE * E, FF, "quantised")
    order = np.zeros((n_tok, NE), np.int32); O.orc_argsort_rows(ptr(probs), ptr(order), NE, n_tok, 1)
    ids = np.ascontiguousarray(order[:, :NU])
    w = np.take_along_axis(probs, ids, axis=1)
    wsum = np.zeros(n_tok, np.float32); O.orc_sum_rows(ptr(np.ascontiguousarray(w)), ptr(wsum), NU, n_tok)
    wn = (w / wsum[:, None]).astype(np.float32)
    up = np.zeros((n_tok, NU, FF), np.float32); gate = np.zeros_like(up)
    O.orc_mul_mat_id(Q4_K, ptr(Wup), ptr(proj), ptr(ids), ptr(up), FF, E, NE, NU, n_tok, 1, NU)
    O.orc_mul_mat_id(Q4_K, ptr(Wgate), ptr(proj), ptr(ids), ptr(gate), FF, E, NE, NU, n_tok, 1, NU)
    act = np.zeros_like(up); O.orc_swiglu(ptr(gate), ptr(up), ptr(act), act.size)
    down = np.zeros((n_tok, NU, E), np.float32)
    O.orc_mul_mat_id(Q6_K, ptr(Wdown), ptr(act), ptr(ids), ptr(down), E, FF, NE, NU, n_tok, NU, NU)
    exw = down * wn[:, :, None]
    want = exw[:, 0, :].copy()
    for i in range(1, NU):
        want = want + exw[:, i, :]
    # ---- executor
    keep = []
    inputs = {"table": (Q4_K, table, E), "w_proj": (wtype, Wp, E), "up_exps": (Q4_K, Wup, E), "gate_exps": (Q4_K, Wgate, E), "down_exps": (Q6_K, Wdown, FF)}
    plain = {"tok": tok, "norm_w": nw, "gate_inp": Wg}

    class Alloc:
        def f32(self, ne):
            d = torch.zeros(int(np.prod(ne)), dtype=torch.float32, device="cuda"); keep.append(d); return G.T(d.data_ptr(), G.F32, ne)

        def i32(self, ne):
            d = torch.zeros(int(np.prod(ne)), dtype=torch.int32, device="cuda"); keep.append(d); return G.T(d.data_ptr(), G.I32, ne)

        def named(self, name, t, ne):
            if name in plain:
                d = dev(plain[name])
            else:
                tt, W, k = inputs[name]
                d = dev(np.concatenate([repack_rows_np(tt, W, k).reshape(-1), np.zeros(64, np.uint8)]))
            keep.append(d)
            return G.T(d.data_ptr(), t, ne)
    nl, out_t = moe_graph.build(G, Alloc(), E, FF, NE, NU, n_tok, VOC, wtype)
    nodes = nl.build()
    ex = G.Executor(0)
    sup = [bool(ex.supports(nodes[i])) for i in range(len(nodes))]
    out_d = [d for d in keep if d.data_ptr() == out_t.ptr][0]
    errs = []
    for rep in range(3):                                   # eager, capture, replay
        out_d.zero_()
        ex.compute(nodes)
        torch.cuda.synchronize()
        errs.append(rel(out_d.cpu().numpy().reshape(n_tok, E), want))
    return {"supports": sup, "errs": errs, "captures": int(ex.captures), "replays": int(ex.replays), "kernels": int(ex.last_kernels)}


def do_nofa(n_tok, n_past):
    """one attention block WITHOUT -fa through the graph executor (tests/nofa_graph.py = what build_attn_mha emits) against the oracle op by op"""
    assert os.environ.get("GGML_B200_WIDE") == "1"
    import nofa_graph
    from refutil import F16 as T_F16
    G = importlib.import_module("llama_box_b200.graph")
    O = oracle()
    rng = np.random.default_rng(n_tok + n_past)
    E, H, HK, D, KV_SIZE = 1024, 8, 2, 128, 256
    NKV = (n_past + n_tok + 31) // 32 * 32
    x = rng.standard_normal((n_tok, E)).astype(np.float32)
    Wq = weights(rng, Q4_K, H * D, E, "quantised"); Wk = weights(rng, Q4_K, HK * D, E, "quantised"); Wv = weights(rng, Q4_K, HK * D, E, "quantised"); Wo = weights(rng, Q4_K, E, H * D, "quantised")
    kc = np.zeros((KV_SIZE, HK * D), np.float16); vt = np.zeros((HK * D, KV_SIZE), np.float16)
    kc[:n_past] = rng.standard_normal((n_past, HK * D)).astype(np.float16); vt[:, :n_past] = rng.standard_normal((HK * D, n_past)).astype(np.float16)
    pos = np.arange(n_past, n_past + n_tok, dtype=np.int32); cells = np.arange(n_past, n_past + n_tok, dtype=np.int64)
    vidx = (np.arange(HK * D, dtype=np.int64)[None, :] * KV_SIZE + cells[:, None]).reshape(-1)
    mask = np.full((64, NKV), -np.inf, np.float32)
    for t in range(n_tok):
        mask[t, :n_past + t + 1] = 0
    rp = dict(n_dims=D, mode=0, n_ctx_orig=8192, freq_base=5e5, freq_scale=1.0, ext_factor=0.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0)
    # ---- oracle
    q = orc_mul_mat(Q4_K, Wq, x, H * D, n_tok, E); k = orc_mul_mat(Q4_K, Wk, x, HK * D, n_tok, E); v = orc_mul_mat(Q4_K, Wv, x, HK * D, n_tok, E)
    qr = np.zeros_like(q); kr = np.zeros_like(k)
    for src, dst, nh in ((q, qr, H), (k, kr, HK)):
        O.orc_rope(ptr(src), ptr(dst), ptr(pos), None, D, nh, n_tok, D, 0, 8192, 5e5, 1.0, 0.0, 1.0, 32.0, 1.0)
    kc_w = kc.copy(); vt_w = vt.copy()
    kc_w[cells] = kr.astype(np.float16); vt_w.reshape(-1)[vidx] = v.reshape(-1).astype(np.float16)
    kq = np.zeros((H, n_tok, NKV), np.float32)
    O.orc_mul_mat_f16(ptr(kc_w), HK * D * 2, D * 2, HK, ptr(qr), H * D * 4, D * 4, ptr(kq), NKV * 4, n_tok * NKV * 4, NKV, n_tok, H, D)
    sm = np.zeros_like(kq); O.orc_soft_max_mask(ptr(kq), ptr(sm), ptr(mask), 0, NKV, NKV, n_tok, H, float(D ** -0.5), 0.0)
    kqv = np.zeros((H, n_tok, D), np.float32)
    O.orc_mul_mat_f16(ptr(vt_w), KV_SIZE * 2, KV_SIZE * D * 2, HK, ptr(sm), NKV * 4, n_tok * NKV * 4, ptr(kqv), D * 4, n_tok * D * 4, D, n_tok, H, NKV)
    cont = np.ascontiguousarray(kqv.transpose(1, 0, 2)).reshape(n_tok, H * D)
    want = orc_mul_mat(Q4_K, Wo, cont, E, n_tok, H * D)
    # ---- executor
    keep = []
    plain = {"x": x, "pos": pos, "k_idx": cells, "v_idx": vidx, "mask": mask}
    wts = {"wq": (Wq, E), "wk": (Wk, E), "wv": (Wv, E), "wo": (Wo, H * D)}
    caches = {"k_cache": kc, "v_cache": vt}
    handles = {}

    class Alloc:
        def f32(self, ne):
            d = torch.zeros(int(np.prod(ne)), dtype=torch.float32, device="cuda"); keep.append(d); return G.T(d.data_ptr(), G.F32, ne)

        def named(self, name, t, ne):
            if name in plain:
                a = plain[name]
                d = torch.from_numpy(np.ascontiguousarray(a)).cuda() if a.dtype == np.int64 else dev(a)
            elif name in caches:
                d = torch.from_numpy(caches[name].view(np.int16).copy()).cuda()
            else:
                W, k_ = wts[name]
                d = dev(np.concatenate([repack_rows_np(Q4_K, W, k_).reshape(-1), np.zeros(64, np.uint8)]))
            keep.append(d); handles[name] = d
            return G.T(d.data_ptr(), t, ne)
    rope_params = [0, D, 0, 0, 8192, G.f32_bits(5e5), G.f32_bits(1.0), G.f32_bits(0.0), G.f32_bits(1.0), G.f32_bits(32.0), G.f32_bits(1.0)]
    nl, out_t = nofa_graph.build(G, Alloc(), E, H, HK, D, n_tok, KV_SIZE, NKV, rope_params)
    nodes = nl.build()
    ex = G.Executor(0)
    sup = [bool(ex.supports(nodes[i])) for i in range(len(nodes))]
    out_d = [d for d in keep if d.data_ptr() == out_t.ptr][0]
    errs = []
    for rep in range(3):
        out_d.zero_()
        ex.compute(nodes)
        torch.cuda.synchronize()
        errs.append(rel(out_d.cpu().numpy().reshape(n_tok, E), want))
    k_ok = bool(np.array_equal(handles["k_cache"].cpu().numpy().view(np.float16).reshape(KV_SIZE, HK * D)[cells].view(np.uint16), kr.astype(np.float16).view(np.uint16)))
    v_ok = bool(np.array_equal(handles["v_cache"].cpu().numpy().view(np.uint16).reshape(-1)[vidx], v.reshape(-1).astype(np.float16).view(np.uint16)))
    return {"supports": sup, "errs": errs, "k_store_exact": k_ok, "v_store_exact": v_ok, "captures": int(ex.captures), "replays": int(ex.replays), "kernels": int(ex.last_kernels)}


def do_glue():
    """the router glue of build_moe_ffn op by op through the C-ABI (glue_ext.cu) against the oracle / exact numpy semantics"""
    rng = np.random.default_rng(9)
    E, NE, NU, NT = 1024, 8, 2, 5
    O = oracle()
    i64a = lambda *v: dev(np.array(v, np.int64))  # noqa: E731  (the stride arrays are read by the host wrapper: keep them on the host)
    h64 = lambda *v: np.array(v, np.int64)  # noqa: E731
    out = {}
    Wg = rng.standard_normal((NE, E)).astype(np.float32); x = rng.standard_normal((NT, E)).astype(np.float32)
    Wd, xd = dev(Wg), dev(x); lg = torch.zeros((NT, NE), dtype=torch.float32, device="cuda")
    ops.check(L.b200_mul_mat_f32(ops.p(Wd), E, ops.p(xd), E, ops.p(lg), NE, NE, E, NT, ops.stream())); torch.cuda.synchronize()
    out["mul_mat_f32"] = rel(lg.cpu().numpy(), x.astype(np.float64) @ Wg.astype(np.float64).T)
    logits = lg.cpu().numpy().copy(); logits[2, 3] = logits[2, 6]
    ld = dev(logits); pd = torch.zeros_like(ld)
    ops.check(L.b200_soft_max_rows(ops.p(ld), NE, ops.p(pd), NE, NE, NT, 1.0, ops.stream())); torch.cuda.synchronize()
    wantp = np.zeros_like(logits); O.orc_soft_max_rows(ptr(logits), ptr(wantp), NE, NT, 1.0)
    out["soft_max"] = float(np.abs(pd.cpu().numpy() - wantp).max())
    wpd = dev(wantp); idd = torch.zeros((NT, NE), dtype=torch.int32, device="cuda")
    ok = True
    for desc in (0, 1):
        ops.check(L.b200_argsort_rows(ops.p(wpd), NE, ops.p(idd), NE, NE, NT, desc, ops.stream())); torch.cuda.synchronize()
        wi = np.zeros((NT, NE), np.int32); O.orc_argsort_rows(ptr(wantp), ptr(wi), NE, NT, desc)
        ok = ok and bool(np.array_equal(idd.cpu().numpy(), wi))
    out["argsort_exact"] = ok
    idx = idd.cpu().numpy()
    wd = torch.zeros((NT, NU), dtype=torch.float32, device="cuda")
    ops.check(L.b200_get_rows_f32_batched(ops.p(wpd), 1, NE, NE, ops.p(idd), NE, ops.p(wd), 1, NU, 1, NU, NT, ops.stream())); torch.cuda.synchronize()
    w = np.take_along_axis(wantp, idx[:, :NU], axis=1)
    out["get_rows_batched_exact"] = bool(np.array_equal(wd.cpu().numpy(), w))
    sd = torch.zeros(NT, dtype=torch.float32, device="cuda")
    ops.check(L.b200_sum_rows(ops.p(wd), NU, ops.p(sd), NU, NT, ops.stream())); torch.cuda.synchronize()
    ws = np.zeros(NT, np.float32); O.orc_sum_rows(ptr(np.ascontiguousarray(w)), ptr(ws), NU, NT)
    out["sum_rows_exact"] = bool(np.array_equal(sd.cpu().numpy(), ws))
    wnd = torch.zeros_like(wd)
    ops.check(L.b200_binary_strided(2, ops.p(wd), ptr(h64(4, 4 * NU, 4 * NU * NT, 4 * NU * NT)), ops.p(sd), ptr(h64(1, NT, 1, 1)), ptr(h64(4, 4, 4 * NT, 4 * NT)), ops.p(wnd),
                                    ptr(h64(NU, NT, 1, 1)), ptr(h64(4, 4 * NU, 4 * NU * NT, 4 * NU * NT)), ops.stream())); torch.cuda.synchronize()
    wn = w / ws[:, None]
    out["div_exact"] = bool(np.array_equal(wnd.cpu().numpy(), wn))
    ex = rng.standard_normal((NT, NU, E)).astype(np.float32); exd = dev(ex); exwd = torch.zeros_like(exd)
    nb3 = h64(4, 4 * E, 4 * E * NU, 4 * E * NU * NT)
    ops.check(L.b200_binary_strided(1, ops.p(exd), ptr(nb3), ops.p(wnd), ptr(h64(1, NU, NT, 1)), ptr(h64(4, 4, 4 * NU, 4 * NU * NT)), ops.p(exwd), ptr(h64(E, NU, NT, 1)), ptr(nb3), ops.stream()))
    torch.cuda.synchronize()
    exw = ex * wn[:, :, None]
    out["mul_bcast_exact"] = bool(np.array_equal(exwd.cpu().numpy(), exw))
    od = torch.zeros((NT, E), dtype=torch.float32, device="cuda")
    sl = h64(4, 4 * E * NU, 4 * E * NU * NT, 4 * E * NU * NT)
    ops.check(L.b200_binary_strided(0, ops.p(exwd), ptr(sl), C.c_void_p(exwd.data_ptr() + 4 * E), ptr(h64(E, NT, 1, 1)), ptr(sl), ops.p(od), ptr(h64(E, NT, 1, 1)),
                                    ptr(h64(4, 4 * E, 4 * E * NT, 4 * E * NT)), ops.stream())); torch.cuda.synchronize()
    out["add_slices_exact"] = bool(np.array_equal(od.cpu().numpy(), exw[:, 0, :] + exw[:, 1, :]))
    xu = (rng.standard_normal(4096) * 5).astype(np.float32); xud = dev(xu); yud = torch.zeros_like(xud)
    worst = 0.0
    for op, sc, bb in ((0, 0.37, 0.0), (0, 2.5, -0.75), (1, 0.0, 0.0), (2, 0.0, 0.0)):
        ops.check(L.b200_unary(op, ops.p(xud), ops.p(yud), xu.size, sc, bb, ops.stream())); torch.cuda.synchronize()
        wu = np.zeros_like(xu); O.orc_unary(op, ptr(xu), ptr(wu), xu.size, sc, bb)
        worst = max(worst, float(np.abs(yud.cpu().numpy() - wu).max() / max(1.0, np.abs(wu).max())))
    out["unary"] = worst
    return out


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


def do_attn_any(kvt, d, n_head, n_head_kv, n_tok, n_kv):
    """FLASH_ATTN_EXT at head sizes the tuned kernels do not carry, F16 / Q8_0 / Q4_0 cache, vs the oracle (F16: vs the f64 value of the same f16 inputs)"""
    rng = np.random.default_rng(kvt * 1000 + d + n_kv)
    rb_row, rb_head = row_bytes(kvt, n_head_kv * d), row_bytes(kvt, d)
    kf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32); vf = rng.standard_normal((n_kv, n_head_kv * d)).astype(np.float32)
    ids = np.arange(n_kv, dtype=np.int64)
    kc = np.zeros((n_kv, rb_row), np.uint8); vc = kc.copy()
    oracle().orc_set_rows(ptr(kf), ptr(ids), ptr(kc), kvt, n_head_kv * d, n_kv, rb_row)
    oracle().orc_set_rows(ptr(vf), ptr(ids), ptr(vc), kvt, n_head_kv * d, n_kv, rb_row)
    q = rng.standard_normal((n_tok, n_head, d)).astype(np.float32)
    npad = (n_tok + 63) // 64 * 64
    mask = np.full((npad, n_kv), -np.inf, np.float32)
    for t in range(n_tok):
        mask[t, :max(1, n_kv - n_tok + t + 1)] = 0
    m16 = mask.astype(np.float16)
    scale = float(1 / np.sqrt(d))
    if kvt == 1:
        Kd = kc.view(np.float16).reshape(n_kv, n_head_kv, d).astype(np.float64); Vd = vc.view(np.float16).reshape(n_kv, n_head_kv, d).astype(np.float64)
        want = np.zeros((n_tok, n_head, d))
        for t in range(n_tok):
            for h in range(n_head):
                sc = Kd[:, h // (n_head // n_head_kv), :] @ q[t, h].astype(np.float16).astype(np.float64) * scale + mask[t].astype(np.float64)
                w = np.exp(sc - sc.max()); w /= w.sum()
                want[t, h] = w @ Vd[:, h // (n_head // n_head_kv), :]
    else:
        want = np.zeros((n_tok, n_head, d), np.float32)
        oracle().orc_flash_attn_ext(ptr(q), n_head * d * 4, d * 4, ptr(kc), rb_row, rb_head, ptr(vc), rb_row, rb_head, ptr(m16.view(np.uint16)), ptr(want),
                                    kvt, d, d, n_head, n_head_kv, n_tok, n_kv, scale, 0.0, 0.0)
    qd = dev(q); md = torch.from_numpy(m16.view(np.int16)).cuda()
    kd = dev(np.concatenate([kc.reshape(-1), np.zeros(64, np.uint8)])); vd = dev(np.concatenate([vc.reshape(-1), np.zeros(64, np.uint8)]))
    dst = torch.zeros((n_tok, n_head, d), dtype=torch.float32, device="cuda")
    ops.check(L.b200_flash_attn_any(kvt, ops.p(qd), n_head * d, d, ops.p(kd), rb_row, rb_head, ops.p(vd), rb_row, rb_head, ops.p(md), n_kv, ops.p(dst),
                                    d, n_head, n_head_kv, n_tok, n_kv, scale, 0.0, 0.0, ops.stream()))
    torch.cuda.synchronize()
    return {"err": rel(dst.cpu().numpy(), want)}


def do_attn_any_suite():
    out = {}
    for case in [(1, 256, 8, 4, 1, 700), (1, 96, 8, 8, 5, 300), (8, 256, 8, 4, 1, 700), (8, 96, 4, 4, 33, 128), (2, 192, 4, 2, 2, 260), (8, 32, 4, 1, 1, 64)]:
        try:
            out[" ".join(map(str, case))] = do_attn_any(*case)
        except Exception as e:                              # noqa: BLE001
            out[" ".join(map(str, case))] = {"error": repr(e)[:300]}
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


def do_repack_models(group):
    return {str(t): do_repack_model(int(t), 2048) for t in str(group).split(",")}


def do_type_suites(group):
    return {str(t): do_type_suite(int(t)) for t in str(group).split(",")}


def main():
    what = sys.argv[1]; a = [int(v) if v.lstrip("-").isdigit() else v for v in sys.argv[2:]]
    torch.cuda.set_device(0)
    fn = {"attn_any_suite": do_attn_any_suite, "repack_models": do_repack_models, "type_suites": do_type_suites, "nofa": do_nofa, "glue": do_glue, "kv_q4_0": do_kv_q4_0, "type_suite": do_type_suite, "mul_mat": do_mul_mat, "mul_mat_id": do_mul_mat_id, "get_rows": do_get_rows, "repack_model": do_repack_model, "executor": do_executor}[what]
    print("RESULT " + json.dumps(fn(*a)))


if __name__ == "__main__":
    main()
