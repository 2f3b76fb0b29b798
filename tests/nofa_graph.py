"""The node list of one attention block WITHOUT -fa as libllama emits it (llama.cpp/src/llama-graph.cpp build_attn_mha non-flash branch,
llama-kv-cache-unified.cpp cpy_k / cpy_v with the transposed V cache): projections, ROPE, K rows and V elements stored into the f16 caches,
KQ = K^T Q over the permuted K view (GQA broadcast), SOFT_MAX with the f32 mask, KQV over the transposed V view, CONT of the permuted result, wo.
Shared by the CPU dry-run test and the GPU child; `A` allocates (see tests/moe_graph.py)."""


def build(G, A, E, H, HK, D, NT, KV_SIZE, NKV, rope_params):
    nl = G.NodeList()
    x = A.named("x", G.F32, [E, NT])
    pos = A.named("pos", G.I32, [NT]); kidx = A.named("k_idx", G.I64, [NT]); vidx = A.named("v_idx", G.I64, [NT * HK * D])
    mask = A.named("mask", G.F32, [NKV, 64])
    kc = A.named("k_cache", G.F16, [HK * D, KV_SIZE]); vc = A.named("v_cache", G.F16, [KV_SIZE, HK * D])      # V cache transposed: row = one (head, d), kv_size cells
    q = nl.add(G.OP_MUL_MAT, A.f32([H * D, NT]), [A.named("wq", G.Q4_K, [E, H * D]), x])
    q3 = nl.view_op(q.reshape([D, H, NT]), q)
    qr = nl.add(G.OP_ROPE, A.f32([D, H, NT]), [q3, pos], rope_params)
    k = nl.add(G.OP_MUL_MAT, A.f32([HK * D, NT]), [A.named("wk", G.Q4_K, [E, HK * D]), x])
    k3 = nl.view_op(k.reshape([D, HK, NT]), k)
    kr = nl.add(G.OP_ROPE, A.f32([D, HK, NT]), [k3, pos], rope_params)
    v = nl.add(G.OP_MUL_MAT, A.f32([HK * D, NT]), [A.named("wv", G.Q4_K, [E, HK * D]), x])
    k2 = nl.view_op(kr.reshape([HK * D, NT]), kr)
    nl.add(G.OP_SET_ROWS, kc, [k2, kidx])
    v1 = nl.view_op(G.T(v.ptr, G.F32, [1, NT * HK * D]), v)
    vflat = nl.view_op(G.T(vc.ptr, G.F16, [1, KV_SIZE * HK * D]), vc)
    nl.add(G.OP_SET_ROWS, vflat, [v1, vidx])
    # K view [D, HK, NKV] of the cache rows, permuted to [D, NKV, HK]; Q permuted to [D, NT, H]
    kp = nl.view_op(G.T(kc.ptr, G.F16, [D, NKV, HK], [2, 2 * HK * D, 2 * D, 2 * HK * D * NKV]), kc)
    qp = nl.view_op(G.T(qr.ptr, G.F32, [D, NT, H], [4, 4 * D * H, 4 * D, 4 * D * H * NT]), qr)
    kq = nl.add(G.OP_MUL_MAT, A.f32([NKV, NT, H]), [kp, qp])
    sm = nl.add(G.OP_SOFT_MAX, A.f32([NKV, NT, H]), [kq, mask], [G.f32_bits(D ** -0.5), G.f32_bits(0.0)])
    vp = nl.view_op(G.T(vc.ptr, G.F16, [NKV, D, HK], [2, 2 * KV_SIZE, 2 * KV_SIZE * D, 2 * KV_SIZE * D * HK]), vc)
    kqv = nl.add(G.OP_MUL_MAT, A.f32([D, NT, H]), [vp, sm])
    perm = nl.view_op(G.T(kqv.ptr, G.F32, [D, H, NT], [4, 4 * D * NT, 4 * D, 4 * D * NT * H]), kqv)
    cont = nl.add(G.OP_CONT, A.f32([D * H, NT]), [perm])
    out = nl.add(G.OP_MUL_MAT, A.f32([E, NT]), [A.named("wo", G.Q4_K, [H * D, E]), cont])
    return nl, out
