"""The node list of one Mixtral-style block as libllama emits it (llama.cpp/src/llama-graph.cpp build_moe_ffn, softmax gating, normalised weights),
preceded by a token-embedding GET_ROWS, an RMS_NORM * w and one MUL_MAT on a wide-only format.  Shared by the CPU dry-run test
(tests/test_wide_plan.py, fake pointers) and the GPU child (tests/wide_gpu_child.py, real tensors): `A` allocates.
    A.f32(ne) / A.i32(ne) -> G.T of a fresh tensor;  A.named(name, type, ne) -> G.T of an input tensor (weights, ids, tokens)"""


def build(G, A, E, FF, NE, NU, NT, VOC, wtype):
    nl = G.NodeList()
    emb = nl.add(G.OP_GET_ROWS, A.f32([E, NT]), [A.named("table", G.Q4_K, [E, VOC]), A.named("tok", G.I32, [NT])])
    nrm = nl.add(G.OP_RMS_NORM, A.f32([E, NT]), [emb], [G.f32_bits(1e-5)])
    cur0 = nl.add(G.OP_MUL, G.T(nrm.ptr, G.F32, [E, NT]), [nrm, A.named("norm_w", G.F32, [E])])
    cur = nl.add(G.OP_MUL_MAT, A.f32([E, NT]), [A.named("w_proj", wtype, [E, E]), cur0])                 # a wide-only format consumes the norm
    # ---- build_moe_ffn
    logits = nl.add(G.OP_MUL_MAT, A.f32([NE, NT]), [A.named("gate_inp", G.F32, [E, NE]), cur])           # ffn_moe_logits
    probs = nl.add(G.OP_SOFT_MAX, A.f32([NE, NT]), [logits], [G.f32_bits(1.0), G.f32_bits(0.0)])         # ffn_moe_probs
    order = nl.add(G.OP_ARGSORT, A.i32([NE, NT]), [probs], [1])                                          # top_k = argsort(desc) + view
    sel = nl.view_op(G.T(order.ptr, G.I32, [NU, NT], [4, 4 * NE, 4 * NE * NT, 4 * NE * NT]), order)      # ffn_moe_topk [n_used, n_tok], row stride n_expert
    probs3 = nl.view_op(G.T(probs.ptr, G.F32, [1, NE, NT], [4, 4, 4 * NE, 4 * NE * NT]), probs)
    w = nl.add(G.OP_GET_ROWS, A.f32([1, NU, NT]), [probs3, sel])                                         # ffn_moe_weights
    w2 = nl.view_op(G.T(w.ptr, G.F32, [NU, NT]), w)
    wsum = nl.add(G.OP_SUM_ROWS, A.f32([1, NT]), [w2])
    wn = nl.add(G.OP_DIV, A.f32([NU, NT]), [w2, wsum])                                                   # ffn_moe_weights_norm
    wn3 = nl.view_op(G.T(wn.ptr, G.F32, [1, NU, NT], [4, 4, 4 * NU, 4 * NU * NT]), wn)
    x3 = nl.view_op(G.T(cur.ptr, G.F32, [E, 1, NT], [4, 4 * E, 4 * E, 4 * E * NT]), cur)
    up = nl.add(G.OP_MUL_MAT_ID, A.f32([FF, NU, NT]), [A.named("up_exps", G.Q4_K, [E, FF, NE]), x3, sel])
    gate = nl.add(G.OP_MUL_MAT_ID, A.f32([FF, NU, NT]), [A.named("gate_exps", G.Q4_K, [E, FF, NE]), x3, sel])
    act = nl.add(G.OP_GLU_SWIGLU, A.f32([FF, NU, NT]), [gate, up], [2, 0])
    down = nl.add(G.OP_MUL_MAT_ID, A.f32([E, NU, NT]), [A.named("down_exps", G.Q6_K, [FF, E, NE]), act, sel])
    exw = nl.add(G.OP_MUL, A.f32([E, NU, NT]), [down, wn3])                                              # ffn_moe_weighted
    out = None
    for i in range(NU):                                                                                  # sum over the used experts: views [E, n_tok] with row stride n_used * E
        sl = nl.view_op(G.T(exw.ptr + 4 * E * i, G.F32, [E, NT], [4, 4 * E * NU, 4 * E * NU * NT, 4 * E * NU * NT]), exw)
        out = sl if out is None else nl.add(G.OP_ADD, A.f32([E, NT]), [out, sl])
    return nl, out
