"""CPU: the executor's fusion logic on node lists shaped like libllama's decode graph, through the dry-run entry point
b200_executor_plan (include/b200_graph.h) — no device is touched, pointers are fake addresses.

The list mirrors what the plug-in hands over for one Llama layer (llama-model.cpp:6004-6095, dump of the real graph in
DESIGN.md §5): view nodes included, and — as ggml-alloc does — the buffer of the un-roped Q is recycled for K and V."""
import importlib

import pytest


def mods():
    from conftest import load_pkg
    load_pkg()
    return importlib.import_module("llama_box_b200.graph")


def llama_layers(G, n_layer, recycle, n_tok=1, kv=None, rope_mode=0, bias=False):
    kv = G.F16 if kv is None else kv
    E, H, HK, D, FF, CTX, NKV = 4096, 32, 8, 128, 14336, 4096, 256
    nl = G.NodeList()
    base = [0x10000000]

    def buf(nbytes):                                   # a fresh "allocation"
        p = base[0]
        base[0] += (nbytes + 0xfffff) & ~0xfffff
        return p
    f32 = lambda ptr, ne: G.T(ptr, G.F32, ne)  # noqa: E731
    W = lambda t, m, k: G.T(buf(m * k), t, [k, m])  # noqa: E731
    pos = G.T(buf(64), G.I32, [n_tok]); idx = G.T(buf(64), G.I64, [n_tok])
    mask32 = G.T(buf(4 * NKV * 64), G.F32, [NKV, 64])
    inp = f32(buf(4 * E * n_tok), [E, n_tok])
    rope_params = [0, D, rope_mode, 0, 8192, G.f32_bits(5e5), G.f32_bits(1.0), G.f32_bits(0.0), G.f32_bits(1.0), G.f32_bits(32.0), G.f32_bits(1.0)]
    kvrow = G.row_size(kv, HK * D); kvhead = G.row_size(kv, D)
    mask16 = None
    for il in range(n_layer):
        nrm = nl.add(G.OP_RMS_NORM, f32(buf(4 * E * n_tok), [E, n_tok]), [inp], [G.f32_bits(1e-5)])
        cur = nl.add(G.OP_MUL, f32(nrm.ptr, [E, n_tok]), [nrm, f32(buf(4 * E), [E])])
        scratch = buf(4 * H * D * n_tok)               # ggml-alloc: Q's buffer, free again once ROPE(Q) has read it
        q = nl.add(G.OP_MUL_MAT, f32(scratch, [H * D, n_tok]), [W(G.Q4_K, H * D, E), cur])
        if bias:
            q = nl.add(G.OP_ADD, f32(buf(4 * H * D * n_tok), [H * D, n_tok]), [q, f32(buf(4 * H * D), [H * D])])
        q3 = nl.view_op(q.reshape([D, H, n_tok]), q)
        qr = nl.add(G.OP_ROPE, f32(buf(4 * H * D * n_tok), [D, H, n_tok]), [q3, pos], rope_params)
        k = nl.add(G.OP_MUL_MAT, f32(scratch if recycle else buf(4 * HK * D * n_tok), [HK * D, n_tok]), [W(G.Q4_K, HK * D, E), cur])
        k3 = nl.view_op(k.reshape([D, HK, n_tok]), k)
        kr = nl.add(G.OP_ROPE, f32(buf(4 * HK * D * n_tok), [D, HK, n_tok]), [k3, pos], rope_params)
        v = nl.add(G.OP_MUL_MAT, f32(scratch if recycle else buf(4 * HK * D * n_tok), [HK * D, n_tok]), [W(G.Q6_K, HK * D, E), cur])
        v3 = nl.view_op(v.reshape([D, HK, n_tok]), v)
        kc = G.T(buf(CTX * kvrow), kv, [HK * D, CTX]); vc = G.T(buf(CTX * kvrow), kv, [HK * D, CTX])
        k2 = nl.view_op(kr.reshape([HK * D, n_tok]), kr)
        nl.add(G.OP_SET_ROWS, kc, [k2, idx])
        v2 = nl.view_op(v3.reshape([HK * D, n_tok]), v3)
        nl.add(G.OP_SET_ROWS, vc, [v2, idx])
        q4 = nl.view_op(qr.reshape([D, H, n_tok]), qr)
        qp = nl.view_op(q4.view([D, n_tok, H], [4, 4 * D * H, 4 * D, 4 * D * H * n_tok]), q4)
        es = G.ELEM_SIZE.get(kv, 34)
        kv1 = nl.view_op(kc.view([D, HK, NKV], [es, kvhead, kvrow, kvrow * NKV]), kc)
        kvp = nl.view_op(kv1.view([D, NKV, HK], [es, kvrow, kvhead, kvrow * NKV]), kv1)
        vv1 = nl.view_op(vc.view([D, HK, NKV], [es, kvhead, kvrow, kvrow * NKV]), vc)
        vvp = nl.view_op(vv1.view([D, NKV, HK], [es, kvrow, kvhead, kvrow * NKV]), vv1)
        if mask16 is None:                             # layer 0: the f16 cast of the mask sits right before the attention,
            mask16 = nl.add(G.OP_CPY, G.T(nrm.ptr, G.F16, [NKV, 64]), [mask32])   # in the recycled buffer of attn_norm
        att = nl.add(G.OP_FLASH_ATTN_EXT, f32(buf(4 * H * D * n_tok), [D, H, n_tok]), [qp, kvp, vvp, mask16], [G.f32_bits(D ** -0.5), G.f32_bits(0.0), G.f32_bits(0.0), 10])
        att2 = nl.view_op(att.reshape([H * D, n_tok]), att)
        wo = nl.add(G.OP_MUL_MAT, f32(buf(4 * E * n_tok), [E, n_tok]), [W(G.Q4_K, E, H * D), att2])
        ffn_inp = nl.add(G.OP_ADD, f32(wo.ptr, [E, n_tok]), [wo, inp])
        n2 = nl.add(G.OP_RMS_NORM, f32(buf(4 * E * n_tok), [E, n_tok]), [ffn_inp], [G.f32_bits(1e-5)])
        c2 = nl.add(G.OP_MUL, f32(n2.ptr, [E, n_tok]), [n2, f32(buf(4 * E), [E])])
        gate = nl.add(G.OP_MUL_MAT, f32(buf(4 * FF * n_tok), [FF, n_tok]), [W(G.Q4_K, FF, E), c2])
        up = nl.add(G.OP_MUL_MAT, f32(buf(4 * FF * n_tok), [FF, n_tok]), [W(G.Q4_K, FF, E), c2])
        h = nl.add(G.OP_GLU_SWIGLU, f32(buf(4 * FF * n_tok), [FF, n_tok]), [gate, up], [2, 0])
        dn = nl.add(G.OP_MUL_MAT, f32(n2.ptr, [E, n_tok]), [W(G.Q6_K, E, FF), h])
        inp = nl.add(G.OP_ADD, f32(n2.ptr, [E, n_tok]), [dn, ffn_inp])
    return nl.build()


@pytest.mark.parametrize("recycle", [False, True])
@pytest.mark.parametrize("bias,rope_mode", [(False, 0), (True, 2)])
def test_decode_layer_is_five_launches_even_with_recycled_buffers(recycle, bias, rope_mode):
    G = mods()
    L = 3
    nodes = llama_layers(G, L, recycle, bias=bias, rope_mode=rope_mode)
    fused = G.plan(nodes, G.EXEC_FUSION)
    plain = G.plan(nodes, 0)
    # QKV(+norm[,+bias]) | rope + KV store + attention | wo(+residual) | gate/up(+norm,+SwiGLU) | down(+residual), + one mask cast
    assert fused == 5 * L + 1, (fused, plain)
    # one kernel per compute node (a biased projection is matvec + ADD), with the activation quantiser shared by Q/K/V and gate/up
    assert plain > 3 * fused


@pytest.mark.parametrize("recycle", [False, True])
def test_persistent_kernel_program_is_a_few_launches_per_list(recycle):
    G = mods()
    nodes = llama_layers(G, 4, recycle=recycle)
    # attention + matvec phases in one persistent launch; the layer-0 mask cast (and the rope it interrupts) split the list once
    assert G.plan(nodes, G.EXEC_FUSION | G.EXEC_MEGAKERNEL | G.EXEC_MEGA_MMV) <= 6
    # attention phases only: one single-phase program per layer next to the four matvec launches
    assert G.plan(nodes, G.EXEC_FUSION | G.EXEC_MEGAKERNEL) <= 5 * 4 + 3


def test_prefill_batch_does_not_take_the_decode_fusions():
    G = mods()
    nodes = llama_layers(G, 1, recycle=True, n_tok=32)
    n = G.plan(nodes, G.EXEC_FUSION)
    assert n > 8        # batched MUL_MAT per projection, rope + KV store fused, attention on its own


def test_plan_rejects_unsupported_nodes():
    G = mods()
    arr = (G.Node * 1)(); arr[0].op = 99
    assert G.plan(arr, G.EXEC_FUSION) < 0
