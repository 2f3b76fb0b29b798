#!/usr/bin/env python3
"""Synthesize a Llama-architecture GGUF with the shapes and quantisation mix of a named model.

TEST / MEASUREMENT INFRASTRUCTURE (no weights ship with the reference, and there is no network):
the file feeds the unmodified reference (oracle/_ref/llama_drv -> libllama + ggml-cpu) and, through
the plug-in, our backend — same bytes, same prompt — for end-to-end token-ID parity and for the
`--impl reference` / `e2e` bench legs.  Tensor names / metadata follow the llama loader
(/root/reference/llama.cpp/src/llama-model.cpp load_tensors LLM_ARCH_LLAMA / LLM_ARCH_QWEN2), the type
mix follows llama-quant.cpp:185-186,203-227,302-364 (see llama_box_b200/model.py:type_mix).

Weights: `--weights blocks` = random *valid* quant blocks (fast; any size), or
`--weights gauss` = N(0, sigma) f32 quantised with the reference quantiser (needs oracle/_ref; small models).
The vocabulary is a synthetic SPM vocab (ids are fed directly; nothing is tokenised).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
from refutil import (BLOCK_BYTES, BLOCK_ELEMS, IQ4_NL, IQ4_XS, Q2_K, Q3_K, Q4_0, Q4_1, Q4_K, Q5_0, Q5_1, Q5_K, Q6_K, Q8_0, rand_blocks,  # noqa: E402
                     row_bytes)

CONFIGS = {
    "llama3-8b": dict(arch="llama", n_embd=4096, n_head=32, n_head_kv=8, head_dim=128, n_ff=14336, n_vocab=128256, n_layer=32, rope_base=500000.0, eps=1e-5, qkv_bias=False, n_ctx_train=8192),
    "tinyllama-1.1b": dict(arch="llama", n_embd=2048, n_head=32, n_head_kv=4, head_dim=64, n_ff=5632, n_vocab=32000, n_layer=22, rope_base=10000.0, eps=1e-5, qkv_bias=False, n_ctx_train=2048),
    "qwen2-72b": dict(arch="qwen2", n_embd=8192, n_head=64, n_head_kv=8, head_dim=128, n_ff=29568, n_vocab=152064, n_layer=80, rope_base=1000000.0, eps=1e-6, qkv_bias=True, n_ctx_train=32768),
    "test-small": dict(arch="llama", n_embd=2048, n_head=16, n_head_kv=4, head_dim=128, n_ff=4096, n_vocab=4096, n_layer=2, rope_base=500000.0, eps=1e-5, qkv_bias=False, n_ctx_train=8192),
    # Mixtral-style mixture of experts on the llama architecture (llama-model.cpp load_tensors LLM_ARCH_LLAMA with n_expert > 0:
    # ffn_gate_inp f32 + ffn_{gate,down,up}_exps; graph: llm_build_llama -> build_moe_ffn): exercises MUL_MAT_ID
    "test-moe": dict(arch="llama", n_embd=2048, n_head=16, n_head_kv=4, head_dim=128, n_ff=2048, n_vocab=4096, n_layer=2, rope_base=500000.0, eps=1e-5, qkv_bias=False, n_ctx_train=8192,
                     n_expert=8, n_expert_used=2),
}
# llama_ftype values (include/llama.h); the single-type ones below put every matrix in that format (the wide path's formats)
FTYPE_ID = {"Q4_0": 2, "Q8_0": 7, "Q4_K_M": 15, "Q4_1": 3, "Q5_1": 9, "Q2_K": 10, "Q3_K_M": 12, "IQ4_NL": 25, "IQ4_XS": 30}
SINGLE_TYPE = {"Q4_0": Q4_0, "Q8_0": Q8_0, "Q4_1": Q4_1, "Q5_1": Q5_1, "Q2_K": Q2_K, "Q3_K_M": Q3_K, "IQ4_NL": IQ4_NL, "IQ4_XS": IQ4_XS}


def use_more_bits(i, n):
    return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def layer_types(ftype, i, n, n_ff=0, is_70b=False):
    """per-layer tensor types of a quantisation mix (llama-quant.cpp:185-186,203-227,302-364) including the reference's fallback
    for rows that are not a multiple of 256 (llama-quant.cpp:442-470: Q4_K -> Q5_0, Q6_K -> Q8_0; Qwen2-72B's ffn_down, n_ff = 29568)
    and the 70B-class bump of attn_v from Q4_K to Q5_K (llama-quant.cpp:305-310)"""
    if ftype == "Q4_K_M":
        hi = Q6_K if use_more_bits(i, n) else Q4_K
        v = hi if hi == Q6_K or not is_70b else Q5_K
        down = hi
        if n_ff % 256 != 0:
            down = Q8_0 if hi == Q6_K else Q5_0
        return dict(attn_q=Q4_K, attn_k=Q4_K, attn_v=v, attn_output=Q4_K, ffn_gate=Q4_K, ffn_up=Q4_K, ffn_down=down)
    t = SINGLE_TYPE[ftype]
    return {k: t for k in ("attn_q", "attn_k", "attn_v", "attn_output", "ffn_gate", "ffn_up", "ffn_down")}


def gauss_quantized(t, m, k, seed, std=0.02):
    """N(0, 0.02^2) f32 weights quantised with the REFERENCE quantiser (ggml_quantize_chunk), generated and converted in
    row chunks on several threads (each chunk has its own counter-based stream, so the file does not depend on the thread count)"""
    from concurrent.futures import ThreadPoolExecutor
    from refutil import ptr, ref
    base, _ = ref()
    out = np.zeros((m, row_bytes(t, k)), dtype=np.uint8)
    rows = max(1, (1 << 22) // k)

    def work(r0):
        r1 = min(m, r0 + rows)
        g = np.random.Generator(np.random.Philox(key=seed, counter=r0))
        w = g.standard_normal((r1 - r0, k), dtype=np.float32) * np.float32(std)
        base.ggml_quantize_chunk(t, ptr(w), ptr(out[r0:r1]), 0, r1 - r0, k, None)
    nthr = max(1, min(32, len(os.sched_getaffinity(0))))
    with ThreadPoolExecutor(nthr) as tp:
        list(tp.map(work, range(0, m, rows)))
    return out


def main():
    import gguf
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="test-small", choices=sorted(CONFIGS))
    ap.add_argument("--ftype", default="Q4_K_M", choices=sorted(FTYPE_ID))
    ap.add_argument("--out", required=True)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--weights", default="blocks", choices=["blocks", "gauss"])
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--residual-scale", action="store_true")
    ap.add_argument("--scale-mul", type=float, default=0.25)
    a = ap.parse_args()
    c = dict(CONFIGS[a.config])
    if a.layers:
        c["n_layer"] = a.layers
    E, H, HK, D, FF, V, L = c["n_embd"], c["n_head"], c["n_head_kv"], c["head_dim"], c["n_ff"], c["n_vocab"], c["n_layer"]
    rng = np.random.default_rng(a.seed)
    qt = {2: gguf.GGMLQuantizationType.Q4_0, 6: gguf.GGMLQuantizationType.Q5_0, 8: gguf.GGMLQuantizationType.Q8_0, 12: gguf.GGMLQuantizationType.Q4_K,
          13: gguf.GGMLQuantizationType.Q5_K, 14: gguf.GGMLQuantizationType.Q6_K, 3: gguf.GGMLQuantizationType.Q4_1, 7: gguf.GGMLQuantizationType.Q5_1,
          10: gguf.GGMLQuantizationType.Q2_K, 11: gguf.GGMLQuantizationType.Q3_K, 20: gguf.GGMLQuantizationType.IQ4_NL, 23: gguf.GGMLQuantizationType.IQ4_XS}
    NE, NU = c.get("n_expert", 0), c.get("n_expert_used", 0)

    w = gguf.GGUFWriter(a.out, c["arch"])
    w.add_name(f"synthetic-{a.config}-{a.ftype}")
    w.add_context_length(c["n_ctx_train"]); w.add_embedding_length(E); w.add_block_count(L); w.add_feed_forward_length(FF)
    w.add_head_count(H); w.add_head_count_kv(HK); w.add_rope_dimension_count(D); w.add_rope_freq_base(c["rope_base"])
    w.add_layer_norm_rms_eps(c["eps"]); w.add_vocab_size(V); w.add_file_type(FTYPE_ID[a.ftype])
    if NE:
        w.add_expert_count(NE); w.add_expert_used_count(NU)
    # synthetic SPM vocabulary: <unk>, <s>, </s>, 256 byte tokens, then plain pieces
    toks, scores, types = [], [], []
    for i in range(V):
        if i == 0: t, ty = "<unk>", gguf.TokenType.UNKNOWN
        elif i == 1: t, ty = "<s>", gguf.TokenType.CONTROL
        elif i == 2: t, ty = "</s>", gguf.TokenType.CONTROL
        elif i < 259: t, ty = f"<0x{i - 3:02X}>", gguf.TokenType.BYTE
        else: t, ty = f"▁w{i}", gguf.TokenType.NORMAL
        toks.append(t.encode()); scores.append(-float(i)); types.append(int(ty))
    w.add_tokenizer_model("llama"); w.add_token_list(toks); w.add_token_scores(scores); w.add_token_types(types)
    w.add_bos_token_id(1); w.add_eos_token_id(2); w.add_unk_token_id(0); w.add_add_bos_token(False)

    cache = {}

    def qblocks(t, m, k, std=0.02):
        key = (t, m, k, std)
        if key not in cache or a.weights == "gauss":
            if a.weights == "gauss":
                cache[key] = gauss_quantized(t, m, k, int(rng.integers(1 << 31)), std)
            else:
                cache[key] = rand_blocks(rng, t, m, k, a.scale_mul)
        return cache[key]

    # --residual-scale: GPT-2 / LLaMA-style initialisation of the projections that write into the residual stream
    # (attn_output, ffn_down): std 0.02 / sqrt(2 * n_layer of the FULL model), so that a layer's update is a fraction of the
    # stream, as in trained models — a 2-layer cut of the 32-layer 8B keeps the 32-layer scaling
    res_std = 0.02 / np.sqrt(2.0 * CONFIGS[a.config]["n_layer"]) if a.residual_scale else 0.02

    def add_q(name, t, m, k):
        std = res_std if (name.endswith("attn_output.weight") or name.endswith("ffn_down.weight")) else 0.02
        w.add_tensor(name, qblocks(t, m, k, std), raw_dtype=qt[t])

    def add_q3(name, t, ne, m, k):
        # expert stack [k, m, n_expert] (llama-model.cpp: ffn_*_exps)
        data = np.stack([qblocks(t, m, k, res_std if "down" in name else 0.02) if a.weights == "gauss" else rand_blocks(rng, t, m, k, a.scale_mul) for _ in range(ne)])
        w.add_tensor(name, data, raw_dtype=qt[t])

    emb_t = Q4_K if a.ftype == "Q4_K_M" else SINGLE_TYPE[a.ftype]
    out_t = Q8_0 if a.ftype == "Q8_0" else Q6_K
    add_q("token_embd.weight", emb_t, V, E)
    for i in range(L):
        ts = layer_types(a.ftype, i, L, FF, a.config == "qwen2-72b")
        w.add_tensor(f"blk.{i}.attn_norm.weight", (1 + 0.05 * rng.standard_normal(E)).astype(np.float32))
        add_q(f"blk.{i}.attn_q.weight", ts["attn_q"], H * D, E)
        add_q(f"blk.{i}.attn_k.weight", ts["attn_k"], HK * D, E)
        add_q(f"blk.{i}.attn_v.weight", ts["attn_v"], HK * D, E)
        if c["qkv_bias"]:
            w.add_tensor(f"blk.{i}.attn_q.bias", (0.1 * rng.standard_normal(H * D)).astype(np.float32))
            w.add_tensor(f"blk.{i}.attn_k.bias", (0.1 * rng.standard_normal(HK * D)).astype(np.float32))
            w.add_tensor(f"blk.{i}.attn_v.bias", (0.1 * rng.standard_normal(HK * D)).astype(np.float32))
        add_q(f"blk.{i}.attn_output.weight", ts["attn_output"], E, H * D)
        w.add_tensor(f"blk.{i}.ffn_norm.weight", (1 + 0.05 * rng.standard_normal(E)).astype(np.float32))
        if NE:
            w.add_tensor(f"blk.{i}.ffn_gate_inp.weight", (0.5 * rng.standard_normal((NE, E))).astype(np.float32))
            add_q3(f"blk.{i}.ffn_gate_exps.weight", ts["ffn_gate"], NE, FF, E)
            add_q3(f"blk.{i}.ffn_down_exps.weight", ts["ffn_down"], NE, E, FF)
            add_q3(f"blk.{i}.ffn_up_exps.weight", ts["ffn_up"], NE, FF, E)
            continue
        add_q(f"blk.{i}.ffn_gate.weight", ts["ffn_gate"], FF, E)
        add_q(f"blk.{i}.ffn_up.weight", ts["ffn_up"], FF, E)
        add_q(f"blk.{i}.ffn_down.weight", ts["ffn_down"], E, FF)
    w.add_tensor("output_norm.weight", (1 + 0.05 * rng.standard_normal(E)).astype(np.float32))
    add_q("output.weight", out_t, V, E)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    print(f"wrote {a.out}: {os.path.getsize(a.out) / 1e9:.3f} GB, {L} layers, {a.ftype}")


if __name__ == "__main__":
    main()
