"""GPU: the graph executor (include/b200_graph.h) on the node list libllama would emit for a small
Llama / Qwen2-shaped model — with and without fusion, with and without CUDA graphs — against a CPU
forward pass composed of oracle ops on the same weights.  north_star bar: identical argmax,
logits within 1e-3 relative."""
import numpy as np
import pytest

from oracle_model import OracleLlama, causal_mask, make_host_weights
from refutil import F16, Q8_0

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def mods():
    import importlib
    from conftest import load_pkg
    load_pkg()
    return importlib.import_module("llama_box_b200.graph"), importlib.import_module("llama_box_b200.model")


def run_steps(G, M, model, ex, flags, steps, prompt, pos0_after_prompt):
    """prefill `prompt` as one ubatch, then greedy-decode `steps` tokens; returns (logits list, tokens)"""
    outs, toks = [], []
    side = torch.cuda.Stream()          # CUDA graphs cannot capture the legacy default stream
    torch.cuda.synchronize()
    n_ctx_pad = lambda p: max(256, (p + 255) // 256 * 256)  # noqa: E731

    def step(tokens, pos0):
        n = len(tokens); n_kv = n_ctx_pad(pos0 + n)
        nodes, io = model.build(n, n_kv)
        io["tokens"].copy_(torch.tensor(tokens, dtype=torch.int32)); io["pos"].copy_(torch.arange(pos0, pos0 + n, dtype=torch.int32))
        io["kv_idx"].copy_(torch.arange(pos0, pos0 + n, dtype=torch.int64)); io["mask"].copy_(torch.from_numpy(causal_mask(n, n_kv, pos0)))
        io["out_ids"].copy_(torch.tensor([n - 1], dtype=torch.int32))
        with torch.cuda.stream(side):
            ex.compute(nodes, flags)
        torch.cuda.synchronize()
        return io["logits"][0].cpu().numpy().copy()
    lg = step(prompt, 0); outs.append(lg); toks.append(int(lg.argmax()))
    pos = len(prompt)
    for _ in range(steps):
        lg = step([toks[-1]], pos); outs.append(lg); toks.append(int(lg.argmax())); pos += 1
    return outs, toks


@pytest.mark.parametrize("ftype,kv,bias", [("Q4_K_M", F16, False), ("Q4_0", F16, False), ("Q8_0", Q8_0, True)])
def test_executor_vs_oracle_forward(ftype, kv, bias):
    G, M = mods()
    cfg = dict(M.CONFIGS["test-small"]); cfg["qkv_bias"] = bias
    if bias:
        cfg["rope_mode"] = 2
    mix, out_t = M.type_mix(ftype, cfg["n_layer"])
    hw, types = make_host_weights(cfg, mix, out_t)
    n_ctx = 512
    prompt = [5, 17, 300, 4000, 9]
    steps = 3
    # oracle
    orc = OracleLlama(cfg, hw, types, n_ctx, kv)
    want, wtoks = [], []
    pos = 0
    seq = list(prompt)
    lg = orc.forward(np.array(seq, np.int32), np.arange(0, len(seq), dtype=np.int32), np.arange(0, len(seq), dtype=np.int64), 256,
                     causal_mask(len(seq), 256, 0).astype(np.float16).view(np.uint16))
    want.append(lg); wtoks.append(int(lg.argmax())); pos = len(seq)
    for _ in range(steps):
        lg = orc.forward(np.array([wtoks[-1]], np.int32), np.array([pos], np.int32), np.array([pos], np.int64), 256,
                         causal_mask(1, 256, pos).astype(np.float16).view(np.uint16))
        want.append(lg); wtoks.append(int(lg.argmax())); pos += 1
    results = {}
    for name, flags in (("plain", 0), ("fused", G.EXEC_FUSION), ("fused+graphs", G.EXEC_FUSION | G.EXEC_CUDA_GRAPHS),
                        ("mega-attn", G.EXEC_FUSION | G.EXEC_MEGAKERNEL | G.EXEC_CUDA_GRAPHS),
                        ("mega", G.EXEC_FUSION | G.EXEC_MEGAKERNEL | G.EXEC_MEGA_MMV), ("mega+graphs", G.EXEC_FUSION | G.EXEC_MEGAKERNEL | G.EXEC_MEGA_MMV | G.EXEC_CUDA_GRAPHS)):
        model = M.SyntheticLlama(cfg, ftype, n_ctx=n_ctx, kv_type=kv, host_weights=hw)
        ex = G.Executor(0)
        got, gtoks = run_steps(G, M, model, ex, flags, steps, prompt, len(prompt))
        results[name] = (got, gtoks, ex.last_kernels, ex.captures, ex.replays, ex.mk_launches, ex.mk_phases)
        # Q8_0 KV: the oracle accumulates V in f32 -> the north_star bar (1e-3 relative) applies directly.
        # F16 KV: the oracle's V accumulator is fp16 (ggml-cpu/ops.cpp:8278-8340, ~4e-4 relative noise per
        # attention output, see test_flash_attn_f16_closer_to_f64_than_oracle); on random weights that noise is
        # amplified ~10x by the following layers, so the comparison is limited by the oracle, not by us.
        tol = 1e-3 if kv == Q8_0 else 3e-2
        for a, b in zip(got, want):
            assert np.isfinite(a).all()
            assert np.abs(a - b).max() <= tol * np.abs(b).max(), (name, np.abs(a - b).max(), np.abs(b).max())
        assert gtoks == wtoks, name
        ex.close()
    # fusion must cut launches, graphs must capture once and replay
    assert results["fused"][2] < results["plain"][2]
    assert results["fused+graphs"][3] >= 1 and results["fused+graphs"][4] >= 1
    # the execution modes agree with each other much more tightly than with the oracle
    for a, b in zip(results["plain"][0], results["fused+graphs"][0]):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()
    # attention-only programs: one launch per layer and token, each a single phase
    assert results["mega-attn"][5] == results["mega-attn"][6] and results["mega-attn"][5] >= cfg["n_layer"]
    for a, b in zip(results["fused"][0], results["mega-attn"][0]):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()
    if ftype == "Q4_K_M":
        # the persistent decode kernel took the Q4_K / Q6_K decode steps (two launches per token: the chain is cut at
        # the last layer's GET_ROWS) and agrees with the per-op path
        assert results["mega"][5] >= 2 * steps and results["mega"][6] > results["mega"][5]
        assert results["mega"][2] < results["fused"][2]
        for a, b in zip(results["fused"][0], results["mega+graphs"][0]):
            assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()
    elif ftype == "Q8_0":
        # other weight types stay on the per-op matvec kernels (Q4_0 models still have a Q6_K output matrix); the fused
        # rope + KV store + attention phase runs as a one-phase program
        assert results["mega"][5] == results["mega"][6]


def test_executor_rejects_unsupported_node():
    G, M = mods()
    import ctypes as C
    n = G.Node(); n.op = 99
    ex = G.Executor(0)
    assert not ex.supports(n)
    arr = (G.Node * 1)(); arr[0].op = 99
    from conftest import load_pkg
    ops = load_pkg().ops
    assert ops.lib.b200_executor_compute(ex.h, arr, 1, None, 0) == -1
    ex.close()
