#!/usr/bin/env python3
"""bench.py — decode tokens/s of the GGUF-quantised hot path on B200 (BASELINE.json metric).

A "step" is one batch-1 decode token of Llama-3-8B Q4_K_M (synthetic random-valid GGUF blocks of that
architecture — no weights ship with the reference and there is no network): the full node list libllama
emits for one token (32 layers + lm_head, 7 quantised matvecs / layer, rms_norm, rope, KV store,
flash-attention over an F16 KV cache), run by the graph executor on hand-written sm_100a kernels.

  value        : tokens/s with every input resident in HBM (token feedback through an on-device argmax)
  e2e          : same metric through the public executor API with HOST inputs (token id, position, KV
                 index, mask row: pinned H2D each step) and the logits read back D2H + host argmax —
                 what libllama does per token (llama-context.cpp:1123-1132)
  roofline     : the dominant kernel (mmvq) — algorithmic weight bytes of one token / CUDA-event time of
                 exactly the matvec launches of one token (weights 4.6 GB >> 126 MB L2, so nothing is
                 cache-resident between launches), against MEASURED_PEAKS.json
  cpu_baseline : the reference's own ggml-cpu path (oracle/_ref: unmodified libllama + libggml-cpu built
                 from /root/reference) on the box's host cores, same architecture / quant mix / prompt shape
  --impl reference : that CPU path as its own arm

N > 1 (torchrun): layer split as the reference's default LLAMA_SPLIT_MODE_LAYER does it — rank r owns a
contiguous range of layers (and its KV), the hidden state is handed to rank r+1 with one NCCL send/recv;
N sequences are kept in flight so every GPU streams its slice of the weights once per pipeline tick
("scaling": "strong" — the work per counted unit, one token through all layers, is fixed while N grows; each GPU
streams 1/N of it per tick.  See DESIGN.md, multi-GPU).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--n-past", type=int, default=512, help="KV positions already in the cache when decoding starts")
    ap.add_argument("--layers", type=int, default=0, help="override layer count (debug only; makes the number INVALID)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--mega", action="store_true", help="experimental: attention as a phase of the persistent decode kernel")
    ap.add_argument("--mega-mmv", action="store_true", help="experimental: attention and the matvecs inside the persistent decode kernel")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, gpu_index=0):
        self.proc, self.lines, self.gpu, self.t_begin = None, [], gpu_index, None

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", str(int(os.environ.get("B200_BENCH_CLOCK_MS", "100"))), "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark_begin(self):
        """the sampler is started before the warm-up (nvidia-smi needs ~0.1 s to produce its first line); samples taken
        from here on are the ones inside the timed region"""
        self.t_begin = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        inside = [ln for (t, ln) in self.lines if self.t_begin is None or self.t_begin <= t <= t_end + 0.02]
        window = "timed region"
        if not inside:                                   # region shorter than the sampling period: fall back to the loaded warm-up samples
            inside = [ln for (t, ln) in self.lines][-3:]; window = "warm-up (timed region shorter than one sample)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference arm
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def cpu_reference(args, gen=16, prompt_len=None):
    """the UNMODIFIED reference on the host cores: oracle/_ref/llama_drv (libllama + ggml-cpu) on a synthetic GGUF
    of the same architecture / quant mix.  Returns (tok/s, cores, sample description) or None."""
    drv = os.path.join(REF_DIR, "llama_drv")
    if not os.path.exists(drv):
        return None
    prompt_len = args.n_past if prompt_len is None else prompt_len
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(shm, f"b200_bench_{args.model}_{args.ftype}{'_L%d' % args.layers if args.layers else ''}.gguf")
    if not os.path.exists(path):
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "make_gguf.py"), "--config", args.model, "--ftype", args.ftype, "--out", path]
        if args.layers:
            cmd += ["--layers", str(args.layers)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-2000:]); return None
    cores = os.cpu_count() or 1
    threads = min(cores, 64)          # ggml-cpu decode is memory-bound; more threads than ~64 do not help and often hurt
    env = dict(os.environ, LD_LIBRARY_PATH=REF_DIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([drv, "--model", path, "--ngl", "0", "--threads", str(threads), "--ctx", str(args.ctx), "--prompt-len", str(prompt_len),
                        "--gen", str(gen)], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-2000:]); return None
    out = json.loads(r.stdout.strip().splitlines()[-1])
    return out["decode_tps"], threads, f"{gen - 1} greedy decode tokens after a {prompt_len}-token prompt, {args.model} {args.ftype} synthetic GGUF, ggml-cpu {threads} threads of {cores} cores", out


def libllama_plugin(args, gen=65):
    """the drop-in path proper: the unmodified reference libllama (llama_decode loop of oracle/drivers/llama_drv.cpp) driving
    libggml-b200.so through ggml's backend C-ABI on the same synthetic GGUF, flash attention on, F16 KV.  Informational
    (`e2e_libllama`); runs after the timed legs, in its own process.  None where oracle/_ref or the plug-in are absent."""
    drv = os.path.join(REF_DIR, "llama_drv"); plugin = os.path.join(ROOT, "llama-box_b200", "libggml-b200.so")
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(shm, f"b200_bench_{args.model}_{args.ftype}{'_L%d' % args.layers if args.layers else ''}.gguf")
    if not (os.path.exists(drv) and os.path.exists(plugin) and os.path.exists(path)):
        return None
    env = dict(os.environ, LD_LIBRARY_PATH=REF_DIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), GGML_BACKEND_PATH=plugin)
    try:
        r = subprocess.run([drv, "--model", path, "--plugin", plugin, "--ngl", "99", "--fa", "--ctx", str(args.ctx), "--prompt-len", str(min(args.n_past, 512)), "--gen", str(gen)],
                           capture_output=True, text=True, env=env, timeout=600)
        if r.returncode != 0:
            return {"value": None, "unit": "tok/s", "error": (r.stderr or r.stdout)[-300:]}
        out = json.loads(r.stdout.strip().splitlines()[-1])
        return {"value": out["decode_tps"], "unit": "tok/s", "prefill_tok_s": out["prefill_tps"],
                "api": "llama_decode of the unmodified reference libllama + libggml-b200.so (ggml backend C-ABI), greedy, %d tokens after a %d-token prompt" % (gen - 1, out["prompt_len"])}
    except Exception as e:  # noqa: BLE001 — informational leg, never fails the bench
        return {"value": None, "unit": "tok/s", "error": str(e)[:300]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    res = cpu_reference(args, gen=max(2, args.steps + 1), prompt_len=args.n_past)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/llama_drv (reference build) or gguf generation missing on this box"}))
        return
    tps, threads, sample, raw = res
    line = {"metric": "decode tok/s Llama-3-8B Q4_K_M bs=1", "value": tps, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / tps if tps else None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "q4_K x q8_K int8 dot, f32 accumulate",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.model} {args.ftype} batch-1 decode, -c {args.ctx}, n_past {args.n_past}, F16 KV", "backend": "ggml-cpu (unmodified reference build)"},
            "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from conftest import load_pkg
    pkg = load_pkg()
    import importlib
    G = importlib.import_module("llama_box_b200.graph"); M = importlib.import_module("llama_box_b200.model")
    ops = pkg.ops

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        return run_b200_pipeline(args, G, M, ops, rank, world, local)

    cfg = args.model
    model = M.SyntheticLlama(cfg, args.ftype, n_ctx=args.ctx, kv_type=G.F16, n_layer=args.layers or None)
    c = model.c
    V = c["n_vocab"]
    # KV cache content for the n_past positions already "decoded": random f16 in a sane range
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    for ly in model.layers:
        for key in ("k_cache", "v_cache"):
            n = args.n_past * G.row_size(G.F16, c["n_head_kv"] * c["head_dim"]) // 2
            ly[key].view(torch.float16)[:n] = (torch.randn(n, device="cuda", generator=gen) * 0.5).half()
    ex = G.Executor(local)
    flags = (0 if args.no_graphs else G.EXEC_CUDA_GRAPHS) | (0 if args.no_fusion else G.EXEC_FUSION) | ((G.EXEC_MEGAKERNEL if (args.mega or args.mega_mmv) else 0) | (G.EXEC_MEGA_MMV if args.mega_mmv else 0) if not args.no_fusion else 0)
    stream = torch.cuda.Stream()
    total = args.warmup + args.steps
    pad = lambda p: (p + 256) // 256 * 256 if True else p  # noqa: E731
    n_kv_of = lambda pos: max(256, (pos + 1 + 255) // 256 * 256)  # noqa: E731
    n_kv_max = n_kv_of(args.n_past + 2 * total + 2)
    assert n_kv_max <= args.ctx, "ctx too small for n_past + steps"
    # per-step input tables, resident in HBM (the `value` leg) and in pinned host memory (the `e2e` leg)
    pos_tab = torch.arange(args.n_past, args.n_past + 2 * total + 2, dtype=torch.int32)
    idx_tab = pos_tab.to(torch.int64)
    pos_dev, idx_dev = pos_tab.cuda(), idx_tab.cuda()
    neg = torch.full((n_kv_max,), float("-inf"))
    graphs = {}

    def nodes_for(n_kv):
        if n_kv not in graphs:
            graphs[n_kv] = model.build(1, n_kv)
        return graphs[n_kv]

    def mask_row(pos, n_kv):
        m = neg[:n_kv].clone(); m[:pos + 1] = 0
        return m
    mask_dev = {}

    def device_step(i):
        """everything already in HBM: 3 tiny D2D copies + the token graph + on-device argmax feeding the next step"""
        pos = args.n_past + i
        n_kv = n_kv_of(pos)
        nodes, io = nodes_for(n_kv)
        io["pos"].copy_(pos_dev[i:i + 1], non_blocking=True)
        io["kv_idx"].copy_(idx_dev[i:i + 1], non_blocking=True)
        io["mask"][0].copy_(mask_dev[i], non_blocking=True)
        ex.compute(nodes, flags, stream=C.c_void_p(stream.cuda_stream))
        ops.check(ops.lib.b200_argmax_f32(ops.p(io["logits"]), ops.p(io["tokens"]), V, 1, C.c_void_p(stream.cuda_stream)))

    # ---- value leg -------------------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for i in range(2 * total + 2):
            mask_dev[i] = mask_row(args.n_past + i, n_kv_of(args.n_past + i)).cuda()
        nodes, io = nodes_for(n_kv_of(args.n_past))
        io["tokens"].fill_(1); io["out_ids"].fill_(0)
        sampler = ClockSampler(local); sampler.start()
        for i in range(args.warmup):
            device_step(i)
        stream.synchronize()
        sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.lib.b200_kernel_launches()
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(args.warmup, total):
            device_step(i)
        e1.record(stream)
        torch.cuda.synchronize()
        launches = ops.lib.b200_kernel_launches() - l0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
    ms_per_step = ms / args.steps
    tps = 1000.0 / ms_per_step
    kernels_per_step = ex.last_kernels + 1

    # ---- roofline leg: exactly the matvec launches of one token, CUDA events on the launching stream
    hbm_peak, peak_src = peaks()
    roof = mmvq_roofline(args, model, ops, G, stream, hbm_peak)
    mid_pos = args.n_past + args.warmup + args.steps // 2
    step_bytes = model.streamed_weight_bytes() + n_kv_of(mid_pos) * model.kv_bytes_per_pos() + model.kv_bytes_per_pos()
    roof["peak_source"] = peak_src
    roof["whole_step"] = {"bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
                          "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak, "frac_of_8TBs": step_bytes / (ms_per_step * 1e-3) / 8e12}

    # ---- e2e leg: host inputs in, logits out, host argmax (the libllama per-token flow) ---------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, model, ex, ops, G, stream, flags, n_kv_of, nodes_for, total, V)

    # ---- CPU baseline (reference build on the host cores), bounded sample ------------------------
    cpu = None
    if not args.no_cpu_baseline:
        res = cpu_reference(args, gen=17, prompt_len=min(args.n_past, 128))
        if res:
            cpu = {"value": res[0], "unit": "tok/s", "cores": res[1], "kind": "reference", "sample": res[2]}
        else:
            cpu = {"value": None, "unit": "tok/s", "cores": 0, "kind": "reference", "sample": "unavailable: oracle/_ref not present on this box"}

    line = {"metric": "decode tok/s Llama-3-8B Q4_K_M bs=1", "value": tps, "unit": "tok/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "q4_K/q6_K x q8_K int8 dot (dp4a), f32 accumulate", "data": "synthetic",
            "config": {"workload": f"{args.model} {args.ftype} batch-1 decode, -c {args.ctx}, n_past {args.n_past}, F16 KV, flash-attn",
                       "n_layer": len(model.layers), "streamed_weight_bytes": model.streamed_weight_bytes(), "kv_bytes_per_pos": model.kv_bytes_per_pos(),
                       "l2_policy": "inputs (4.6 GB of weights per step) larger than L2; no flush needed", "cuda_graphs": not args.no_graphs, "fusion": not args.no_fusion, "persistent_decode_kernel": ("attention+matvec" if args.mega_mmv else ("attention" if args.mega else "off")),
                       "parallelism": "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "kernels_per_step": int(kernels_per_step),
            "graph_captures": int(ex.captures), "graph_replays": int(ex.replays), "roofline": roof, "cpu_baseline": cpu}
    if not args.no_cpu_baseline:
        line["e2e_libllama"] = libllama_plugin(args)   # own process, own copy of the model (4.9 GB more HBM)
    if args.layers:
        line["config"]["INVALID"] = "layer count overridden (debug run)"
    print(json.dumps(line))


def mmvq_roofline(args, model, ops, G, stream, hbm_peak):
    """time exactly the quantised matvec launches of one decode token (fused forms, real weights, all layers),
    CUDA events on the launching stream; the 4.6 GB working set defeats L2 between launches"""
    import torch
    c = model.c
    E, FF = c["n_embd"], c["n_ff"]
    st = C.c_void_p(stream.cuda_stream)
    x = torch.randn(1, max(E, FF), device="cuda")
    acts = {}
    for k in {E, FF}:
        a = torch.zeros(ops.act_col_bytes(0, k) + 64, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(stream):
            ops.check(ops.lib.b200_quantize_act(0, ops.p(x), max(E, FF), ops.p(a), k, 1, st))
        acts[k] = a
    q = torch.zeros(c["n_head"] * c["head_dim"], device="cuda"); kk = torch.zeros(c["n_head_kv"] * c["head_dim"], device="cuda"); v = torch.zeros_like(kk)
    h = torch.zeros(FF, device="cuda"); o = torch.zeros(E, device="cuda"); lg = torch.zeros(c["n_vocab"], device="cuda")

    def one_token():
        n = 0
        for ly in model.layers:
            descs = (ops.MmvDesc * 3)()
            for d, (w, dst) in zip(descs, ((ly["wq"], q), (ly["wk"], kk), (ly["wv"], v))):
                d.W, d.dst, d.bias, d.m, d.type = w.buf.data_ptr(), dst.data_ptr(), None, w.m, w.type
            ops.check(ops.lib.b200_mul_mat_vec_q_multi(descs, 3, ops.p(acts[E]), None, E, 1, st))
            ops.check(ops.lib.b200_mul_mat_vec_q(ly["wo"].type, ops.p(ly["wo"].buf), ops.p(acts[E]), ops.p(o), E, None, ops.p(o), E, E, 1, st))
            ops.check(ops.lib.b200_mul_mat_vec_q_swiglu(ly["gate"].type, ops.p(ly["gate"].buf), ly["up"].type, ops.p(ly["up"].buf), ops.p(acts[E]), None, ops.p(h), FF, E, 1, st))
            ops.check(ops.lib.b200_mul_mat_vec_q(ly["down"].type, ops.p(ly["down"].buf), ops.p(acts[FF]), ops.p(o), E, None, ops.p(o), E, FF, 1, st))
            n += 4
        ops.check(ops.lib.b200_mul_mat_vec_q(model.output.type, ops.p(model.output.buf), ops.p(acts[E]), ops.p(lg), c["n_vocab"], None, None, c["n_vocab"], E, 1, st))
        return n + 1
    with torch.cuda.stream(stream):
        for _ in range(3):
            nl = one_token()
        stream.synchronize()
        # replayed from a CUDA graph so the measurement is device-bound, as in the decode step itself
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            one_token()
    with torch.cuda.stream(stream):
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        stream.synchronize()
        e0.record(stream)
        for _ in range(reps):
            g.replay()
        e1.record(stream)
        stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wbytes = sum(sum(ly[k].nbytes for k in ("wq", "wk", "wv", "wo", "gate", "up", "down")) for ly in model.layers) + model.output.nbytes
    ach = wbytes / (ms * 1e-3) / 1e9
    # DRAM bytes per algorithmic byte of the QKV / wo / gate+up / down launches in the round's `ncu --set full` capture
    # (profiles/r1_mmvq_ncu_full_metrics.tsv: 123.60 MB moved for 122.68 MB of weights; the surplus is activations + outputs)
    dram_per_alg = 1.0075
    return {"bound": "hbm", "kernel": "mmvq_kernel (quantised matvec, all %d launches of one token)" % nl, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "frac_of_8TBs": ach / 8000.0, "traffic": int(wbytes * dram_per_alg), "traffic_source": "ncu dram__bytes_read+write ratio x bytes_per_token (profiles/)",
            "bytes_per_token": wbytes, "launches_per_token": nl,
            "avg_launch_us": ms * 1e3 / nl, "ms_per_token_matvec_only": ms}


def run_e2e(args, model, ex, ops, G, stream, flags, n_kv_of, nodes_for, total, V):
    import numpy as np
    import torch
    st = C.c_void_p(stream.cuda_stream)
    n_kv_max = n_kv_of(args.n_past + 2 * total + 2)
    h_tok = torch.zeros(1, dtype=torch.int32).pin_memory(); h_pos = torch.zeros(1, dtype=torch.int32).pin_memory(); h_idx = torch.zeros(1, dtype=torch.int64).pin_memory()
    h_mask = torch.full((n_kv_max,), float("-inf")).pin_memory(); h_logits = torch.zeros(V).pin_memory()
    h_tok[0] = 1

    def step(i):
        pos = args.n_past + total + i        # continue after the value leg's positions
        n_kv = n_kv_of(pos)
        nodes, io = nodes_for(n_kv)
        h_pos[0] = pos; h_idx[0] = pos; h_mask[:n_kv].fill_(float("-inf")); h_mask[:pos + 1] = 0
        with torch.cuda.stream(stream):
            io["tokens"].copy_(h_tok, non_blocking=True); io["pos"].copy_(h_pos, non_blocking=True); io["kv_idx"].copy_(h_idx, non_blocking=True)
            io["mask"][0].copy_(h_mask[:n_kv], non_blocking=True)
            ex.compute(nodes, flags, stream=st)
            h_logits.copy_(io["logits"][0], non_blocking=True)
        stream.synchronize()
        h_tok[0] = int(torch.argmax(h_logits))          # greedy sampling on the host (httpserver.hpp:4285-4299)
        return 4 + 4 + 8 + n_kv * 4, V * 4
    for i in range(min(args.warmup, 4)):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    h2d = d2h = 0
    n = args.steps
    for i in range(min(args.warmup, 4), min(args.warmup, 4) + n):
        a, b = step(i); h2d += a; d2h += b
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3) / n     # host work between steps is part of end-to-end time
    return {"value": 1000.0 / ms, "unit": "tok/s", "h2d_bytes_per_step": h2d // n, "d2h_bytes_per_step": d2h // n, "ms_per_step": ms,
            "api": "b200_executor_compute (C-ABI, include/b200_graph.h) with pinned host inputs, logits D2H, host argmax"}


def run_b200_pipeline(args, G, M, ops, rank, world, local):
    """N > 1: the reference's LLAMA_SPLIT_MODE_LAYER (llama-model.cpp:1917-1958): rank r owns a contiguous range of
    layers and their KV; the hidden state [n_embd] f32 is handed to rank r+1 with one NCCL send/recv per boundary
    (the reference: cudaMemcpyPeerAsync, ggml-cuda.cu:2556-2577), the sampled token goes back to rank 0.
    `world` independent sequences are kept in flight, one per pipeline slot, so that every GPU streams its slice of
    the weights once per tick; `value` = tokens emitted by the last stage per second (all sequences)."""
    import torch
    import torch.distributed as dist
    full = M.CONFIGS[args.model]
    L = args.layers or full["n_layer"]
    lo, hi = rank * L // world, (rank + 1) * L // world
    first, last = rank == 0, rank == world - 1
    model = M.SyntheticLlama(args.model, args.ftype, n_ctx=args.ctx, kv_type=G.F16, n_layer=L, layer_range=(lo, hi), first=first, last=last, n_seq=world)
    c = model.c
    V, E = c["n_vocab"], c["n_embd"]
    gen = torch.Generator(device="cuda"); gen.manual_seed(7 + rank)
    for ly in model.layers:
        for caches in (ly["k_caches"], ly["v_caches"]):
            for cache in caches:
                n = args.n_past * G.row_size(G.F16, c["n_head_kv"] * c["head_dim"]) // 2
                cache.view(torch.float16)[:n] = (torch.randn(n, device="cuda", generator=gen) * 0.5).half()
    ex = G.Executor(local)
    flags = (0 if args.no_graphs else G.EXEC_CUDA_GRAPHS) | (0 if args.no_fusion else G.EXEC_FUSION) | ((G.EXEC_MEGAKERNEL if (args.mega or args.mega_mmv) else 0) | (G.EXEC_MEGA_MMV if args.mega_mmv else 0) if not args.no_fusion else 0)
    n_kv_of = lambda pos: max(256, (pos + 1 + 255) // 256 * 256)  # noqa: E731
    e2e_steps = 0 if args.no_e2e else args.steps
    total_ticks = args.warmup + world - 1 + args.steps + e2e_steps   # pipeline fill + warm-up + timed region (+ the e2e leg)
    steps_per_seq = (total_ticks + world - 1) // world + 1
    assert n_kv_of(args.n_past + steps_per_seq) <= args.ctx
    stream = torch.cuda.current_stream()
    st = C.c_void_p(stream.cuda_stream)
    graphs = {}
    neg = torch.full((n_kv_of(args.n_past + steps_per_seq),), float("-inf"), device="cuda")

    def nodes_for(n_kv, seq):
        key = (n_kv, seq)
        if key not in graphs:
            graphs[key] = model.build(1, n_kv, seq=seq)
        return graphs[key]
    tok = torch.ones(1, dtype=torch.int32, device="cuda")
    seq_pos = [args.n_past] * world
    handoff_ev = []

    pending = [None]

    def isend(x, dst):
        if pending[0] is not None:
            pending[0].wait()
        pending[0] = dist.isend(x, dst=dst)

    import numpy as np
    n_kv_cap = n_kv_of(args.n_past + steps_per_seq)
    host_in = dict(pos=torch.zeros(1, dtype=torch.int32).pin_memory(), idx=torch.zeros(1, dtype=torch.int64).pin_memory(),
                   mask=torch.zeros(n_kv_cap, dtype=torch.float32).pin_memory(), tok=torch.ones(1, dtype=torch.int32).pin_memory())
    host_logits = torch.zeros(V, dtype=torch.float32).pin_memory() if last else None
    io_bytes = [0, 0]

    def tick(t, timed, e2e=False):
        seq = (t - rank) % world
        if t < rank:
            return                                                  # pipeline fill
        pos = seq_pos[seq]; seq_pos[seq] += 1
        n_kv = n_kv_of(pos)
        nodes, io = nodes_for(n_kv, seq)
        if first:
            if t >= world:                                          # the token this sequence sampled `world` ticks ago
                dist.recv(tok, src=world - 1)
            io["tokens"].copy_(tok)
        else:
            dist.recv(io["hidden_in"], src=rank - 1)
        if e2e:
            # the caller's side of the boundary: this tick's inputs come from pinned host memory ...
            host_in["pos"][0] = pos; host_in["idx"][0] = pos
            host_in["mask"][:n_kv].fill_(float("-inf")); host_in["mask"][:pos + 1] = 0
            io["pos"].copy_(host_in["pos"], non_blocking=True); io["kv_idx"].copy_(host_in["idx"], non_blocking=True)
            io["mask"][0].copy_(host_in["mask"][:n_kv], non_blocking=True)
            io_bytes[0] = 4 + 8 + 4 * n_kv + (4 if first else 0)
        else:
            io["pos"].fill_(pos); io["kv_idx"].fill_(pos)
            m = neg[:n_kv].clone(); m[:pos + 1] = 0
            io["mask"][0].copy_(m)
        ex.compute(nodes, flags, stream=st)
        if last:
            ops.check(ops.lib.b200_argmax_f32(ops.p(io["logits"]), ops.p(tok), V, 1, st))
            if e2e:
                # ... and the logits go back to the host, which picks the token (llama_get_logits + greedy sampling)
                host_logits.copy_(io["logits"][0], non_blocking=True)
                stream.synchronize()
                host_in["tok"][0] = int(np.argmax(host_logits.numpy()))
                io_bytes[1] = 4 * V
            if t + 1 < total_ticks:                                 # rank 0 stops receiving after the last tick
                isend(tok, 0)
        elif t + 1 < total_ticks:                                   # rank r+1 consumes it at tick t + 1
            isend(io["hidden_out"], rank + 1)

    sampler = ClockSampler(local); sampler.start()
    for t in range(args.warmup + world - 1):
        tick(t, False)
    torch.cuda.synchronize(); dist.barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.lib.b200_kernel_launches()
    e0.record(stream)
    for t in range(args.warmup + world - 1, args.warmup + world - 1 + args.steps):
        tick(t, True)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                     # device time, max over ranks
    launches = torch.tensor([ops.lib.b200_kernel_launches() - l0], device="cuda", dtype=torch.int64)
    dist.all_reduce(launches)
    clocks = sampler.stop()
    # ---- e2e leg: same ticks, host buffers on both sides of the boundary
    e2e = None
    if e2e_steps:
        t_base = args.warmup + world - 1 + args.steps
        dist.barrier(); torch.cuda.synchronize()
        e0.record(stream)
        for t in range(t_base, t_base + e2e_steps):
            tick(t, True, e2e=True)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        iob = torch.tensor(io_bytes, device="cuda", dtype=torch.int64)
        dist.barrier()
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX); dist.all_reduce(iob)
        e2e = {"value": 1000.0 * e2e_steps / float(ms2.item()), "unit": "tok/s", "h2d_bytes_per_step": int(iob[0].item()), "d2h_bytes_per_step": int(iob[1].item()),
               "ms_per_step": float(ms2.item()) / e2e_steps, "api": "b200_executor_compute per pipeline stage, pinned host inputs on every stage, logits D2H + host argmax on the last"}
    wb = torch.tensor([model.streamed_weight_bytes() + n_kv_of(args.n_past + args.steps // world) * model.kv_bytes_per_pos()], device="cuda", dtype=torch.int64)
    dist.all_reduce(wb)
    if rank == 0:
        ms_per_step = float(ms.item()) / args.steps
        hbm_peak, peak_src = peaks()
        wbytes = None
        line = {"metric": "decode tok/s Llama-3-8B Q4_K_M bs=1", "value": 1000.0 / ms_per_step, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "q4_K/q6_K x q8_K int8 dot (dp4a), f32 accumulate", "data": "synthetic",
                "config": {"workload": f"{args.model} {args.ftype} batch-1 decode per sequence, {world} sequences in flight (one per pipeline stage), -c {args.ctx}, n_past {args.n_past}, F16 KV",
                           "parallelism": f"layer split over {world} GPUs (--tensor-split {','.join(['1'] * world)}), NCCL send/recv hidden-state handoff",
                           "l2_policy": "inputs larger than L2"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.item()),
                "roofline": {"bound": "hbm", "kernel": "whole pipeline tick (all stages run concurrently, each streams its layer range)", "achieved": int(wb.item()) / (ms_per_step * 1e-3) / 1e9,
                             "peak": hbm_peak * world, "unit": "GB/s", "frac": int(wb.item()) / (ms_per_step * 1e-3) / 1e9 / (hbm_peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs"},
                "cpu_baseline": None,
                "note": "a step = one pipeline tick: every GPU streams its 1/N slice of the weights for one sequence; one token leaves the last stage per tick"}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    run_b200(args)


if __name__ == "__main__":
    main()
