#!/usr/bin/env python3
"""Debug aid: per-row logits deviation of the plug-in vs ggml-cpu for one synthetic model, under several executor switches
(default / GGML_B200_DISABLE_FUSION / GGML_B200_DISABLE_GRAPHS / GGML_B200_DISABLE_PDL), to localise a parity break.
usage: tools/parity_probe.py --config llama3-8b --ftype Q8_0 --kv q8_0 [--layers 2] [--prompt-len 24] [--gen 6] [--verify 1]"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "llama-box_b200", "libggml-b200.so")


def run(gguf, a, plugin, env_extra, tag):
    out = f"/tmp/probe_{tag}.bin"
    cmd = [os.path.join(REF, "llama_drv"), "--model", gguf, "--ctx", "1024", "--prompt-len", str(a.prompt_len), "--gen", str(a.gen), "--fa", "--ctk", a.kv, "--ctv", a.kv,
           "--verify-batch", str(a.verify), "--logits-out", out, "--no-repack"]
    cmd += ["--plugin", PLUGIN, "--ngl", str(a.ngl)] if plugin else ["--ngl", "0", "--threads", "32"]
    env = dict(os.environ, LD_LIBRARY_PATH=REF, **env_extra)
    env.pop("GGML_BACKEND_PATH", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        print(tag, "FAILED:", r.stderr[-600:]); return None
    res = json.loads(r.stdout.strip().splitlines()[-1])
    return np.fromfile(out, np.float32).reshape(-1, res["n_vocab"]), res["tokens"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="llama3-8b"); ap.add_argument("--ftype", default="Q4_K_M"); ap.add_argument("--kv", default="f16")
    ap.add_argument("--layers", type=int, default=2); ap.add_argument("--prompt-len", type=int, default=24); ap.add_argument("--gen", type=int, default=6)
    ap.add_argument("--verify", type=int, default=1); ap.add_argument("--ngl", type=int, default=99)
    a = ap.parse_args()
    gguf = f"/dev/shm/probe_{a.config}_{a.ftype}_L{a.layers}.gguf"
    if not os.path.exists(gguf):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_gguf.py"), "--config", a.config, "--ftype", a.ftype, "--weights", "gauss", "--layers", str(a.layers), "--out", gguf],
                           capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=REF))
        assert r.returncode == 0, r.stderr[-1000:]
    cpu = run(gguf, a, False, {}, "cpu")
    print(f"== {a.config} {a.ftype} kv={a.kv} L={a.layers} prompt={a.prompt_len} verify={a.verify} ngl={a.ngl}")
    for tag, env in (("default", {}), ("nofusion", {"GGML_B200_DISABLE_FUSION": "1"}), ("nographs", {"GGML_B200_DISABLE_GRAPHS": "1"}), ("nopdl", {"GGML_B200_DISABLE_PDL": "1"})):
        g = run(gguf, a, True, env, tag)
        if g is None or cpu is None:
            continue
        rel = [float(np.abs(x - y).max() / np.abs(y).max()) for x, y in zip(g[0], cpu[0])]
        print(f"{tag:9s} tokens_equal={g[1] == cpu[1]}  rel per row: " + " ".join(f"{v:.1e}" for v in rel[:10]))


if __name__ == "__main__":
    main()
