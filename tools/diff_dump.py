#!/usr/bin/env python3
"""Debug aid: run oracle/_ref/llama_drv twice (ggml-cpu and the B200 plug-in) with --dump (a checksum of every f32 graph node,
ggml eval callback) on the same GGUF / prompt and print the first nodes whose checksums deviate.
usage: tools/diff_dump.py model.gguf [--prompt-len N] [--gen G] [--kv f16|q8_0] [--tol 1e-5]"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "llama-box_b200", "libggml-b200.so")


def run(gguf, plugin, dump, a):
    cmd = [os.path.join(REF, "llama_drv"), "--model", gguf, "--ctx", "1024", "--prompt-len", str(a.prompt_len), "--gen", str(a.gen), "--fa", "--ctk", a.kv, "--ctv", a.kv, "--dump", dump]
    cmd += ["--plugin", PLUGIN, "--ngl", str(a.ngl), "--no-repack"] if plugin else ["--ngl", "0", "--threads", "16", "--no-repack"]
    if plugin:
        cmd += ["--threads", "16"]
    env = dict(os.environ, LD_LIBRARY_PATH=REF)
    env.pop("GGML_BACKEND_PATH", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.exit(r.stderr[-2000:])


def parse(path):
    out = []
    for ln in open(path):
        m = re.match(r"(\S+)\s+(\S+)\s+\[(.*?)\] sum=(\S+) abs=(\S+)", ln)
        if m:
            out.append((m.group(1), m.group(2), m.group(3), float(m.group(4)), float(m.group(5))))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("gguf"); ap.add_argument("--prompt-len", type=int, default=24); ap.add_argument("--gen", type=int, default=2)
    ap.add_argument("--kv", default="f16"); ap.add_argument("--tol", type=float, default=1e-5); ap.add_argument("--max", type=int, default=25); ap.add_argument("--ngl", type=int, default=99)
    a = ap.parse_args()
    run(a.gguf, False, "/tmp/dump_cpu.txt", a); run(a.gguf, True, "/tmp/dump_gpu.txt", a)
    c, g = parse("/tmp/dump_cpu.txt"), parse("/tmp/dump_gpu.txt")
    print(f"cpu nodes {len(c)}, gpu nodes {len(g)}")
    step, shown, first = 0, 0, True
    for i, (x, y) in enumerate(zip(c, g)):
        if x[0] != y[0]:
            print("node order differs at", i, x[0], y[0]); break
        rel = abs(x[4] - y[4]) / max(abs(x[4]), 1e-30)
        rels = abs(x[3] - y[3]) / max(abs(x[4]), 1e-30)
        dev = max(rel, rels)
        if dev > a.tol and shown < a.max:
            print(f"  step {step} node {i:4d} {x[0]:24s} {x[1]:14s} [{x[2]}] abs-sum rel {rel:.2e} sum rel {rels:.2e}")
            shown += 1
        if x[0] == "result_output":
            print(f"step {step}: result_output checksum deviation {dev:.2e}")
            step += 1; shown = 0


if __name__ == "__main__":
    main()
