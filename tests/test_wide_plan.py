"""CPU: the executor's handling of the wide path (MUL_MAT on Q4_1/Q5_1/Q2_K/Q3_K/IQ4_NL/IQ4_XS/MXFP4, MUL_MAT_ID, GET_ROWS
on quantised tables) through b200_executor_supports / b200_executor_plan — dry run, no device.  The switch GGML_B200_WIDE is read once per
process, so each case runs in a child interpreter."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import importlib, json, sys
sys.path.insert(0, %r)
from conftest import load_pkg
load_pkg()
G = importlib.import_module("llama_box_b200.graph")
import moe_graph
E, FF, NE, NU, NT, VOC = 4096, 2048, 8, 2, %d, 32000


class Fake:
    base = 0x10000000

    def buf(self, n):
        p = Fake.base; Fake.base += (n + 0xfffff) & ~0xfffff; return p

    def f32(self, ne):
        return G.T(self.buf(4 * ne[0] * ne[1] * (ne[2] if len(ne) > 2 else 1)), G.F32, ne)

    def i32(self, ne):
        return G.T(self.buf(4 * ne[0] * ne[1]), G.I32, ne)

    def named(self, name, t, ne):
        rows = 1
        for d in ne[1:]:
            rows *= d
        return G.T(self.buf(G.row_size(t, ne[0]) * rows), t, ne)


out = {"wide": int(G._lib.b200_executor_wide_enabled())}
nl, _ = moe_graph.build(G, Fake(), E, FF, NE, NU, NT, VOC, %d)
nodes = nl.build()
out["supports"] = [int(G._lib.b200_executor_supports(nodes[i])) for i in range(len(nodes))]
out["ops"] = [int(nodes[i].op) for i in range(len(nodes))]
out["plan"] = G.plan(nodes, G.EXEC_FUSION)
print(json.dumps(out))
"""


def run_child(wide, n_tok=1, wtype=3):
    env = dict(os.environ)
    env.pop("GGML_B200_WIDE", None)
    if wide:
        env["GGML_B200_WIDE"] = "1"
    r = subprocess.run([sys.executable, "-c", CHILD % (os.path.join(ROOT, "tests"), n_tok, wtype)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("wtype", [3, 7, 10, 11, 20, 23, 39])
def test_wide_nodes_are_supported_and_planned_when_switched_on(wtype):
    o = run_child(True, 1, wtype)
    assert o["wide"] == 1 and all(o["supports"]), o
    # GET_ROWS | RMS_NORM*w materialised (its consumer is not a tuned matvec, so it is NOT elided) | wide MUL_MAT | router: f32 MUL_MAT, SOFT_MAX, ARGSORT,
    # GET_ROWS (batched), SUM_ROWS, DIV | 2 x MUL_MAT_ID | SwiGLU | MUL_MAT_ID | weighted MUL | ADD of the two expert slices
    assert o["plan"] == 15, o


def test_wide_nodes_batch():
    o = run_child(True, 5, 23)
    assert all(o["supports"]) and o["plan"] == 15, o


def test_wide_nodes_are_refused_when_switched_off():
    o = run_child(False)
    G_OP_GET_ROWS, G_OP_MUL_MAT, G_OP_MUL_MAT_ID, G_OP_SOFT_MAX, G_OP_ARGSORT, G_OP_SUM_ROWS, G_OP_DIV = 9, 1, 11, 12, 13, 14, 15
    assert o["wide"] == 0
    for op, s in zip(o["ops"], o["supports"]):
        if op in (G_OP_GET_ROWS, G_OP_MUL_MAT, G_OP_MUL_MAT_ID, G_OP_SOFT_MAX, G_OP_ARGSORT, G_OP_SUM_ROWS, G_OP_DIV):
            assert s == 0, o              # ggml's scheduler keeps these nodes on its CPU backend, exactly as before this change
    assert o["plan"] < 0


NOFA_CHILD = r"""
import importlib, json, sys
sys.path.insert(0, %r)
from conftest import load_pkg
load_pkg()
G = importlib.import_module("llama_box_b200.graph")
import nofa_graph


class Fake:
    base = 0x10000000

    def buf(self, n):
        p = Fake.base; Fake.base += (n + 0xfffff) & ~0xfffff; return p

    def f32(self, ne):
        n = 4
        for d in ne:
            n *= d
        return G.T(self.buf(n), G.F32, ne)

    def named(self, name, t, ne):
        rows = 1
        for d in ne[1:]:
            rows *= d
        return G.T(self.buf(G.row_size(t, ne[0]) * rows), t, ne)


rp = [0, 128, 0, 0, 8192, G.f32_bits(5e5), G.f32_bits(1.0), G.f32_bits(0.0), G.f32_bits(1.0), G.f32_bits(32.0), G.f32_bits(1.0)]
nl, _ = nofa_graph.build(G, Fake(), 4096, 32, 8, 128, %d, 4096, 544, rp)
nodes = nl.build()
names = ["NONE", "MUL_MAT", "RMS_NORM", "MUL", "ADD", "ROPE", "SET_ROWS", "FA", "GLU", "GET_ROWS", "CPY", "MUL_MAT_ID", "SOFT_MAX", "ARGSORT", "SUM_ROWS", "DIV", "CONT"]
print(json.dumps({"wide": int(G._lib.b200_executor_wide_enabled()), "supports": [[names[nodes[i].op], int(G._lib.b200_executor_supports(nodes[i]))] for i in range(len(nodes))],
                  "plan": G.plan(nodes, G.EXEC_FUSION)}))
"""


@pytest.mark.parametrize("n_tok", [1, 5])
def test_attention_without_fa_is_supported_when_switched_on(n_tok):
    import json
    env = dict(os.environ, GGML_B200_WIDE="1")
    r = subprocess.run([sys.executable, "-c", NOFA_CHILD % (os.path.join(ROOT, "tests"), n_tok)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    o = json.loads(r.stdout.strip().splitlines()[-1])
    assert o["wide"] == 1 and all(s for _, s in o["supports"]), o
    # QKV in one launch | ROPE(q) | ROPE(k) | SET_ROWS(k) | element scatter of V | KQ | SOFT_MAX | KQV | CONT | wo
    assert 9 <= o["plan"] <= 12, o


def test_attention_without_fa_is_refused_when_switched_off():
    import json
    env = dict(os.environ)
    env.pop("GGML_B200_WIDE", None)
    r = subprocess.run([sys.executable, "-c", NOFA_CHILD % (os.path.join(ROOT, "tests"), 1)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    o = json.loads(r.stdout.strip().splitlines()[-1])
    refused = {name for name, s in o["supports"] if not s}
    assert {"SOFT_MAX", "CONT"} <= refused and o["plan"] < 0, o
