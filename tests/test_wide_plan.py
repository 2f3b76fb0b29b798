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
E, FF, NE, NU, NT, VOC = 4096, 2048, 8, 2, %d, 32000
base = [0x10000000]
def buf(n):
    p = base[0]; base[0] += (n + 0xfffff) & ~0xfffff; return p
f32 = lambda ne: G.T(buf(4 * ne[0] * ne[1] * (ne[2] if len(ne) > 2 else 1)), G.F32, ne)
def W(t, ne):
    rb = G.row_size(t, ne[0]); rows = ne[1] * (ne[2] if len(ne) > 2 else 1)
    return G.T(buf(rb * rows), t, ne)
out = {"wide": int(G._lib.b200_executor_wide_enabled())}
nl = G.NodeList()
tok = G.T(buf(64), G.I32, [NT])
emb = nl.add(G.OP_GET_ROWS, f32([E, NT]), [W(G.Q4_K, [E, VOC]), tok])                       # token embedding on the device
nrm = nl.add(G.OP_RMS_NORM, f32([E, NT]), [emb], [G.f32_bits(1e-5)])
cur = nl.add(G.OP_MUL, G.T(nrm.ptr, G.F32, [E, NT]), [nrm, f32([E, 1])])
proj = nl.add(G.OP_MUL_MAT, f32([E, NT]), [W(%d, [E, E]), cur])                              # a wide-only format consumes the norm
ids = G.T(buf(4 * NU * NT), G.I32, [NU, NT])
x3 = nl.view_op(G.T(proj.ptr, G.F32, [E, 1, NT], [4, 4 * E, 4 * E, 4 * E * NT]), proj)
up = nl.add(G.OP_MUL_MAT_ID, f32([FF, NU, NT]), [W(G.Q4_K, [E, FF, NE]), x3, ids])
gate = nl.add(G.OP_MUL_MAT_ID, f32([FF, NU, NT]), [W(G.Q4_K, [E, FF, NE]), x3, ids])
act = nl.add(G.OP_GLU_SWIGLU, f32([FF, NU, NT]), [gate, up], [2, 0])
down = nl.add(G.OP_MUL_MAT_ID, f32([E, NU, NT]), [W(G.Q6_K, [FF, E, NE]), act, ids])
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
    # GET_ROWS | RMS_NORM*w materialised (its consumer is not a tuned matvec, so it is NOT elided) | wide MUL_MAT | 2 x MUL_MAT_ID | SwiGLU | MUL_MAT_ID
    assert o["plan"] == 7, o


def test_wide_nodes_batch():
    o = run_child(True, 5, 23)
    assert all(o["supports"]) and o["plan"] == 7, o


def test_wide_nodes_are_refused_when_switched_off():
    o = run_child(False)
    G_OP_NONE, G_OP_GET_ROWS, G_OP_MUL_MAT, G_OP_MUL_MAT_ID = 0, 9, 1, 11
    assert o["wide"] == 0
    for op, s in zip(o["ops"], o["supports"]):
        if op in (G_OP_GET_ROWS, G_OP_MUL_MAT, G_OP_MUL_MAT_ID):
            assert s == 0, o              # ggml's scheduler keeps these nodes on its CPU backend, exactly as before this change
    assert o["plan"] < 0
