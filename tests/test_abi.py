"""CPU: the C-ABI libraries load and export every symbol their headers declare (no compute calls —
there is no GPU here), and the product refuses to run without CUDA instead of falling back."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+|ggml_backend_[a-z0-9_]+)\s*\(", src)))


def test_b200_ops_exports_every_declared_symbol(b200):
    names = [n for n in declared_symbols("b200_ops.h")]
    assert len(names) >= 30
    for n in names:
        assert hasattr(b200.lib, n), n
    assert set(names) == set(b200.SIGNATURES), set(names) ^ set(b200.SIGNATURES)
    assert b200.lib.b200_abi_version() == 1


def test_geometry_matches_ggml_blocks(b200):
    # ggml/src/ggml-common.h:170-175,219-224,295-344
    for t, (elems, size) in {2: (32, 18), 8: (32, 34), 12: (256, 144), 13: (256, 176), 14: (256, 210)}.items():
        assert b200.lib.b200_block_elems(t) == elems and b200.lib.b200_block_bytes(t) == size
        assert b200.lib.b200_row_bytes(t, 4096) == 4096 // elems * size
    assert b200.lib.b200_row_bytes(12, 100) == -1
    assert b200.lib.b200_act_kind_for(12) == 0 and b200.lib.b200_act_kind_for(8) == 1 and b200.lib.b200_act_kind_for(0) < 0
    assert b200.lib.b200_act_col_bytes(0, 4096) == 4096 + 64 + 512
    assert b200.lib.b200_act_col_bytes(1, 4096) == 4096 + 512 + 256


def test_no_cpu_fallback_without_cuda(b200):
    import torch
    if torch.cuda.is_available():
        return
    assert b200.lib.b200_device_count() == 0
    buf = (C.c_float * 1024)()
    st = b200.lib.b200_rms_norm(C.addressof(buf), None, C.addressof(buf), 1024, 1, 1024, 1024, 1e-5, None)
    assert st == -3 and b"CUDA" in b200.lib.b200_last_error()


def test_b200_graph_exports_every_declared_symbol(b200):
    """include/b200_graph.h (executor C-ABI): every declared entry point is exported and mirrored in graph.py"""
    import importlib
    G = importlib.import_module("llama_box_b200.graph")
    names = declared_symbols("b200_graph.h")
    assert len(names) >= 8
    for n in names:
        assert hasattr(b200.lib, n), n
    assert set(names) == set(G.GRAPH_SYMBOLS), set(names) ^ set(G.GRAPH_SYMBOLS)


def test_ggml_plugin_exports_backend_entry_points():
    """include/ggml_b200.h: the dynamic-backend entry points ggml's loader dlsym()s (ggml-backend-reg.cpp:243-274).
    The plug-in links against the reference's libggml-base, so it only exists where oracle/_ref was built."""
    so = os.path.join(ROOT, "llama-box_b200", "libggml-b200.so")
    if not os.path.exists(so):
        import pytest
        pytest.skip("libggml-b200.so not built (needs oracle/_ref)")
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    for n in declared_symbols("ggml_b200.h"):
        assert n in exported, n
    assert {"ggml_backend_init", "ggml_backend_score"} <= exported
