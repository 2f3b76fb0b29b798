"""GPU: the wide path — MUL_MAT on Q4_1 / Q5_1 / Q2_K / Q3_K / IQ4_NL / IQ4_XS / MXFP4, MUL_MAT_ID (expert routing read on the device) and
GET_ROWS on quantised tables, and the KV cache type q4_0 (llama-box_b200/csrc/mmvq_ext.cu, fattn_ext.cu) — against the C oracle, through the C-ABI
(b200_mul_mat_vec_wide / b200_mul_mat_id / b200_get_rows_q), through the graph executor, and through the plug-in with the reference's own
test-backend-ops harness.

STATUS: these kernels were written after round 2's GPU budget had been spent.  Their per-format arithmetic (extfmt.cuh) is the same code
tests/test_extfmt_hostsim.py checks on the CPU against the reference, but the kernels themselves have NOT run on hardware yet, which is why
  * the path is off by default (GGML_B200_WIDE=1 switches it on),
  * every case runs in a child process (a faulting kernel must not poison the CUDA context of the rest of the GPU suite), and
  * the cases are marked xfail(strict=False): XPASS = verified on this box; XFAIL = a defect to fix next round, not a regression of the
    verified path.  The file sorts last on purpose."""
import json
import os
import re
import subprocess
import sys

import pytest

from refutil import REF_DIR, ROOT, have_ref

UNVERIFIED = "wide path: written after the round's GPU budget was spent, first hardware run is this one (see module docstring)"
pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason=UNVERIFIED)]
CHILD = os.path.join(ROOT, "tests", "wide_gpu_child.py")
CORE = [2, 6, 8, 12, 13, 14]
EXT = [3, 7, 10, 11, 20, 23, 39]


def child(*args, wide=False, timeout=600):
    env = dict(os.environ)
    env.pop("GGML_B200_WIDE", None)
    if wide:
        env["GGML_B200_WIDE"] = "1"
    r = subprocess.run([sys.executable, CHILD] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=timeout)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_numpy_model_of_the_library_layout_matches_the_repack_kernel():
    o = child("repack_models", "2,6,8,14")
    assert all(v["equal"] for v in o.values()), o


@pytest.mark.parametrize("group", ["3,7,20,39", "10,11,23", "2,6,8", "12,13,14"])
def test_wide_kernels_vs_oracle(group):
    """per format: b200_mul_mat_vec_wide (4 shapes x {random bit patterns, reference-quantised weights}, with and without bias + residual; rows that are
    not a multiple of 256 for the 32-element formats), b200_mul_mat_id (shared / per-expert activations, 1 and 5 tokens), b200_get_rows_q.  A few formats
    per child process (the interpreter + torch start-up would otherwise dominate the run time)."""
    bad = {}
    o = child("type_suites", group, timeout=1800)
    for t, cases in o.items():
        assert len(cases) >= 13, (t, len(cases))
        for name, r in cases.items():
            key = f"type {t}: {name}"
            if "error" in r:
                bad[key] = r
            elif name.startswith("mul_mat_id"):
                if r["err"] > 2e-5:
                    bad[key] = r
            elif name.startswith("mul_mat"):
                if r["plain"] > 2e-5 or r["bias_residual"] > 2e-5:
                    bad[key] = r
            elif not r["bit_exact"]:
                bad[key] = r
    assert not bad, bad


@pytest.mark.parametrize("shape", [(128, 32, 8, 1, 768), (128, 8, 2, 5, 4096), (64, 32, 4, 1, 1024), (64, 8, 8, 33, 256)])
def test_q4_0_kv_cache_set_rows_and_flash_attn_vs_oracle(shape):
    """KV cache type q4_0 (`-ctk q4_0 -ctv q4_0`): b200_set_rows_q4_0 bytes == ggml's from_float; b200_flash_attn_q4_0 within the Q8_0-cache tolerance"""
    o = child("kv_q4_0", *shape)
    assert o["set_rows_equal"], o
    assert o["attn_err"] <= 2e-5, o


def test_flash_attn_any_head_size_vs_oracle():
    """b200_flash_attn_any: head sizes 32 / 96 / 192 / 256 over F16, Q8_0 and Q4_0 caches (the tuned kernels carry 64 and 128)"""
    o = child("attn_any_suite")
    bad = {k: v for k, v in o.items() if "error" in v or v["err"] > 2e-5}
    assert len(o) == 6 and not bad, bad


@pytest.mark.parametrize("n_tok,n_past", [(1, 100), (5, 37)])
def test_attention_without_fa_block_vs_oracle(n_tok, n_past):
    """one attention block as libllama emits it WITHOUT -fa (tests/nofa_graph.py) through the graph executor: K rows and transposed-V elements stored
    exactly, output within the chained-op tolerance (the probabilities are rounded to f16 before KQV on both sides, like ggml-cpu does)"""
    o = child("nofa", n_tok, n_past, wide=True)
    assert all(o["supports"]), o
    assert o["v_store_exact"], o
    assert max(o["errs"]) <= 2e-3, o
    assert o["captures"] >= 1 and o["replays"] >= 1, o


def test_moe_router_glue_vs_oracle():
    """glue_ext.cu through the C-ABI: f32 router matmul, SOFT_MAX, ARGSORT (ties in the reference's order), batched GET_ROWS, SUM_ROWS, DIV, the
    broadcast MUL and the ADD over strided expert slices"""
    o = child("glue")
    assert o["mul_mat_f32"] <= 1e-5 and o["soft_max"] <= 1e-6 and o["unary"] <= 1e-6, o
    assert all(o[k] for k in ("argsort_exact", "get_rows_batched_exact", "sum_rows_exact", "div_exact", "mul_bcast_exact", "add_slices_exact")), o


@pytest.mark.parametrize("wtype,n_tok", [(3, 1), (11, 4), (23, 1)])
def test_executor_moe_block_vs_oracle(wtype, n_tok):
    """a whole Mixtral-style block (tests/moe_graph.py = the node list build_moe_ffn emits) through the graph executor: 15 launches, nothing leaves the device"""
    o = child("executor", wtype, n_tok, wide=True)
    assert all(o["supports"]), o
    assert max(o["errs"]) <= 2e-3, o                         # chained ops: a 1e-6 difference in an f32 intermediate may flip one int8 re-quantisation downstream
                                                             # (~1e-3 of the output scale); the per-op cases above hold the tight bounds
    assert o["captures"] >= 1 and o["replays"] >= 1 and o["kernels"] == 15, o


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("op,min_ok", [("MUL_MAT_ID", 10), ("GET_ROWS,SET_ROWS,ADD,MUL,DIV,SOFT_MAX,ARGSORT,SUM_ROWS,CONT", 40)])
def test_reference_backend_ops_harness_with_the_wide_path(op, min_ok):
    """the reference's own parity harness (tests/test-backend-ops.cpp) against the plug-in with GGML_B200_WIDE=1.  MUL_MAT / FLASH_ATTN_EXT are not re-run here
    (minutes of CPU reference time that tests/test_gpu_plugin.py already spends on them): the wide formats and the q4_0 cache reach the plug-in through the libllama cases below."""
    plugin = os.path.join(ROOT, "llama-box_b200", "libggml-b200.so")
    env = dict(os.environ, LD_LIBRARY_PATH=REF_DIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), GGML_BACKEND_PATH=plugin, GGML_B200_WIDE="1")
    r = subprocess.run([os.path.join(REF_DIR, "test-backend-ops"), "test", "-b", "B2000", "-o", op], capture_output=True, text=True, env=env, timeout=1800)
    plain = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout + r.stderr)
    ok = len(re.findall(r"\): OK", plain)); fail = len(re.findall(r"FAIL", plain))
    assert fail == 0 and r.returncode == 0, plain[-4000:]
    assert ok >= min_ok, (ok, plain[-2000:])


# ---- the product boundary: the unmodified libllama over the plug-in with GGML_B200_WIDE=1, against ggml-cpu on the same file and prompt ----------
@pytest.fixture(scope="module")
def product():
    import test_gpu_product as P
    for key in [k for k, path in P._models.items() if not os.path.exists(path)]:     # files that module's own clean-up already removed
        del P._models[key]
    yield P
    for path in list(P._models.values()):
        try:
            os.remove(path)
        except OSError:
            pass


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("config,ftype,kv,must_be_on_device", [
    ("test-moe", "Q4_K_M", "f16", "MUL_MAT_ID"),            # Mixtral-style experts: build_moe_ffn -> 3 x MUL_MAT_ID per layer
    ("test-small", "Q4_1", "f16", "MUL_MAT"), ("test-small", "Q5_1", "f16", "MUL_MAT"), ("test-small", "Q2_K", "f16", "MUL_MAT"),
    ("test-small", "Q3_K_M", "f16", "MUL_MAT"), ("test-small", "IQ4_NL", "f16", "MUL_MAT"), ("test-small", "IQ4_XS", "f16", "MUL_MAT"),
    ("test-small", "Q4_K_M", "q4_0", "FLASH_ATTN"),         # -ctk q4_0 -ctv q4_0
    ("test-small", "Q4_K_M", "f16", "SOFT_MAX"),            # WITHOUT -fa (llama-box's default): KQ / masked SOFT_MAX / KQV / CONT on the device
])
def test_libllama_over_the_wide_path(tmp_path, tmp_path_factory, product, config, ftype, kv, must_be_on_device):
    """llama_decode (prefill of 24 tokens, then greedy decode) on files whose matrices are in the wide path's formats / a mixture-of-experts
    FFN / a q4_0 KV cache.  The scheduler must place the named op on the B200 backend (GGML_SCHED_DEBUG=2 assignment dump — otherwise the test
    would pass through ggml's CPU fallback), and the run must be as close to ggml-cpu as ggml-cpu's own other builds are."""
    P = product
    gguf = P.model_file(tmp_path_factory, config, ftype, 2)
    fa = must_be_on_device != "SOFT_MAX"
    cpu = P.drv(gguf, str(tmp_path / "cpu"), False, kv=kv, gen=9, fa=fa)
    others = P.cpu_builds(tmp_path, tmp_path_factory, gguf, kv=kv, gen=9, fa=fa)
    gpu = P.drv(gguf, str(tmp_path / "gpu"), True, kv=kv, gen=9, fa=fa, extra_env=dict(GGML_B200_WIDE="1", GGML_SCHED_DEBUG="2", LLAMA_DRV_LOG_DEBUG="1"))
    placed = re.findall(r"node #\s*\d+ \(\s*" + must_be_on_device + r"\w*\):.*?\[\s*(\w+)", gpu["stderr"])
    assert placed and all(b.startswith("B200") for b in placed), (len(placed), sorted(set(placed)))
    P.assert_within_reference_self_consistency(gpu, cpu, others, f"wide path: {config} {ftype} kv={kv}")
