"""debug: per-phase timeline of the persistent decode kernel on the bench model (GGML_B200_MK_TRACE=1)"""
import ctypes as C, os, sys, importlib
os.environ["GGML_B200_MK_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
from conftest import load_pkg
pkg = load_pkg(); ops = pkg.ops
G = importlib.import_module("llama_box_b200.graph"); M = importlib.import_module("llama_box_b200.model")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_past = int(sys.argv[2]) if len(sys.argv) > 2 else 512
model = M.SyntheticLlama("llama3-8b", "Q4_K_M", n_ctx=4096, kv_type=G.F16, n_layer=layers)
ex = G.Executor(0)
st = torch.cuda.Stream()
n_kv = (n_past + 256) // 256 * 256
nodes, io = model.build(1, n_kv)
io["tokens"].fill_(1); io["pos"].fill_(n_past); io["kv_idx"].fill_(n_past); io["out_ids"].fill_(0)
m = torch.full((n_kv,), float("-inf")); m[:n_past + 1] = 0; io["mask"][0].copy_(m)
flags = G.EXEC_FUSION | G.EXEC_MEGAKERNEL
with torch.cuda.stream(st):
    for _ in range(5):
        ex.compute(nodes, flags, stream=C.c_void_p(st.cuda_stream))
torch.cuda.synchronize()
ops.lib.b200_mk_trace_dump.restype = None
ops.lib.b200_mk_trace_dump(None)
print("mk launches", ex.mk_launches, "phases", ex.mk_phases, "kernels/step", ex.last_kernels)
