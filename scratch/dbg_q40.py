import sys, numpy as np, torch, importlib
sys.path.insert(0, 'tests')
from conftest import load_pkg
load_pkg()
G = importlib.import_module("llama_box_b200.graph"); M = importlib.import_module("llama_box_b200.model")
from oracle_model import OracleLlama, causal_mask, make_host_weights
from refutil import *
ops = load_pkg().ops
ft = sys.argv[1] if len(sys.argv) > 1 else "Q4_0"
cfg = dict(M.CONFIGS["test-small"]); cfg["n_layer"] = 1
mix, out_t = M.type_mix(ft, cfg["n_layer"])
hw, types = make_host_weights(cfg, mix, out_t)
model = M.SyntheticLlama(cfg, ft, n_ctx=512, kv_type=F16, host_weights=hw)
ex = G.Executor(0)
toks=[5, 17, 300, 4000, 9]; n=len(toks)
nodes, io = model.build(n, 256, want_all_logits=True)
io["tokens"].copy_(torch.tensor(toks, dtype=torch.int32)); io["pos"].copy_(torch.arange(0, n, dtype=torch.int32))
io["kv_idx"].copy_(torch.arange(0, n, dtype=torch.int64)); io["mask"].copy_(torch.from_numpy(causal_mask(n, 256, 0)))
ex.compute(nodes, 0); torch.cuda.synchronize()
orc = OracleLlama(cfg, hw, types, 512, F16)
tr = {}
lg = orc.forward(np.array(toks, np.int32), np.arange(0, n, dtype=np.int32), np.arange(0, n, dtype=np.int64), 256, causal_mask(n, 256, 0).astype(np.float16).view(np.uint16), trace=tr)
for key, ten in model.bufs.items():
    name = key[0].split("@")[0]
    if name in tr:
        got = ten.cpu().numpy().reshape(tr[name].shape)
        w = tr[name]
        print(f"{name:10s} maxabs {np.abs(got-w).max():.3e} scale {np.abs(w).max():.3e}")
got = io["logits"].cpu().numpy()
print("logits(last)", np.abs(got[-1]-lg).max(), np.abs(lg).max())
