"""Timeline of one attention CTA (debug mode 4): clock64 deltas per key tile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ctrlsim_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0")
G, n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 32
Lc = n_t * 72
qkv = torch.randn(G, Lc, 768, device=dev)
O = torch.empty(G, Lc, 256, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, st)
lib.ctrlsim_debug_attn(4)
lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, st)
torch.cuda.synchronize()
tr = np.zeros((64, 8), dtype=np.int64)
lib.ctrlsim_debug_attn_trace(tr.ctypes.data)
lib.ctrlsim_debug_attn(0)
t0 = tr[0, 7]
names = ["s_full", "sweep1", "p_ready", "o_full", "o_acc", "prod_done", "mma_p_rdy", "sm_top"]
print("tile  " + "  ".join(f"{n:>9s}" for n in ["sm_top", "s_full", "sweep1", "p_ready", "mma_p_rdy", "o_full", "o_acc", "prod_done"]))
for j in range(36):
    r = tr[j] - t0
    print(f"{j:4d}  " + "  ".join(f"{int(r[k]):9d}" for k in [7, 0, 1, 2, 6, 3, 4, 5]))
d = np.diff(tr[:36, 7])
print("per-tile period (clk): median", np.median(d), "mean", d.mean())
