"""Timeline of CTA 0 of the linear-layer GEMM (debug mode 1): clock64 per k-slab."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ctrlsim_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0")
M, N, K = 256 * 2304, int(sys.argv[1]) if len(sys.argv) > 1 else 768, int(sys.argv[2]) if len(sys.argv) > 2 else 256
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / math.sqrt(K); b = torch.randn(N, device=dev)
if os.environ.get("TRACE_WLO", "1") == "1" and (N, K) in ((768, 256), (1024, 256), (256, 1024)):
    # use a REGISTERED weight so that the W_lo tiles come by TMA (the product path)
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    cfg = default_config(); model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
    p = "decoder.transformer_decoder.layers.0."
    W = model.tensors[p + {(768, 256): "self_attn.in_proj_weight", (1024, 256): "linear1.weight", (256, 1024): "linear2.weight"}[(N, K)]]
    print("W_lo by TMA (registered weight)")
Cm = torch.empty(M, N, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cm.data_ptr(), M, N, K, 0, st)
lib.ctrlsim_debug_gemm(int(os.environ.get('GEMM_DEBUG', '1')))
lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cm.data_ptr(), M, N, K, 0, st)
torch.cuda.synchronize()
tr = np.zeros((192, 4), dtype=np.int64)
lib.ctrlsim_debug_gemm_trace(tr.ctypes.data)
lib.ctrlsim_debug_gemm(0)
t0 = tr[0, 0]
print(f"M={M} N={N} K={K}: slab  tma_issue  landed  lo_done  mma_issued   (clk, relative)")
for i in range(8, 72):
    r = tr[i] - t0
    print(f"{i:4d} {int(r[0]):9d} {int(r[1]):9d} {int(r[2]):9d} {int(r[3]):9d}   land-issue {int(r[1]-r[0]):5d}  lo {int(r[2]-r[1]):4d}  mma-after-lo {int(r[3]-r[2]):5d}")
ep = tr[128:]
print("epilogue of CTA 0: tile  acc_ready  first_ld  first_chunk  released  (relative) | ld latency, chunk, total")
for i in range(2, 10):
    r = ep[i] - t0
    print(f"{i:4d} {int(r[0]):9d} {int(r[1]):9d} {int(r[2]):9d} {int(r[3]):9d} | {int(r[1]-r[0]):5d} {int(r[2]-r[1]):5d} {int(r[3]-r[0]):6d}")
d = np.diff(tr[16:120, 3])
print("slab period (clk): median", np.median(d), "mean", d.mean(), " ideal MMA time 813 (12 x 68) -> now 8 instr: 4 x (136 + 68) = 816")
