import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0")
G, n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 32
Lc = n_t * 72
qkv = torch.randn(G, Lc, 768, device=dev)
O = torch.empty(G, Lc, 256, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
vis = sum(72 * (72 * tw + 24) + 72 for tw in range(n_t))
print(f"causal G={G}: {ms:.3f} ms, useful {G*8*vis*128/ms/1e9:.1f} TFLOP/s mode={os.environ.get('CTRLSIM_ATTN','tc')}")
