import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0")
def run(q, kv, pad, mode):
    lib.ctrlsim_debug_attn(mode)
    G, Lq, Lk = q.shape[0], q.shape[1], kv.shape[1]
    O = torch.full((G, Lq, 256), -7.0, device=dev)
    p8 = pad.to(torch.uint8).contiguous()
    rc = lib.ctrlsim_attn_padded(q.data_ptr(), 256, kv.data_ptr(), kv.data_ptr() + 256 * 4, 512, p8.data_ptr(), O.data_ptr(), G, Lq, Lk, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); assert rc == 0
    lib.ctrlsim_debug_attn(0)
    return O
torch.manual_seed(0)
G, Lq, Lk = 1, 128, 64
pad = torch.zeros(G, Lk, dtype=torch.bool, device=dev)
q = torch.randn(G, Lq, 256, device=dev); kv = torch.randn(G, Lk, 512, device=dev)
S = (q[0, :, :32] @ kv[0, :, :32].T)  # head 0 raw scores [128, 64]
O1 = run(q, kv, pad, 1)
print("mode1 S dump err (head0, keys 0..31):", (O1[0, :, :32] - S[:, :32]).abs().max().item(), O1[0, 0, :4].tolist(), S[0, :4].tolist())
q0 = torch.zeros_like(q); kv1 = torch.ones_like(kv)
O3 = run(q0, kv1, pad, 3); print("mode3 P_hi readback (expect 1.0):", O3[0, 0, :4].tolist(), O3[0, 77, 28:32].tolist())
O2 = run(q0, kv1, pad, 2); print("mode2 raw O_main (expect 64):", O2[0, 0, :4].tolist(), O2[0, 77, 28:32].tolist())
kv2 = kv1.clone(); kv2[..., 256:] = torch.arange(256, device=dev, dtype=torch.float32)[None, None, :]
O2 = run(q0, kv2, pad, 2); print("mode2 raw O_main V=dim (expect 64*dim):", O2[0, 0, :6].tolist())
kv3 = kv1.clone(); kv3[..., 256:] = torch.arange(Lk, device=dev, dtype=torch.float32)[None, :, None]
O2 = run(q0, kv3, pad, 2); print("mode2 raw O_main V=key (expect 2016):", O2[0, 0, :6].tolist())

import math
def ref(q, kv, pad):
    G, Lq, Lk = q.shape[0], q.shape[1], kv.shape[1]
    qh = q.view(G, Lq, 8, 32).transpose(1, 2); kh = kv[..., :256].reshape(G, Lk, 8, 32).transpose(1, 2); vh = kv[..., 256:].reshape(G, Lk, 8, 32).transpose(1, 2)
    s_ = (qh / math.sqrt(32)) @ kh.transpose(-1, -2); s_ = s_.masked_fill(pad[:, None, None, :], float("-inf"))
    return (torch.softmax(s_, -1) @ vh).transpose(1, 2).reshape(G, Lq, 256)
for Lk2 in (64, 224):
    q = torch.randn(2, 150, 256, device=dev); kv = torch.randn(2, Lk2, 512, device=dev); pad2 = torch.rand(2, Lk2, device=dev) < 0.3; pad2[:, 0] = False
    O = run(q, kv, pad2, 0); print(Lk2, "random full: max err", (O - ref(q, kv, pad2)).abs().max().item())
