import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0")
def run(q, kv, pad):
    G, Lq, Lk = q.shape[0], q.shape[1], kv.shape[1]
    O = torch.full((G, Lq, 256), -7.0, device=dev)
    p8 = pad.to(torch.uint8).contiguous()
    rc = lib.ctrlsim_attn_padded(q.data_ptr(), 256, kv.data_ptr(), kv.data_ptr() + 256 * 4, 512, p8.data_ptr(), O.data_ptr(), G, Lq, Lk, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); assert rc == 0
    return O
def ref(q, kv, pad):
    G, Lq, Lk = q.shape[0], q.shape[1], kv.shape[1]
    qh = q.view(G, Lq, 8, 32).transpose(1, 2); kh = kv[..., :256].reshape(G, Lk, 8, 32).transpose(1, 2); vh = kv[..., 256:].reshape(G, Lk, 8, 32).transpose(1, 2)
    s = (qh / math.sqrt(32)) @ kh.transpose(-1, -2); s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, Lq, 256)
torch.manual_seed(0)
for Lk in (64, 128, 224):
    G, Lq = 1, 128
    pad = torch.zeros(G, Lk, dtype=torch.bool, device=dev)
    q = torch.zeros(G, Lq, 256, device=dev); kv = torch.ones(G, Lk, 512, device=dev)
    O = run(q, kv, pad); print(Lk, "Q=0,V=1 ->", O[0, 0, :4].tolist(), O[0, 127, -4:].tolist(), "min/max", O.min().item(), O.max().item())
    kv2 = kv.clone(); kv2[..., 256:] = torch.arange(Lk, device=dev, dtype=torch.float32)[None, :, None]
    O = run(q, kv2, pad); print(Lk, "Q=0,V=key idx -> expect", (Lk - 1) / 2, O[0, 0, :4].tolist(), O[0, 100, 40:44].tolist())
    kv3 = kv.clone(); kv3[..., 256:] = torch.arange(256, device=dev, dtype=torch.float32)[None, None, :]
    O = run(q, kv3, pad); print(Lk, "Q=0,V=dim idx -> expect dim", O[0, 0, :6].tolist(), O[0, 5, 250:].tolist())
    q = torch.randn(G, Lq, 256, device=dev); kv = torch.randn(G, Lk, 512, device=dev)
    O = run(q, kv, pad); R = ref(q, kv, pad); print(Lk, "random: max err", (O - R).abs().max().item())
