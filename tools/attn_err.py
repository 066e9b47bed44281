"""Max |error| of the attention kernels against an fp64 torch reference (accuracy tracking across kernel versions)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200 import lib as L
from oracle.model_port import causal_mask_rule
lib = L.load(); dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for n_t, sc in [(32, 1.0), (32, 3.0), (7, 1.0)]:
    G, Lc = 2, n_t * 72
    g = torch.Generator(device="cpu").manual_seed(n_t)
    qkv = (torch.randn(G, Lc, 768, generator=g) * sc).to(dev)
    O = torch.empty(G, Lc, 256, device=dev)
    assert lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, st) == 0
    allowed = causal_mask_rule(24, n_t, 3).to(dev)
    d = qkv.double()
    qh = d[..., :256].reshape(G, Lc, 8, 32).transpose(1, 2)
    kh = d[..., 256:512].reshape(G, Lc, 8, 32).transpose(1, 2)
    vh = d[..., 512:].reshape(G, Lc, 8, 32).transpose(1, 2)
    s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(~allowed, float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, Lc, 256)
    ref32 = (torch.softmax(s.float(), -1) @ vh.float()).transpose(1, 2).reshape(G, Lc, 256)
    print(f"causal n_t={n_t} scale={sc}: max|err| kernel {(O.double()-ref).abs().max().item():.3e}  (torch fp32 softmax/matmul {(ref32.double()-ref).abs().max().item():.3e})", flush=True)
G, Lq, Lk = 3, 2304, 224
g = torch.Generator(device="cpu").manual_seed(1)
q = torch.randn(G, Lq, 256, generator=g).to(dev)
kv = torch.randn(G, Lk, 512, generator=g).to(dev)
pad = (torch.rand(G, Lk, generator=g) < 0.3).to(dev); pad[:, 0] = False
pad[1, :130] = True; pad[1, 140] = False
O = torch.empty(G, Lq, 256, device=dev)
padu8 = pad.to(torch.uint8).contiguous()
assert lib.ctrlsim_attn_padded(q.data_ptr(), 256, kv.data_ptr(), kv.data_ptr() + 256 * 4, 512, padu8.data_ptr(), O.data_ptr(), G, Lq, Lk, st) == 0
qh = q.double().view(G, Lq, 8, 32).transpose(1, 2)
kh = kv.double()[..., :256].reshape(G, Lk, 8, 32).transpose(1, 2)
vh = kv.double()[..., 256:].reshape(G, Lk, 8, 32).transpose(1, 2)
s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(pad[:, None, None, :], float("-inf"))
ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, Lq, 256)
print(f"padded Lq={Lq} Lk={Lk} (group 1: first two key tiles fully padded): max|err| {(O.double()-ref).abs().max().item():.3e}", flush=True)
