"""Correctness + speed of the linear-layer GEMM on REGISTERED weights (the product path: W_lo tiles fetched by TMA).
Shapes of the decoder; M values with odd / ragged 128-row tile counts."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200 import lib as L
from ctrlsim_b200.config import default_config
from ctrlsim_b200.weights import make_weights
from ctrlsim_b200.model import DeviceModel
lib = L.load(); dev = torch.device("cuda:0")
cfg = default_config(); model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
p = "decoder.transformer_decoder.layers.0."
cases = [("self_attn.in_proj_weight", "self_attn.in_proj_bias", 768, 256), ("linear1.weight", "linear1.bias", 1024, 256),
         ("linear2.weight", "linear2.bias", 256, 1024), ("self_attn.out_proj.weight", "self_attn.out_proj.bias", 256, 256)]
st = torch.cuda.current_stream().cuda_stream
for wn, bn, N, K in cases:
    W, b = model.tensors[p + wn], model.tensors[p + bn]
    for M in (256 * 2304, 1024 + 37, 128 * 9, 2304 * 3):
        A = torch.randn(M, K, device=dev)
        C = torch.full((M, N), float("nan"), device=dev)
        for relu in (0, 1):
            rc = lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, relu, st)
            assert rc == 0, lib.ctrlsim_last_error()
            torch.cuda.synchronize()
            idx = torch.cat([torch.arange(0, min(M, 300)), torch.arange(max(0, M - 300), M)]).to(dev)
            ref = torch.nn.functional.linear(A[idx].double(), W.double(), b.double())
            if relu: ref = torch.relu(ref)
            err = (C[idx].double() - ref).abs().max().item()
            assert not torch.isnan(C).any().item(), (wn, M, relu, "NaN left in C")
            assert err < 5e-5, (wn, M, relu, err)
        if M == 256 * 2304:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"N={N} K={K} M={M}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s  max|err|={err:.2e}  mode={os.environ.get('CTRLSIM_GEMM','default')}", flush=True)
print("all shapes ok")
