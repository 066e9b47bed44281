"""Micro-benchmark of ctrlsim_linear (the GEMM behind every nn.Linear) on decoder-sized shapes."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200 import lib as L
lib = L.load()
dev = torch.device("cuda:0")
shapes = [(256 * 2304, 768, 256), (256 * 2304, 256, 256), (256 * 2304, 1024, 256), (256 * 2304, 256, 1024)]
if len(sys.argv) > 1 and sys.argv[1] == "ffn2":  # the launch profiles/ncu_traffic.json describes: FFN2 of 90 focal groups
    shapes = [(90 * 2304, 256, 1024)]
elif len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
dbg = int(os.environ.get("GEMM_DEBUG", "0"))  # experiment bits of gemm_tc.cu (results are wrong when set)
model = None
if os.environ.get("GEMM_MODEL", "0") == "1":  # REGISTERED weights (lo tiles exist): the product path incl. the weight-stationary kernel
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    cfg = default_config(); model = DeviceModel(cfg, make_weights(cfg, seed=0), dev); lib = model.lib
    pre = "decoder.transformer_decoder.layers.0."
    names = {(768, 256): "self_attn.in_proj_weight", (256, 256): "self_attn.out_proj.weight", (1024, 256): "linear1.weight", (256, 1024): "linear2.weight"}
for M, N, K in shapes:
    A = torch.randn(M, K, device=dev)
    W = model.tensors[pre + names[(N, K)]] if model else torch.randn(N, K, device=dev) / math.sqrt(K)
    b = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, st)
    lib.ctrlsim_debug_gemm(dbg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    lib.ctrlsim_debug_gemm(0)
    ref = torch.nn.functional.linear(A[:512].double(), W.double(), b.double())
    err = (C[:512].double() - ref).abs().max().item()
    print(f"M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s  max|err|={err:.2e}  mode={os.environ.get('CTRLSIM_GEMM','default')} debug={dbg} registered={model is not None} ws={os.environ.get('CTRLSIM_GEMM_WS','1')}", flush=True)
