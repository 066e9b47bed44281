"""Micro-benchmark of the two HBM-bound attention kernels: attn_step (24 rows of the current step against the first pass'
K/V rows) and map_pool (the polyline pooling attention = BASELINE.json's "encoder-attn"), both at bench-like sizes with
inputs far larger than the 126 MB L2.  Prints achieved algorithmic GB/s against MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ctrlsim_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


A = 24
G = int(sys.argv[1]) if len(sys.argv) > 1 else 90
for ti, own in ((31, 1), (31, 0), (8, 1)):
    Lc = (ti + 1) * 72
    qkv = torch.randn(G, Lc, 768, device=dev)
    rows = torch.randn(G, A, 768, device=dev)
    O = torch.empty(G, A, 256, device=dev)
    ms = timeit(lambda: lib.ctrlsim_attn_step(qkv.data_ptr(), 768, 256, 512, Lc, rows.data_ptr(), O.data_ptr(), G, ti, own, st))
    nk = ti * 72 + A + (A if own else 0)
    by = G * (nk * 256 * 4 * 2 + A * 256 * 4 * 2)  # K + V rows of every visible key, Q in, O out
    print(f"attn_step G={G} ti={ti} own={own}: {ms*1e3:.1f} us, {by/ms/1e6:.0f} GB/s = {by/ms/1e6/peak:.3f} of measured HBM peak")
n_poly = G * 200
feats = torch.randn(n_poly, 100, 256, device=dev)
pv = (torch.rand(n_poly, 100, device=dev) > 0.1).to(torch.uint8)
ok = torch.ones(n_poly, dtype=torch.uint8, device=dev)
U = torch.randn(8, 256, device=dev) * 0.1
out = torch.empty(n_poly, 8, 256, device=dev)
ms = timeit(lambda: lib.ctrlsim_map_pool(feats.data_ptr(), pv.data_ptr(), ok.data_ptr(), U.data_ptr(), out.data_ptr(), n_poly, st))
by = n_poly * (100 * 256 * 4 + 100 + 8 * 256 * 4)
print(f"map_pool {n_poly} polylines: {ms*1e3:.1f} us, {by/ms/1e6:.0f} GB/s = {by/ms/1e6/peak:.3f} of measured HBM peak")
# the product path's polyline encoder, fused from raw points (map_encoder.cu)
from ctrlsim_b200.config import default_config
from ctrlsim_b200.weights import make_weights
from ctrlsim_b200.model import DeviceModel
cfg = default_config(); model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
pts = torch.randn(n_poly, 100, 3, device=dev)
pts[..., 2] = (torch.rand(n_poly, 100, device=dev) > 0.1).float()
pooled = torch.empty(n_poly, 8, 256, device=dev)
ms = timeit(lambda: model.lib.ctrlsim_map_encode_pool(model.handle, pts.data_ptr(), ok.data_ptr(), pooled.data_ptr(), n_poly, st))
by = n_poly * (100 * 3 * 4 + 1 + 8 * 256 * 4)
fl = n_poly * 100 * (2 * 3 * 256 + 2 * 256 * 8 + 2 * 8 * 256)
print(f"map_encode_pool {n_poly} polylines: {ms*1e3:.1f} us, {by/ms/1e6:.0f} GB/s algorithmic, {fl/ms/1e9:.1f} fp32 TFLOP/s")
