"""SM clock and board power WHILE one kernel class runs back to back for a few seconds (pynvml sampled every 20 ms from
a thread): tells a structural bound (clock stays near its maximum, power below the cap) from the 1 kW power cap
(power at the cap, SM clock pulled down; a kernel change that needs fewer joules per flop then shows up as clock, not
as a better pipe utilisation)."""
import sys, os, math, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from ctrlsim_b200.config import default_config
from ctrlsim_b200.weights import make_weights
from ctrlsim_b200.model import DeviceModel

dev = torch.device("cuda:0")
cfg = default_config()
model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
lib = model.lib
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], [False]


def poll():
    while not stop[0]:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.02)


def run(name, fn, flops, seconds=3.0, dbg=0):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        fn(st)
    torch.cuda.synchronize()
    lib.ctrlsim_debug_gemm(dbg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(st); torch.cuda.synchronize()
    e0.record(); fn(st); e1.record(); torch.cuda.synchronize()
    one = e0.elapsed_time(e1)
    n = max(10, int(seconds * 1000 / one))
    del samples[:]
    stop[0] = False
    th = threading.Thread(target=poll); th.start()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn(st)
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    stop[0] = True; th.join()
    lib.ctrlsim_debug_gemm(0)
    ms = e0.elapsed_time(e1) / n
    mid = [s for s in samples if t0 + 0.4 * (t1 - t0) <= s[0] <= t1]  # second half: after the clocks settled
    clk = statistics.median(s[1] for s in mid) if mid else float("nan")
    pw = statistics.median(s[2] for s in mid) if mid else float("nan")
    print(f"{name:46s} {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s  first launch {one:7.3f} ms  SM clock {clk:6.0f} MHz  power {pw:6.0f} W  ({len(mid)} samples)", flush=True)
    time.sleep(1.0)


M = 256 * 2304
pre = "decoder.transformer_decoder.layers.0."
for (N, K, nm) in ((1024, 256, "linear1.weight"), (256, 1024, "linear2.weight")):
    A = torch.randn(M, K, device=dev); W = model.tensors[pre + nm]; b = torch.randn(N, device=dev); C = torch.empty(M, N, device=dev)
    f = lambda st, A=A, W=W, b=b, C=C, N=N, K=K: lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, st)
    run(f"gemm {M}x{N}x{K} 3xTF32", f, 2.0 * M * N * K)
    run(f"gemm {M}x{N}x{K} no A_lo x W_hi MMAs (dbg 8)", f, 2.0 * M * N * K, dbg=8)
    run(f"gemm {M}x{N}x{K} no main MMAs (dbg 32)", f, 2.0 * M * N * K, dbg=32)
    run(f"gemm {M}x{N}x{K} no MMAs at all (dbg 40)", f, 2.0 * M * N * K, dbg=40)
    run(f"gemm {M}x{N}x{K} no stores (dbg 2)", f, 2.0 * M * N * K, dbg=2)
    del A, C
G, n_t = 64, 32
Lc = n_t * 72
qkv = torch.randn(G, Lc, 768, device=dev); O = torch.empty(G, Lc, 256, device=dev)
vis = sum(72 * (72 * tw + 24) + 72 for tw in range(n_t))
run("decoder self-attention (causal, G=64)", lambda st: lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, st), G * 8 * vis * 128.0)
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b2 = torch.empty_like(a)
run("torch copy 2 GiB (HBM reference)", lambda st: b2.copy_(a), 0.0)
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); y = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
run("cuBLAS bf16 8192^3 (reference)", lambda st: torch.matmul(x, y), 2.0 * 8192 ** 3)
torch.backends.cuda.matmul.allow_tf32 = True
xf = torch.randn(8192, 8192, device=dev); yf = torch.randn(8192, 8192, device=dev)
run("cuBLAS tf32 8192^3 (reference)", lambda st: torch.matmul(xf, yf), 2.0 * 8192 ** 3)
