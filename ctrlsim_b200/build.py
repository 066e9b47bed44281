"""In-tree build of the CUDA extension: nvcc -> ctrlsim_b200/lib/libctrlsim_b200.so (sm_100a only).

    python -m ctrlsim_b200.build            # rebuild what is stale

No GPU is needed to build (nvcc cross-compiles).  sim.cu is compiled with -fmad=false (bit-faithful simulator
arithmetic); everything else with the default contraction.  -lineinfo keeps ncu's source page mapped to our code.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
LIB = os.path.join(OUT, "libctrlsim_b200.so")
LIB_WIDE = os.path.join(OUT, "libctrlsim_b200_wide.so")  # same sources with -DCTRLSIM_WIDE (64 agents / 256 polylines)
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "--extended-lambda", "-I", os.path.join(os.path.dirname(HERE), "include")]
UNITS = {"gemm.cu": [], "gemm_tc.cu": [], "attention.cu": [], "attention_mma.cu": [], "attention_tc.cu": [], "attention_step.cu": [], "map_encoder.cu": [], "tokens.cu": [], "sample.cu": [], "model.cu": [], "api.cu": [],
         "sim.cu": ["-fmad=false"]}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    tm = os.path.getmtime(target)
    return any(os.path.getmtime(d) > tm for d in deps)


def build(verbose: bool = False, force: bool = False, wide: bool = False) -> str:
    """Build one variant of the library (``wide``: the 64-agent / 256-polyline geometry, common.cuh)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT, exist_ok=True)
    lib, suffix, defs = (LIB_WIDE, "_wide.o", ["-DCTRLSIM_WIDE"]) if wide else (LIB, ".o", [])
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ctrlsim_b200.h"))
    objs, procs = [], []
    for unit, extra in UNITS.items():
        src = os.path.join(SRC, unit)
        obj = os.path.join(OUT, unit.replace(".cu", suffix))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *defs, *extra, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for unit, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {unit} ---\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(lib, objs):
        subprocess.check_call([nvcc, *ARCH, "-shared", "-o", lib, *objs, "-lcudart", "-ldl"])
    return lib


def build_all(verbose: bool = False, force: bool = False):
    return [build(verbose, force, wide=False), build(verbose, force, wide=True)]


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
