"""profiles/ncu_traffic.json from `ncu --set full` captures: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
per launch of the two kernels bench.py reports a roofline for, next to the algorithmic work of the captured launch, so
that bench.py can scale the figure to the launches it timed (the captures run 8 scenes = 90 focal groups per launch,
the bench 256 groups per chunk).

    python tools/ncu_traffic.py gpurun_out/r02v_gemm_ffn2.ncu-rep gpurun_out/r02a_map_pool.ncu-rep
(the GEMM capture: `ncu --set full -k regex:gemm_tc_ta -s 2 -c 1 ... python tools/gemm_bench.py ffn2`, which launches
FFN2 at M = 90 x 2304 with registered weights)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v, u = float(r[ix[name]]), units[ix[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return rows[2:], ix, val


def main(paths):
    res = {"source": [os.path.basename(p) for p in paths],
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of one captured launch (ncu --set full, 8 scenes = 90 "
                   "focal groups); bench.py scales bytes_per_unit to the launches it timed"}
    for p in paths:
        rows, ix, val = rows_of(p)
        best = None
        for r in rows:
            name = r[ix["Kernel Name"]]
            by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            if best is None or by > best[1]:
                best = (name, by, r)
        name, by, r = best
        key = "gemm_tc_ta_kernel" if "gemm_tc" in name else "map_pool_kernel" if "map_pool" in name else name.split("(")[0]
        ent = {"kernel": name.split("(")[0], "dram_bytes_per_launch": by, "dram_read": val(r, "dram__bytes_read.sum"),
               "dram_write": val(r, "dram__bytes_write.sum"), "duration_us": float(r[ix["gpu__time_duration.sum"]])}
        if key == "gemm_tc_ta_kernel":
            # the launch with the most DRAM traffic is FFN2 (reads the 1024-wide hidden): M = 90 groups x 2304 rows,
            # N = 256, K = 1024; FFN1 (template <1, 1>, ReLU) moves the same bytes the other way round
            M, N, K = (90 * 2304, 1024, 256) if "kernel<1" in name.replace(" ", "") else (90 * 2304, 256, 1024)
            ent.update(shape=[M, N, K], algorithmic_flops=2.0 * M * N * K, algorithmic_bytes=4.0 * (M * K + N * K + M * N),
                       bytes_per_unit=by / (2.0 * M * N * K), unit="flop")
        elif key == "map_pool_kernel":
            n_poly = 90 * 200
            alg = n_poly * (100 * 256 * 4 + 100 + 8 * 256 * 4)
            ent.update(polylines=n_poly, algorithmic_bytes=alg, bytes_per_unit=by / alg, unit="algorithmic byte")
        res[key] = ent
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
