"""Per-kernel summary tables (markdown) from .ncu-rep files: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = set()
    print(f"## {path.split('/')[-1]}\n")
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        key = (name, r[ix["Grid Size"]] if "Grid Size" in ix else "")
        if key in seen:
            continue
        seen.add(key)
        print(f"kernel: `{name}`  grid {key[1]}\n\n| metric | value | unit |\n|---|---|---|")
        for m in WANT:
            if m in ix:
                print(f"| {m} | {r[ix[m]]} | {units[ix[m]]} |")
        print()
