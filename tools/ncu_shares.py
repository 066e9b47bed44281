"""Per-kernel device-time shares of the LAST full simulator step in an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...).   usage: python tools/ncu_shares.py X.csv [step_from_end]"""
import collections
import csv
import sys

path = sys.argv[1]
back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))


def short(n):
    n = n.split("(")[0].replace("void ", "").replace("ctrlsim::", "")
    return n


idx = [i for i, r in enumerate(rows) if short(r["Kernel Name"]).startswith("observe_kernel")]
sel = rows[idx[-back]:idx[-back + 1]] if back > 1 else rows[idx[-1]:]
# a simulator step ends with sim_step_kernel: drop what the bench launches after the last step (isolated micro-benchmarks)
ends = [i for i, r in enumerate(sel) if short(r["Kernel Name"]).startswith("sim_step_kernel")]
if ends:
    sel = sel[:ends[-1] + 1]
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in sel:
    v, u = float(r["Metric Value"]), r["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
    n = short(r["Kernel Name"])
    tot[n] += ms
    cnt[n] += 1
T = sum(tot.values())
print(f"{len(sel)} launches, {T:.3f} ms total\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for n, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"| {n} | {cnt[n]} | {v:.3f} | {100 * v / T:.1f}% |")
