"""ncu launch list (the CSV of `--metrics gpu__time_duration.sum --csv --log-file x.csv`) -> the per-kernel markdown table kept
under profiles/:

    python scripts/launch_list_md.py gpurun_out/launches.csv "title line" > profiles/rNN_ncu_launch_list_X.md
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    depth, out = 0, []
    for ch in name:                      # cut at the argument list: the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out)[:120]


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "kernel launch list")
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(rows)
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total, n = sum(tot.values()), sum(cnt.values())
    ours = sum(v for k, v in tot.items() if k.startswith("coma::"))
    print(f"# {title}\n")
    print(f"`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` (eager launches, one step between "
          f"cudaProfilerStart/Stop; per-launch times are cold-cache and serialised: compare SHARES).  {n} launches, {total:.2f} ms in "
          f"total, {ours:.2f} ms ({100 * ours / total:.0f} %) in this repo's kernels (`coma::*`), the rest ATen / cuBLAS glue.\n")
    print("| ms | share | launches | kernel |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| {v:.3f} | {100 * v / total:.1f} % | {cnt[k]} | `{k}` |")


if __name__ == "__main__":
    main()
