"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py profiles/r01a_launches.csv"""
import collections
import csv
import re
import sys

EPI = {"0": "EPI_STORE (merger fc1/fc2)", "1": "EPI_RESID (patch-embed, proj, down)", "2": "EPI_QKV", "3": "EPI_SWIGLU (gate|up)",
       "100": "EPI_RESID deep ring (down)", "4": "EPI_LOGPROB", "5": "EPI_STORE_MULTI", "6": "EPI_MROPE"}


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"]
        m = re.search(r"gemm_kernel<\(?(?:int\))?(\d+), *\(?(?:int\))?(\d+)", name)
        m2 = re.search(r"gemm2_kernel<\(?(?:int\))?(\d+)", name)
        if m2:
            key = f"gemm2_kernel<{EPI.get(m2.group(1), m2.group(1))}> (CTA pair)"
        elif m:
            key = f"gemm_kernel<BN={m.group(1)}, {EPI.get(m.group(2), m.group(2))}>"
        else:
            key = re.sub(r"\(.*", "", name.split("::")[-1])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / v[0]:.1f} | {v[1] / tot * 100:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
