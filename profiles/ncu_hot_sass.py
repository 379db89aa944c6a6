"""Top SASS instructions by warp-stall samples for each kernel of an `ncu --page source --csv --print-source sass` dump."""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, hdr, rows = None, None, []


def flush():
    if not rows:
        return
    si = hdr.index("# Samples")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in rows)
    print("====", kern[:100], "total samples", tot)
    agg = {}
    for r in rows:
        for i in stalls:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("  stall mix:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for idx, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1][si] or 0))[:topn]:
        top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
        print(f"  {idx:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  {r[1].strip()[:70]:70s} {top}")


for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        flush()
        kern, hdr, rows = r[1], None, []
    elif r and r[0] == "Address":
        hdr = r
    elif hdr is not None and r:
        rows.append(r)
flush()
