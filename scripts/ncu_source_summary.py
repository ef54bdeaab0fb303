"""Compact summary of an `ncu --page source --csv` export (scripts/r3_prof_kernel.sh): warp-stall reasons, opcode mix
(executed instructions and stall samples) and the hottest SASS instructions of the first kernel in the file.

    python scripts/ncu_source_summary.py gpurun_out/TAG_source.csv [top] > profiles/TAG_source_summary.txt
"""
import collections
import csv
import sys


def main(path, top=14):
    rows = list(csv.reader(open(path)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 10:
            cur["rows"].append(r)
    k = kern[0]
    h = k["hdr"]
    idx = {n: i for i, n in enumerate(h)}
    R = k["rows"]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    for r in R:
        for s in stalls:
            if r[idx[s]]:
                tot[s] += int(r[idx[s]])
    S = sum(tot.values())
    print("kernel:", k["name"][:110])
    print("warp-stall samples:", S)
    print("  " + "  ".join("%s %.1f%%" % (s[6:], 100 * v / S) for s, v in tot.most_common(10)))
    op, ops = collections.Counter(), collections.Counter()
    for r in R:
        o = r[idx["Source"]].split()
        name = (o[1] if o[0].startswith("@") else o[0]).split(".")[0]
        op[name] += int(r[idx["Instructions Executed"]] or 0)
        ops[name] += int(r[idx["# Samples"]] or 0)
    T = sum(op.values())
    print("warp instructions executed:", T)
    print("%-10s %8s %8s" % ("opcode", "instr %", "sample %"))
    for name, v in op.most_common(18):
        print("%-10s %7.1f%% %7.1f%%" % (name, 100 * v / T, 100 * ops[name] / S))
    print("hottest instructions (samples, executed, SASS, top stall reasons):")
    order = sorted(range(len(R)), key=lambda i: -int(R[i][idx["# Samples"]] or 0))
    for i in order[:top]:
        r = R[i]
        st = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print("%6s %9s  %-64s %s" % (r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:64], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
