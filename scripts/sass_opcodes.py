"""Per-kernel histogram of the Blackwell-specific SASS opcodes in the built library (profiles/sass_opcodes.txt).

    python scripts/sass_opcodes.py [nerf_qa_b200/libnerfqa_b200.so] > profiles/sass_opcodes.txt

UTCHMMA / UTCHMMA.2CTA = tcgen05.mma (cta_group::1 / ::2), UTMALDG / UTMASTG = TMA loads / stores, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, HMMA = warp-level mma.sync (the banded-Toeplitz A-DISTS window kernel), LDSM = ldmatrix,
LDGSTS = cp.async, FFMA2 / FMUL2 / FADD2 = packed fp32.  Needs `cuobjdump` (CUDA toolkit); runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

WATCH = ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "FFMA2",
         "FMUL2", "FADD2", "MUFU")


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "nerf_qa_b200", "libnerfqa_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    if w == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                        continue
                    cur[w] += 1
                    break
    print(f"# {os.path.relpath(lib, root)}: SASS opcode counts per kernel (static, sm_100a)")
    print("kernel".ljust(64) + "".join(w.rjust(13) for w in ("instr",) + WATCH))
    tot = collections.Counter()
    for name, c in kernels.items():
        if not any(c[w] for w in WATCH):
            continue
        print(name[:63].ljust(64) + str(c["_total"]).rjust(13) + "".join(str(c[w]).rjust(13) for w in WATCH))
        tot.update(c)
    print("TOTAL".ljust(64) + str(tot["_total"]).rjust(13) + "".join(str(tot[w]).rjust(13) for w in WATCH))


if __name__ == "__main__":
    main()
