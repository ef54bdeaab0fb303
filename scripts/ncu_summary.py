"""Condense an ncu report (`ncu --set full`) into the per-launch table committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_summary.json
"""
import csv, io, json, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

COLS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_tensor_op_hmma.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg.per_second"]
SCALE = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) if c in hdr else None for c in COLS]
    out_units, out_rows = [], []
    for c, i in zip(COLS, idx):
        u = units[i] if i is not None else None
        out_units.append("Gbyte" if u in ("byte", "Kbyte", "Mbyte", "Gbyte") else "ms" if u in ("ns", "us", "ms", "s") else u)
    for d in data:
        r = []
        for c, i in zip(COLS, idx):
            if i is None:
                r.append(None)
                continue
            v, u = d[i], units[i]
            if c == "Kernel Name":
                v = v.replace("nqa::", "")[:48]
            elif u in SCALE and v not in ("", None):
                v = "%.6f" % (float(v.replace(",", "")) * SCALE[u])
            r.append(v)
        out_rows.append(r)
    with open(out, "w") as f:
        # csrc_sha16: fingerprint of the kernel sources this capture was taken from; bench.py quotes `traffic` from a summary
        # only while it matches the checkout
        from bench import csrc_sha16
        json.dump({"csrc_sha16": csrc_sha16(), "dram_unit": "Gbyte", "columns": COLS, "units": out_units, "rows": out_rows}, f,
                  indent=0)
    for r in out_rows:
        print(r[0][:44].ljust(44), r[1].ljust(14), "ms", r[2], "tensor%", r[3], "rd GB", r[4], "wr GB", r[5], "fma%", r[10])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
