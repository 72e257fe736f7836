"""Dev tool: turn an .ncu-rep (ncu --set full) into the per-kernel summary committed under profiles/.
    python tools/ncu_summary.py gpurun_out/prof_all_r1c.ncu-rep profiles/r1_ncu_full_C2.md"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time us"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    ix = {c: h.index(c) for c, _ in COLS if c in h}
    kn = h.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none: {rep.split('/')[-1]}", "",
             "(per-launch, cold-cache, serialised replays: use the SHARES, not the absolutes)", "",
             "| kernel | " + " | ".join(n for c, n in COLS if c in ix) + " |",
             "|---|" + "---|" * len(ix)]
    tot = 0.0
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").split("::")[-1]
        vals = []
        for c, _ in COLS:
            if c not in ix:
                continue
            v = r[ix[c]]
            u = units[ix[c]]
            try:
                f = float(v.replace(",", ""))
                if c.startswith("dram__bytes"):
                    f = f * {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
                if c == "gpu__time_duration.sum":
                    f = f * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
                    tot += f
                v = f"{f:.2f}" if f < 1000 else f"{f:.0f}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(f"| {name} | " + " | ".join(vals) + " |")
    lines += ["", f"total of listed launches: {tot:.1f} us"]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
