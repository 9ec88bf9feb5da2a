"""Digest of an ncu --set full report: one row per captured launch (duration, DRAM bytes, instructions, occupancy, issue rate, top stall
reasons per issued instruction). Usage: python tools/ncu_digest.py report.ncu-rep [more.ncu-rep ...] > profiles/rNN_xxx.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("smsp__inst_executed.sum", "warp inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]


def digest(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"\n### {path}\n")
    print("| kernel | grid x block | " + " | ".join(n for _, n in COLS) + " | top stalls (cycles per issued instruction) |")
    print("|---|---|" + "---|" * (len(COLS) + 1))
    for r in rows[2:]:
        g = {h: v for h, v in zip(hdr, r)}
        u = {h: v for h, v in zip(hdr, units)}
        cells = []
        for k, _ in COLS:
            v = g.get(k, "")
            try:
                f = float(v)
                cells.append(f"{f:,.0f}" if f >= 1000 else f"{f:.3g}")
            except ValueError:
                cells.append(v)
            if k.startswith("dram__bytes") and u.get(k):
                cells[-1] += " " + u[k]
        st = sorted(((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v)) for h, v in g.items()
                     if h.startswith("smsp__average_warps_issue_stalled") and v not in ("", "n/a")), key=lambda x: -x[1])
        print(f"| `{g['Kernel Name'][:60]}` | {g.get('launch__grid_size')} x {g.get('launch__block_size')} | " + " | ".join(cells) + " | " +
              ", ".join(f"{k} {v:.2f}" for k, v in st[:5]) + " |")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        digest(p)
