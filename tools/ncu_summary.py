"""Turns an `ncu --set full` capture of tools/prof_kernels.py (one repetition = 14 launches) into profiles/*.json / *.md.
Usage (here, no GPU needed): python tools/ncu_summary.py gpurun_out/r01_full.ncu-rep profiles/r01_ncu_full_summary"""
import csv, io, json, subprocess, sys

LABELS = ["hv_pyr_fused_kernel(2 images 752x480)", "hv_pyr_fused_kernel(2 images 752x480)",
          "hv_lk_cta_kernel<31>(150 features, temporal, initial flow)", "hv_lk_cta_kernel<31>(150 features, stereo)",
          "ekf_predict_kernel(10 samples + normalisations)"]
for n, l in ((8, 34), (20, 55), (40, 90), (84, 160)):
    LABELS += [f"ekf_update_cluster2_kernel check n={n} l={l}", f"ekf_update_cluster2_kernel check+update n={n} l={l}"]
LABELS += ["ekf_update_cluster2_kernel symmetrise+augment"]
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
           "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}
kernels = []
# the capture is a window of the periodic launch sequence of prof_kernels.py: align the labels on the first pyramid launch
names = [r[col["Kernel Name"]] for r in data]
first_pyr = next((i for i, nm in enumerate(names) if "hv_pyr" in nm and "hv_pyr" in names[(i + 1) % len(names)]), 0)
for k, r in enumerate(data):
    e = {"launch": LABELS[(k - first_pyr) % len(LABELS)] if len(data) % len(LABELS) == 0 else r[col["Kernel Name"]], "kernel": r[col["Kernel Name"]]}
    for m in METRICS:
        if m in col and r[col[m]] not in ("", "n/a"):
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            if m.startswith("dram__bytes") or m == "lts__t_bytes.sum":
                v *= BYTES.get(u, 1.0); u = "byte"
            if m == "gpu__time_duration.sum":
                v *= TIME.get(u, 1.0); u = "us"
            if m == "launch__shared_mem_per_block_dynamic":
                v *= BYTES.get(u, 1.0) / 1e3; u = "Kbyte"
            e[m] = v
    e["dram_bytes"] = e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    kernels.append(e)
doc = {"source": "ncu --set full --clock-control none --import-source on -k regex:hv_|ekf_ (tools/prof_kernels.py, 2nd repetition); cold caches, "
                 "serialised launches; bytes in byte, time in us", "kernels": kernels}
json.dump(doc, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write("# ncu --set full summary (B200, tools/prof_kernels.py)\n\nCold-cache, serialised launches: compare shares, not absolutes. "
            "`dram` = dram__bytes_read.sum + dram__bytes_write.sum per launch.\n\n| launch | time us | grid | regs | dyn smem KB | dram bytes | warps active % | warp instr |\n|---|---|---|---|---|---|---|---|\n")
    for e in kernels:
        f.write(f"| {e['launch']} | {e.get('gpu__time_duration.sum', 0):.1f} | {int(e.get('launch__grid_size', 0))} | {int(e.get('launch__registers_per_thread', 0))} | "
                f"{e.get('launch__shared_mem_per_block_dynamic', 0):.1f} | {int(e['dram_bytes'])} | {e.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                f"{int(e.get('smsp__inst_executed.sum', 0))} |\n")
print("wrote", out + ".json", out + ".md", len(kernels), "launches")
