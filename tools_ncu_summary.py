"""Summarise an .ncu-rep (ncu --set full) per launch: key counters used in profiles/*.txt."""
import csv, io, subprocess, sys
rep = sys.argv[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
idx = {k: hdr.index(k) for k in keys if k in hdr}
name_i = hdr.index("Kernel Name")
units = rows[1]
stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled")]
for r in rows[2:]:
    print(r[name_i][:70], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for k, i in idx.items():
        print("   %-68s %16s %s" % (k, r[i], units[i]))
    st = sorted(((float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0, hdr[i]) for i in stall), reverse=True)[:5]
    print("   top stalls:", ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", ""), v) for v, h in st))
