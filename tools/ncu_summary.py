"""Text summary of an .ncu-rep (ncu --set full): one block per profiled launch with the counters DESIGN.md quotes.
usage: ncu_summary.py report.ncu-rep > profiles/<name>_summary.txt"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
stalls = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
for r in rows[2:]:
    print("== %s" % r[hdr.index("Kernel Name")][:110])
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print("   %-26s %s %s" % (label, r[i], units[i]))
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stalls), reverse=True)[:6]
    print("   stalled warps per issue    " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in st))
