"""Dev tool: text summary of an .ncu-rep (run here, no GPU needed).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [max_kernels] > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
]
rep = sys.argv[1]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full --clock-control none summary of {rep}")
for r in rows[2:2 + limit]:
    d = dict(zip(hdr, r))
    print(f"\n== {d.get('Kernel Name', '?')}  (ID {d.get('ID', '?')})")
    for k, label in KEYS:
        if k in d:
            print(f"  {label:32s} {d[k]:>16s} {units[hdr.index(k)]}")
    try:
        t = float(d["gpu__time_duration.sum"].replace(",", ""))
        u = units[hdr.index("gpu__time_duration.sum")]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-3)
        def gb(x, key):
            v = float(d[key].replace(",", "")); uu = units[hdr.index(key)]
            return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}.get(uu, 1.0)
        tr = gb(0, "dram__bytes_read.sum") + gb(0, "dram__bytes_write.sum")
        print(f"  {'dram traffic (read+write)':32s} {tr:16.3f} GB  -> {tr / t_s:8.1f} GB/s under the profiler")
    except Exception as e:  # noqa
        pass
