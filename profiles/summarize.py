"""Turn ncu captures in gpurun_out/ into small committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/r01_launches.csv profiles/r01_launches_bench.md
    python profiles/summarize.py kernel   gpurun_out/r01_gemv_q40_t1.ncu-rep profiles/r01_gemv_q40_t1.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        try:
            d.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in d.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum per kernel, --clock-control none\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | avg us | min us | max us | share of listed time |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {min(v) / 1e3:.2f} | {max(v) / 1e3:.2f} | {100 * sum(v) / total:.1f}% |\n")


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        name_i = hdr.index("Kernel Name")
        for r in data:
            f.write(f"## {r[name_i][:100]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            f.write("\nwarp stall reasons (warps stalled per issue-active cycle): " +
                    ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]) + "\n\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
