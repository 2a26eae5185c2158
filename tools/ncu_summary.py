"""Summarise an `ncu --set full` report into the small CSV kept under profiles/ (one row per captured launch).
    python tools/ncu_summary.py gpurun_out/r02_prof_conv_tc.ncu-rep > profiles/r02_ncu_full_conv_tc.csv"""
import csv
import io
import subprocess
import sys

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
           "launch__cluster_size", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max.per_second",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    cols = [m for m in METRICS if m in idx]
    out = csv.writer(sys.stdout)
    out.writerow(["Kernel Name"] + cols)
    out.writerow([""] + [units[idx[c]] for c in cols])
    for r in rows[2:]:
        out.writerow([r[idx["Kernel Name"]][:100]] + [r[idx[c]] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1])
