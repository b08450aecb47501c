"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed): one line per profiled launch with the metrics the
profiling recipe names.  usage: python bench/ncu_summary.py gpurun_out/prof_x.ncu-rep [out.csv]"""
import csv
import io
import subprocess
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    idx = [(w, h.index(w)) for w in WANT if w in h]
    out = io.StringIO()
    wr = csv.writer(out)
    wr.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in idx])
    for r in rows[2:]:
        wr.writerow([r[i][:90] for _, i in idx])
    text = out.getvalue()
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
