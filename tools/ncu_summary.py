"""Turn ncu artefacts from gpurun_out/ into the compact summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r1_launches.csv  > profiles/r1_launches_summary.txt
    python tools/ncu_summary.py full     gpurun_out/r1_prof.ncu-rep  > profiles/r1_prof_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__cluster_x",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def short(name):
    name = name.replace("void ", "").replace("gssd::", "")
    return name[:name.index("(")] if "(" in name else name


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        k = (short(r[4]), r[7], r[8])
        per.setdefault(k, []).append(float(r[14]) / 1e3)
    total = sum(sum(v) for v in per.values())
    print("%-64s %-14s %-14s %6s %10s %10s %7s" % ("kernel", "block", "grid", "n", "avg us", "total us", "share"))
    for (n, b, g), v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print("%-64s %-14s %-14s %6d %10.2f %10.1f %6.1f%%" % (n[:64], b, g, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))
    print("total device time of %d launches: %.1f us" % (len(rows), total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s  grid %s block %s" % (short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for m in FULL_METRICS:
            if m in hdr:
                print("   %-72s %16s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
