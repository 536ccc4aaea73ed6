"""Summarise .ncu-rep captures (ncu --set full) into the text tables kept under profiles/.
usage: python scripts/ncu_summary.py rep1.ncu-rep [rep2.ncu-rep ...] > profiles/<name>.txt"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("%s  [%s]" % (r[hdr.index("Kernel Name")][:110], rep.split("/")[-1]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, r[i], units[i]))
        print()
