"""Prints the handful of ncu --set full metrics DESIGN.md / profiles/ quote, for every kernel in a .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {rep}: {d.get('Kernel Name')}")
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
        try:   # achieved DRAM bandwidth of the launch = (bytes read + written) / duration
            SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
            val = lambda k: float(d[k].replace(",", "")) * SC[units[hdr.index(k)]]
            gbs = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / val("gpu__time_duration.sum") / 1e9
            print(f"  achieved DRAM bandwidth = {gbs:.0f} GB/s")
        except (KeyError, ValueError, ZeroDivisionError):
            pass
