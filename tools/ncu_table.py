"""Print a compact per-launch table from an .ncu-rep (raw page)."""
import csv, subprocess, sys
path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"), ("lts__t_bytes.sum", "L2MB"),
        ("l1tex__t_bytes.sum", "L1MB"), ("launch__grid_size", "grid"), ("smsp__inst_executed.avg", "inst/smsp")]
print(" ".join(f"{n:>10s}" for _, n in cols), " kernel")
for r in rows[2:]:
    vals = []
    for c, n in cols:
        if c in idx:
            v, u = r[idx[c]], units[idx[c]]
            try:
                f = float(v.replace(",", ""))
                if u == "Gbyte": f *= 1000
                if u == "Kbyte": f /= 1000
                if u == "byte": f /= 1e6
                if u in ("ms", "msecond"): f *= 1000
                if u in ("ns", "nsecond"): f /= 1000
                vals.append(f"{f:10.1f}")
            except ValueError:
                vals.append(f"{v[:10]:>10s}")
        else:
            vals.append(f"{'-':>10s}")
    print(" ".join(vals), " ", r[idx["Kernel Name"]][:60])
