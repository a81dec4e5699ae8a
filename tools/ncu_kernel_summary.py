"""Summarise an ncu report (run here, after gpurun): per kernel name -> launches, mean duration, DRAM bytes read / written, tensor-pipe
and issue utilisation, resident warps, registers.
usage: python tools/ncu_kernel_summary.py gpurun_out/x.ncu-rep [--json out.json --kernel <substring> --alg-bytes N --shape "..."]"""
import argparse, csv, io, json, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--json")
ap.add_argument("--kernel", default="")
ap.add_argument("--alg-bytes", type=float, default=0)
ap.add_argument("--shape", default="")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3, "s": 1e6}


def val(r, name, default=0.0):
    if name not in h:
        return default
    i = h.index(name)
    try:
        return float(r[i].replace(",", "")) * scale.get(units[i], 1)
    except ValueError:
        return default


agg = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    e = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, tensor=0.0, issue=0.0, warps=0.0, regs=0))
    e["n"] += 1
    e["us"] += val(r, "gpu__time_duration.sum")
    e["rd"] += val(r, "dram__bytes_read.sum")
    e["wr"] += val(r, "dram__bytes_write.sum")
    e["tensor"] += val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    e["issue"] += val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active", val(r, "smsp__issue_active.avg.pct"))
    e["warps"] += val(r, "sm__warps_active.avg.per_cycle_active")
    e["regs"] = int(val(r, "launch__registers_per_thread"))
print(f"{'kernel':70s} {'n':>3s} {'us':>9s} {'tensor%':>8s} {'issue%':>7s} {'warps/SM':>8s} {'dramRdMB':>9s} {'dramWrMB':>9s} {'regs':>5s}")
for name, e in agg.items():
    n = e["n"]
    print(f"{name[:70]:70s} {n:3d} {e['us'] / n:9.2f} {e['tensor'] / n:8.1f} {e['issue'] / n:7.1f} {e['warps'] / n:8.1f} {e['rd'] / n / 1e6:9.2f} {e['wr'] / n / 1e6:9.2f} {e['regs']:5d}")
if a.json:
    pick = [(k, e) for k, e in agg.items() if a.kernel in k]
    if not pick:
        sys.exit(f"no kernel matching {a.kernel!r}")
    k, e = pick[0]
    n = e["n"]
    json.dump({"kernel": k[:100], "shape": a.shape, "launches": n, "dram_bytes_per_launch": (e["rd"] + e["wr"]) / n,
               "dram_read_per_launch": e["rd"] / n, "dram_write_per_launch": e["wr"] / n, "algorithmic_bytes_per_launch": a.alg_bytes,
               "us_per_launch_under_ncu": e["us"] / n, "note": "ncu --set full --clock-control none, cold-cache serialised replay"},
              open(a.json, "w"), indent=1)
