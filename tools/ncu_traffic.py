"""Extract per-launch DRAM traffic of the captured GEMM launches from an ncu report (run here, after gpurun).
usage: python tools/ncu_traffic.py gpurun_out/gemm_fc1.ncu-rep M N K > profiles/r01_gemm_traffic.json"""
import csv, io, json, subprocess, sys
rep, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
units = rows[1]
def col(name): return h.index(name)
res = []
for r in rows[2:]:
    def val(name):
        i = col(name); v = float(r[i].replace(",", "")); u = units[i]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3}.get(u, 1)
    res.append({"kernel": r[col("Kernel Name")][:80], "us": val("gpu__time_duration.sum"),
                "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                "tensor_pct": float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])})
n = len(res)
alg = 2 * (M * K + N * K + M * N)     # bf16 A, B read once + bf16 D written once
print(json.dumps({"kernel": res[0]["kernel"], "shape": [M, N, K], "launches": n,
                  "dram_bytes_per_launch": sum(x["dram_read"] + x["dram_write"] for x in res) / n,
                  "dram_read_per_launch": sum(x["dram_read"] for x in res) / n,
                  "dram_write_per_launch": sum(x["dram_write"] for x in res) / n,
                  "algorithmic_bytes_per_launch": alg, "us_per_launch_under_ncu": sum(x["us"] for x in res) / n,
                  "tensor_pipe_active_pct": sum(x["tensor_pct"] for x in res) / n,
                  "note": "cold-cache, serialised ncu replay of tools/bench_gemm_parts.py (L2 flushed between launches)"}, indent=1))
