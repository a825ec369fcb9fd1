"""Summarise an ncu --set full report of the demod kernel into profiles/*.json (+ a few lines of text).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/demod_pipe_ncu_summary.json [channels chunk_len]
"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
channels = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 48000
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
get = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name, scale=None):
    v, u = get[name]
    x = float(v.replace(",", ""))
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}
    return x * mult.get(u, 1)


s = {
    "report": rep, "kernel": get["Kernel Name"][0], "channels": channels, "chunk_len": chunk,
    "grid": get["launch__grid_size"][0], "block": get["launch__block_size"][0],
    "registers_per_thread": int(float(get["launch__registers_per_thread"][0])),
    "duration_s_under_ncu": num("gpu__time_duration.sum"),
    "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
    "algorithmic_bytes": 8.0 * channels * chunk,
    "dram_throughput_pct_of_peak": float(get["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]),
    "issue_active_pct": float(get["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
    "warps_active_pct": float(get["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
    "inst_executed": float(get["smsp__inst_executed.sum"][0]),
    "pipe_fma_pct": float(get["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]),
    "pipe_alu_pct": float(get["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"][0]),
    "shared_bank_conflicts": float(get["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0]),
}
s["traffic_over_algorithmic"] = (s["dram_bytes_read"] + s["dram_bytes_write"]) / s["algorithmic_bytes"]
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))
