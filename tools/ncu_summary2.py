"""Generic summary of one-kernel ncu --set full reports: python tools/ncu_summary2.py <rep> <out.json> [key=value ...]
Extra key=value pairs (e.g. algorithmic_bytes=..., algorithmic_flops=...) are copied into the JSON."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
extra = dict(a.split("=", 1) for a in sys.argv[3:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
get = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
MULT = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1,
        "Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6}
def num(name):
    if name not in get:
        cand = [h for h in get if h.endswith(name)]
        if not cand: return None
        name = cand[0]
    v, u = get[name]
    try: return float(v.replace(",", "")) * MULT.get(u, 1)
    except ValueError: return None
s = {"report": rep, "kernel": get["Kernel Name"][0], "grid": get["launch__grid_size"][0], "block": get["launch__block_size"][0],
     "registers_per_thread": num("launch__registers_per_thread"),
     "dynamic_smem_bytes": num("launch__shared_mem_per_block_dynamic"),
     "duration_s_under_ncu": num("gpu__time_duration.sum"),
     "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
     "dram_throughput_pct_of_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "l2_throughput_pct_of_peak": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     "l2_to_sm_bytes": num("l1tex__m_xbar2l1tex_read_bytes.sum"),
     "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "inst_executed": num("smsp__inst_executed.sum"),
     "pipe_fma_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
     "pipe_fp64_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
     "pipe_xu_pct": num("sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active"),
     "pipe_tensor_pct": num("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
     "tensor_hmma_cycles_active_avg": num("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"),
     "sm_cycles_active_avg": num("sm__cycles_active.avg"),
     "xu_inst_pct": num("sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed"),
     "shared_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
     "sm_clock_hz": num("sm__cycles_elapsed.avg.per_second")}
for k, v in extra.items():
    try: s[k] = float(v)
    except ValueError: s[k] = v
if "algorithmic_bytes" in s and s["dram_bytes_read"] is not None:
    s["traffic_over_algorithmic"] = (s["dram_bytes_read"] + s["dram_bytes_write"]) / s["algorithmic_bytes"]
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))
