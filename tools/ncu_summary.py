"""Summary JSON of one `ncu --set full` capture (the metrics profiles/*_ncu_summary.json hold): runs here, no GPU.
  python tools/ncu_summary.py gpurun_out/r45_k_sp_search.ncu-rep "what was captured" [units_in_launch] > profiles/....json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio")

rep, what = sys.argv[1], sys.argv[2]
units = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
head, unit, vals = rows[0], rows[1], rows[2]
out = {}
for h, u, v in zip(head, unit, vals):
    if h in KEEP or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        out[f"{h} [{u}]" if u else h] = v
out["kernel"] = vals[head.index("Kernel Name")] if "Kernel Name" in head else ""
res = {"what": what, "metrics": out}
if units:
    def num(key):
        for k, v in out.items():
            if k.startswith(key):
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(k[k.find("[") + 1:-1], 1.0) if "[" in k else 1.0
                return float(v.replace(",", "")) * scale
        return 0.0
    res["metrics"]["_derived"] = {"units_in_launch": units,
                                  "warp_instructions_per_unit": num("smsp__inst_executed.sum") / units,
                                  "dram_bytes_per_unit": (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / units}
print(json.dumps(res, indent=1))
