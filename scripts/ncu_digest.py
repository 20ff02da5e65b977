"""Digest of an ncu report (raw page) into a small JSON for profiles/: duration, pipe / issue / stall metrics, DRAM bytes.
usage: python scripts/ncu_digest.py <report.ncu-rep> <out.json> [note]"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
res = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    k = {"kernel": d.get("Kernel Name"), "note": note, "metrics": {}}
    for h in WANT:
        if h in d and d[h] not in ("", "n/a"):
            k["metrics"][h] = {"value": d[h], "unit": u[h]}
    stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: round(float(v), 3) for h, v in d.items()
              if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a") and float(v) >= 0.05}
    k["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    res.append(k)
json.dump(res if len(res) > 1 else res[0], open(out, "w"), indent=1)
print(json.dumps(res, indent=1)[:1800])
