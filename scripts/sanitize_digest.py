"""Summaries of the compute-sanitizer logs scripts/sanitize.sh leaves in gpurun_out/ (san_<tool>_{smoke,tests}_<tag>.log / .out).
usage: python scripts/sanitize_digest.py <tag> [key]   -> prints a JSON object; with `key`, also stores it under that key in
profiles/r02_sanitizer.json (the record of every sanitizer pass of the round)."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = {"command": f"gpurun -- bash scripts/sanitize.sh {tag}", "tools": {}}
for tool in ("memcheck", "racecheck", "synccheck", "initcheck"):
    for part in ("smoke", "tests"):
        log = os.path.join(ROOT, "gpurun_out", f"san_{tool}_{part}_{tag}.log")
        if not os.path.exists(log):
            continue
        lines = open(log, errors="replace").read().splitlines()
        summary = [re.sub(r"^=+\s*", "", l) for l in lines if "SUMMARY" in l]
        prog = open(log[:-4] + ".out", errors="replace").read().strip().splitlines()
        out["tools"].setdefault(tool, {})[part] = {"sanitizer_summary": summary, "program_tail": prog[-1] if prog else ""}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    path = os.path.join(ROOT, "profiles", "r02_sanitizer.json")
    rec = json.load(open(path))
    rec[sys.argv[2]] = out
    json.dump(rec, open(path, "w"), indent=1)
