"""Multiplier-pipe probes (lab build: make lab; ZKP_B200_LIB=zk-paillier_b200/libzkp_b200_lab.so python scripts/imad_probes.py).
Each line: multiply-adds of the probed kind per second, and that rate as warp instructions per cycle per SM sub-partition
at the SM clock nvidia-smi reports under the probe."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_paillier_b200 as zk

NAMES = {0: "IMAD.WIDE.U32 independent (16 acc, 8 warps/SMSP)", 1: "IMAD.WIDE.U32.X carry-chained rows", 2: "IMAD 32-bit",
         7: "IMAD.HI.U32 independent", 8: "IMAD.WIDE.U32 zero addend (mul.wide)", 9: "IMAD.WIDE.U32 : IADD3 = 1:1",
         10: "IMAD.WIDE.U32 : LOP3 = 1:2", 11: "IMAD.WIDE.U32 32 accumulators", 12: "IMAD.WIDE.U32 loop-invariant multiplicands",
         13: "IMAD.WIDE.U32 : IMAD = 1:1 (rate counts the IMAD.WIDE only)", 14: "IMAD.WIDE.U32 independent, 1 warp/SMSP",
         15: "IMAD.WIDE.U32 independent, 2 warps/SMSP", 16: "IMAD.WIDE.U32 independent, 4 warps/SMSP",
         17: "DFMA independent (16 acc)", 18: "exact 52x52 product: 2 DFMA + DADD + 2 64-bit adds (rate counts products)",
         19: "DFMA : IMAD.WIDE.U32 = 1:1 (rate counts the DFMA only)"}
ctx = zk.native.Context(0)
sms = ctx.sm_count
out = {}
for v in sorted(NAMES):
    try:
        r = ctx.imad_peak(v)
    except Exception as e:
        out[NAMES[v]] = {"error": str(e)}
        continue
    mhz = float(subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip() or 0)
    per_smsp_clk = r / 32.0 / (sms * 4) / (1965e6)
    out[NAMES[v]] = {"mads_per_s": r, "warp_instr_per_clk_per_smsp_at_1965MHz": round(per_smsp_clk, 4), "cycles_per_warp_instr": round(1 / per_smsp_clk, 3), "sm_mhz_after": mhz}
    print(NAMES[v], json.dumps(out[NAMES[v]]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/imad_probes.json", "w"), indent=1)
