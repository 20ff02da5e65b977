"""Opcode histogram of the innermost loops of one kernel's SASS (offline check before spending GPU time).
usage: python scripts/sass_loops.py <object.o> <mangled-function-substring>"""
import collections, re, subprocess, sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print("==", name, len(ins), "instructions")
    loops = []
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    for lo, hi in sorted(loops):
        body = [t for a, t in ins if lo <= a <= hi]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
        tag = "inner" if (lo, hi) in inner else "outer"
        wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
        print(f"-- loop {lo:#x}..{hi:#x} ({tag}) {len(body)} instr, {wide} IMAD.WIDE")
        if tag == "inner":
            print("   ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
