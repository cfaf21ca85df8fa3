#!/usr/bin/env python
"""Instruction mix of the loops of a SASS listing (profiles/*.sass):
    python tools/sass_mix.py profiles/r2_mdpp_jit_rollout_fp64.sass
Prints every backward-branch loop (start, end, static instructions) and the
opcode histogram of each, so the per-chunk figures quoted in DESIGN.md /
profiles/README.md can be re-derived from the committed listings."""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
loops = []
for a, t in ins:
    m = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
print(f"{len(ins)} instructions, {len(loops)} loops")
for lo, hi in loops:
    body = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for a, t in ins if lo <= a <= hi]
    c = collections.Counter(op if not op.startswith("IMAD") else
                            ("IMAD.WIDE" if "WIDE" in op else "IMAD.MOV" if "MOV" in op else "IMAD")
                            for op in (b.split(".")[0] if not b.startswith("IMAD") else b for b in body))
    print(f"loop 0x{lo:x}-0x{hi:x}: {len(body)} instructions: "
          + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
