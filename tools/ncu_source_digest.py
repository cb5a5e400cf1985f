"""Digest of the SASS source page of one captured kernel (`ncu --set full --import-source on`):

    python tools/ncu_source_digest.py x.ncu-rep out.txt [--top 60]

Reads `ncu -i x.ncu-rep --page source --csv` (works without a GPU) and writes
  * warp-instructions executed per opcode (what the issue slots are spent on),
  * the instructions with the most warp-stall samples, each with its dominant stall reasons,
  * stall samples summed per opcode.
The full page is tens of MB; this keeps the part that explains a kernel's time in a few KB."""

import argparse
import csv
import io
import re
import subprocess
from collections import Counter, defaultdict

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("out")
ap.add_argument("--top", type=int, default=60)
args = ap.parse_args()

if args.report.endswith(".csv"):  # an already exported page
    raw = open(args.report).read()
else:
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
kernel = lines[0] if start > 0 else ""
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
col = {name: i for i, name in enumerate(hdr)}
stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]


def num(v: str) -> float:
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def opcode(src: str) -> str:
    s = re.sub(r"^@!?U?P\d+\s+", "", src.strip())
    return s.split()[0].rstrip(";") if s else "?"


executed, samples_by_op = Counter(), Counter()
per_line = []
total_samples = 0.0
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    op = opcode(src)
    ex = num(r[col["Instructions Executed"]])
    smp = num(r[col["# Samples"]])
    executed[op] += ex
    samples_by_op[op] += smp
    total_samples += smp
    stalls = sorted(((num(r[i]), n[6:]) for n, i in stall_cols), reverse=True)[:3]
    per_line.append((smp, r[col["Address"]][-5:], src.strip(), ex, stalls))

out = [f"# {kernel}", f"# report {args.report}: {len(per_line)} SASS instructions, {int(total_samples)} warp samples", ""]
tot_ex = sum(executed.values())
out.append(f"warp-instructions executed: {tot_ex:,.0f}")
out.append("  share      executed  opcode")
for op, n in executed.most_common(28):
    out.append(f"  {100 * n / tot_ex:5.1f}%  {n:12,.0f}  {op}")
out.append("")
out.append("warp-stall samples per opcode")
for op, n in samples_by_op.most_common(20):
    out.append(f"  {100 * n / max(total_samples, 1):5.1f}%  {op}")
out.append("")
out.append(f"top {args.top} instructions by samples (share, address, executed, instruction, dominant stall reasons)")
for smp, addr, src, ex, stalls in sorted(per_line, reverse=True)[: args.top]:
    why = ", ".join(f"{n} {int(v)}" for v, n in stalls if v > 0)
    out.append(f"  {100 * smp / max(total_samples, 1):5.2f}%  {addr}  {ex:10,.0f}  {src[:70]:70s}  {why}")
open(args.out, "w").write("\n".join(out) + "\n")
print("\n".join(out[:45]))
