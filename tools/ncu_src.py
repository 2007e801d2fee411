"""Summarise an `ncu --page source --csv` dump: executed-instruction share per opcode
and the hottest SASS lines by stall samples.  usage: ncu_src.py file.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: i for i, n in enumerate(hdr)}
ops = collections.Counter(); tot = 0; lines = []
for r in rows[hdr_i + 1:]:
  if len(r) < len(hdr) or r[0] == "Address" or not r[0].startswith("0x"): continue
  src = r[ci["Source"]].strip()
  ex = int(float(r[ci["Instructions Executed"]] or 0))
  st = int(float(r[ci["Warp Stall Sampling (All Samples)"]] or 0))
  toks = src.split()
  op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
  op = op.split(".")[0].rstrip(";")
  ops[op] += ex; tot += ex
  lines.append((st, ex, src))
print("total warp instructions executed:", tot)
for op, n in ops.most_common(22):
  print(f"  {op:10s} {n:10d}  {100.0*n/tot:5.1f}%")
print("hottest lines by stall samples:")
ssum = sum(l[0] for l in lines)
for st, ex, src in sorted(lines, reverse=True)[:top]:
  print(f"  {st:6d} ({100.0*st/max(ssum,1):4.1f}%) ex={ex:8d}  {src[:100]}")
