"""Aggregate an `ncu --page source --csv` dump by SASS opcode (weighted by executed warp instructions)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = 0; ops = collections.Counter(); samp = collections.Counter()
for r in rows[2:]:
    if len(r) <= iN: continue
    try: n = int(r[iN]); s = int(r[iSamp])
    except ValueError: continue
    src = r[iS].strip()
    toks = src.split()
    op = toks[0]
    if op.startswith('@'): op = toks[1]
    op = op.split('.')[0] if not op.startswith('MUFU') else op
    ops[op] += n; samp[op] += s; tot += n
npx = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp instr {tot}", f" -> {tot*32/npx:.1f} thread-instr/pixel" if npx else "")
for op, n in ops.most_common(40):
    print(f"{op:14s} {n:12d} {100*n/tot:6.2f}%  samples {samp[op]:7d}" + (f"  {n*32/npx:7.2f}/px" if npx else ""))
