"""Per-role stall breakdown of a warp-specialised kernel from `ncu --page source --csv` (stdin): rows are grouped by their
execution count (each role's loop body executes a distinct number of times)."""
import csv, collections, sys
r = csv.reader(sys.stdin); next(r); hdr = next(r)
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); ist = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and '(' not in h]
rows = [row for row in r if len(row) > ia]
tot = sum(int(x[ist]) for x in rows); print("total samples", tot)
cnt = collections.Counter(int(row[ia]) for row in rows)
for c, n in cnt.most_common(7):
    if c == 0: continue
    sel = [row for row in rows if int(row[ia]) == c]
    agg = collections.Counter()
    for row in sel:
        for i in stall_cols: agg[hdr[i][6:]] += int(row[i] or 0)
    print(f"exec x{c}: {n} instr, {sum(int(x[ist]) for x in sel)} samples:", agg.most_common(8))
top = sorted(rows, key=lambda x: -int(x[ist]))[:int(sys.argv[1]) if len(sys.argv) > 1 else 20]
for x in top:
    st = sorted(((int(x[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(x[ist], x[ia], x[isrc].strip()[:80], st)
