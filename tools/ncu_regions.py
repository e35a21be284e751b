"""tools/ncu_regions.py src.csv -- instruction and stall-sample share per straight-line code region of an ncu source page."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))[2:]
tot = sum(int(r[5]) for r in rows); samp = sum(int(r[2]) for r in rows)
print("total inst %.2fM samples %d" % (tot / 1e6, samp))
blocks = []; cur = None
for i, r in enumerate(rows):
    n = int(r[5]); s = int(r[2])
    if cur and abs(n - cur['n']) <= 0.02 * max(n, cur['n'], 1):
        cur['len'] += 1; cur['inst'] += n; cur['samp'] += s; cur['end'] = i
    else:
        cur = {'start': i, 'end': i, 'n': n, 'len': 1, 'inst': n, 'samp': s}; blocks.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
for b in blocks:
    if b['inst'] > thr * tot or b['samp'] > thr * samp:
        print("%5d-%5d len %4d exec/instr %8d inst %6.2fM (%4.1f%%) samples %6d (%4.1f%%)  first: %s" % (
            b['start'], b['end'], b['len'], b['n'], b['inst'] / 1e6, 100 * b['inst'] / tot, b['samp'], 100 * b['samp'] / samp, rows[b['start']][1].strip()[:60]))
