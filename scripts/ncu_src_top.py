"""Top SASS instructions of an `ncu --page source --csv` export by a chosen column.
usage: python scripts/ncu_src_top.py src.csv "L1 Wavefronts Shared" [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
H = rows[hi]
def num(x):
    try: return float(x.replace(',', ''))
    except ValueError: return 0.0
data = [r for r in rows[hi + 1:] if len(r) == len(H) and r[0] != 'Address']
key = sys.argv[2]; n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
k, ie, src, smp = H.index(key), H.index('Instructions Executed'), H.index('Source'), H.index('# Samples')
print('total', key, sum(num(r[k]) for r in data), ' total instr', sum(num(r[ie]) for r in data), 'samples', sum(num(r[smp]) for r in data))
for r in sorted(data, key=lambda r: -num(r[k]))[:n]:
    print(r[k].rjust(12), r[ie].rjust(10), r[smp].rjust(7), r[0][-5:], r[src][:120])
