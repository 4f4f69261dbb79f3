import csv, json, sys
rows = list(csv.reader(open('gpurun_out/launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, gi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size')
seq = [(r[ki][:60], float(r[vi].replace(',', '')), r[gi]) for r in data if len(r) > vi]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
for name, t, g in seq[-n:]:
    print(f"{t/1e3:10.1f} us  {g:>12}  {name}")
l = [x for x in open('gpurun_out/bench.log') if x.startswith('{')][0]
d = json.loads(l)
print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
for k, v in d['stages'].items():
    print(' ', k, {a: round(b, 4) for a, b in v.items()})
print('roofline', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d['roofline'].items() if k in ('achieved', 'frac', 'share_of_step')})
