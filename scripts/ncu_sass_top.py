"""Summarise the SASS page of an ncu report: python scripts/ncu_sass_top.py rep.ncu-rep [launch_index] [min_pct]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; k = int(sys.argv[2]) if len(sys.argv) > 2 else 0; minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:120])
H = rows[1]
isrc, isamp, iex = H.index('Source'), H.index('# Samples'), H.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > isamp]
half = len(data)
# ncu prints the function twice in some versions; keep the first copy
for i in range(1, len(data)):
    if data[i][H.index('Address')] == data[0][H.index('Address')]:
        half = i; break
data = data[:half]
def n(x):
    try: return int(x)
    except Exception: return 0
tot = sum(n(r[isamp]) for r in data)
print('instructions', len(data), 'samples', tot)
key = re.compile(r'UTMALDG|UTCHMMA|UTCBAR|SYNCS|LDTM|STTM|BAR\.|ELECT|EXIT|LDG|STG|LDS|MUFU')
for i, r in enumerate(data):
    s = n(r[isamp])
    if s >= tot * minpct / 100 or (key.search(r[isrc]) and s > 0 and '--all' in sys.argv):
        print(f"{i:5d} {s:7d} {100*s/tot:5.1f}% ex={r[iex]:>9} {r[isrc][:100]}")
