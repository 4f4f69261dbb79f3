"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the metrics the
roofline discussion in DESIGN.md uses.  usage: python scripts/ncu_summary.py raw.csv [grep-substring...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H, U, data = rows[0], rows[1], rows[2:]
if len(sys.argv) > 2 and sys.argv[2] == '--list':
    for h, u in zip(H, U):
        if all(s in h for s in sys.argv[3:]): print(h, '[', u, ']')
    sys.exit(0)
M = [('time_ms', 'gpu__time_duration.sum'), ('dramR_GB', 'dram__bytes_read.sum'), ('dramW_GB', 'dram__bytes_write.sum'),
     ('sm%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
     ('tensor%', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed'),
     ('l1%', 'l1tex__throughput.avg.pct_of_peak_sustained_active'), ('l2%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
     ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), ('l2hit%', 'lts__t_sector_hit_rate.pct'),
     ('issue%', 'sm__inst_issued.avg.pct_of_peak_sustained_active'), ('regs', 'launch__registers_per_thread'),
     ('occ%', 'sm__warps_active.avg.pct_of_peak_sustained_active'), ('clk_MHz', 'sm__cycles_elapsed.avg.per_second')]
M = [(a, b) for a, b in M if b in H]
ki = H.index('Kernel Name')
print('kernel'.ljust(46), ' '.join(a.rjust(9) for a, _ in M))
for r in data:
    name = r[ki].replace('snvc::<unnamed>::', '').replace('void ', '').replace('unnamed>::', '')[:45]
    vals = []
    for a, b in M:
        i = H.index(b); v = r[i].replace(',', '')
        try:
            f = float(v)
            if U[i] == 'byte': f /= 1e9
            if U[i] == 'Kbyte': f /= 1e6
            if U[i] == 'Mbyte': f /= 1e3
            if U[i] == 'Tbyte': f *= 1e3
            if U[i] in ('ns', 'nsecond'): f /= 1e6
            if U[i] in ('us', 'usecond'): f /= 1e3
            if U[i] in ('s', 'second'): f *= 1e3
            if 'per_second' in b and U[i] in ('Ghz', 'GHz'): f *= 1e3
            if 'per_second' in b and U[i] in ('hz', 'Hz'): f /= 1e6
            vals.append(f'{f:9.3f}')
        except ValueError:
            vals.append(v[:9].rjust(9))
    print(name.ljust(46), ' '.join(vals))
