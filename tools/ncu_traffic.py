"""stdin: `ncu --page raw --csv`; stdout: JSON with per-kernel duration and DRAM bytes, and the totals."""
import csv, json, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
def col(name):
    return hdr.index(name)
out = {'kernels': [], 'dram_bytes': 0.0, 'duration_ms': 0.0}
units = rows[1]
def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
def to_ms(v, u):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1e-6)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    rd = to_bytes(r[col('dram__bytes_read.sum')], units[col('dram__bytes_read.sum')])
    wr = to_bytes(r[col('dram__bytes_write.sum')], units[col('dram__bytes_write.sum')])
    ms = to_ms(r[col('gpu__time_duration.sum')], units[col('gpu__time_duration.sum')])
    out['kernels'].append({'name': r[col('Kernel Name')][:60], 'ms': ms, 'dram_read': rd, 'dram_write': wr,
                           'fp64_pipe_pct': float(r[col('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')]),
                           'regs': int(float(r[col('launch__registers_per_thread')]))})
    out['dram_bytes'] += rd + wr
    out['duration_ms'] += ms
print(json.dumps(out, indent=1))
