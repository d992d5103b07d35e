"""Attribute executed warp-instructions / stall samples of a kernel in an
.ncu-rep to CUDA source lines (via nvdisasm -g line info of the built .so).
usage: python tools/ncu_lines.py rep.ncu-rep mangled_kernel_substring [topN]"""
import csv, re, subprocess, sys, os, tempfile, glob
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
tmp = tempfile.mkdtemp()
so = os.environ.get('LWB200_LIB') or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'lightweaver_b200', 'liblwb200.so')
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
dis = subprocess.run(['nvdisasm', '-g', '-c'] + glob.glob(tmp + '/*.cubin'), capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith('\t.section\t.text.') and kern in l)
lines = []
cur = ('?', 0)
for l in dis[start + 1:]:
    if l.startswith('\t.section') or l.startswith('//-------'):
        if lines:
            break
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
kid = os.environ.get('NCU_KERNEL_ID')  # e.g. ::regex:ray_kernel:1 when the report holds several kernels
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--kernel-id', kid] if kid else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
iI, iN, iS = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
body = rows[hi + 1:]
print('sass rows', len(body), 'disasm instrs', len(lines))
agg = {}
tot = 0
totS = 0
ops = {}
for r, loc in zip(body, lines):
    n, s = int(r[iI]), int(r[iN])
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += n; a[1] += s; a[2] += 1
    tot += n; totS += s
    op = r[iS].split()[0] if not r[iS].strip().startswith('@') else r[iS].split()[1]
    op = op.split('.')[0]
    ops[op] = ops.get(op, 0) + n
print('total warp-inst %.4g samples %d' % (tot, totS))
print('%7s %7s %5s  location' % ('inst%', 'stall%', 'sass'))
src_cache = {}
def src(loc):
    f, ln = loc
    for d in ('lightweaver_b200/csrc',):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < ln <= len(src_cache[p]):
                return src_cache[p][ln - 1].strip()[:90]
    return ''
for loc, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%6.2f%% %6.2f%% %5d  %s:%d  %s' % (100 * n / tot, 100 * s / max(totS, 1), c, loc[0], loc[1], src(loc)))
print('opcode mix:', ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))
