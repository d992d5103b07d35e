"""Steady-state device time of the phases of one iteration (development aid): the same call
sequence repeated back to back without host synchronisation, so differences between the rows
are what each phase adds on the GPU.    python tools/phase_times.py [c1|c2] [reps]"""
import sys, time
sys.path.insert(0, '.')
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
p = getattr(synth, 'config_' + wl)()
ctx = Context(p)
for _ in range(3):
    ctx.fs_iter_device(want_dJ=False); ctx.stat_eq_device()
def T(f):
    for _ in range(10): f()
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync(); return (time.perf_counter() - t0) / reps * 1e6
rows = [
    ('fs_iter, finalise deferred, no dJ', lambda: ctx.fs_iter_device(want_dJ=False, deferFinalise=True)),
    ('fs_iter, no dJ', lambda: ctx.fs_iter_device(want_dJ=False)),
    ('fs_iter, dJ async', lambda: ctx.fs_iter_device(asyncDJ=True)),
    ('fs_iter, dJ async + stat_eq async', lambda: (ctx.fs_iter_device(asyncDJ=True), ctx.stat_eq_device(wait=False))),
    ('fs_iter, dJ sync', lambda: ctx.fs_iter_device(want_dJ=True)),
    ('fs_iter, dJ async + stat_eq sync', lambda: (ctx.fs_iter_device(asyncDJ=True), ctx.stat_eq_device())),
    ('fs_iter, dJ async + finalise again', lambda: (ctx.fs_iter_device(asyncDJ=True), ctx.finalise())),
    ('fs_iter, dJ async + 2 x stat_eq async', lambda: (ctx.fs_iter_device(asyncDJ=True), ctx.stat_eq_device(wait=False), ctx.stat_eq_device(wait=False))),
    ('fs_iter, dJ async + stat_eq async, atom 0 only', lambda: (ctx.fs_iter_device(asyncDJ=True), ctx.stat_eq_device(atom=0, wait=False))),
    ('stat_eq async alone', lambda: ctx.stat_eq_device(wait=False)),
    ('finalise alone', lambda: ctx.finalise()),
]
for name, f in rows:
    print('%-40s %8.1f us   (launch set %.1f us)' % (name, T(f), ctx.kernel_time_ms() * 1e3), flush=True)
