"""Quick GPU parity probe (development aid): CUDA path vs the C oracle."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
from oracle import oraclelib
from tests.util import compare_problems

for name, mk in (('tiny', lambda fs: synth.tiny_problem(formal_solver=fs)),
                 ('c1', lambda fs: synth.config_c1(formal_solver=fs))):
    for fs in (2, 1, 0):
        p = mk(fs); q = p.clone()
        ctx = Context(p); orc = oraclelib.OracleContext(q)
        for it in range(3):
            t0 = time.time(); u = ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0)); t1 = time.time()
            q.prefill_gamma(); dJ, idx = orc.fs_iter(lambdaIterate=(it == 0))
            e = compare_problems(p, q)
            print(name, 'fs', fs, 'it', it, 'dJ %.6e/%.6e idx %d/%d' % (u.dJMax, dJ, u.dJMaxIdx, idx),
                  ' '.join('%s=%.1e' % kv for kv in e.items()), 'gpu %.1f ms' % ((t1 - t0) * 1e3), flush=True)
            ctx.stat_equil(); orc.stat_eq()
            print('    after stat_eq n=%.1e' % compare_problems(p, q)['n'], flush=True)
        ctx.close()
