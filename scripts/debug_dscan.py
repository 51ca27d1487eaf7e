import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
ge.build()
from epseon_backend_b200 import cabi
from oracle import Oracle
from tests import workloads as W
w = W.c2()
orc = Oracle(omp=True, form=1)
A, i0, n, vmin = orc.prep(w['V'], w['s'])
ctx = cabi.Context(0)
ctx.set_option(ctx.OPT_FORM, 1)
ctx.set_potentials(w['V'], w['s'])
nE = 4096
dE = (w['E_hi'] - w['E_lo']) / (nE - 1)
n_o, m_o, x_o = orc.sweep_uniform(A, w['s'], w['E_lo'], dE, 0, nE)
for seg in (1, 0, 4, 18):
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, seg)
    for tails in (False, True):
        n_g, m_g, x_g = ctx.sweep_uniform(w['E_lo'], w['E_hi'], nE, tails=tails)
        bad = np.flatnonzero(n_g[0] != n_o)
        print('seg', seg, 'tails', tails, 'bad', bad.size, bad[:5], n_g[0][bad[:5]], n_o[bad[:5]], 'flagged', ctx.counter(ctx.CNT_SCAN_FLAGGED), 'scanl', ctx.counter(ctx.CNT_SCAN_LAUNCHES))
ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)
lev, wid, nb = ctx.solve_levels(w['E_lo'], w['E_hi'], 4096, 0, 16, 256, 1e-13, 12)
lev_o, *_ = orc.solve_levels(A, w['s'], w['E_lo'], w['E_hi'], 4096, 0, 16, 256, 1e-13, 12)
print(lev[0][:4], lev_o[:4], nb)
ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 1)
lev, wid, nb = ctx.solve_levels(w['E_lo'], w['E_hi'], 4096, 0, 16, 256, 1e-13, 12)
print('noscan', lev[0][:4], nb)
