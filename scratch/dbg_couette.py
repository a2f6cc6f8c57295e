import json, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'merzbild.jl_b200'); sys.path.insert(0, 'tests')
import merzbild_b200 as mb
from oracle import oracle
AR = 66.3e-27
g = json.load(open('tests/golden/sparta_couette.json')); su = g['setup']; sp = np.array(g['cells'])
ctx = mb.Context(0, 4321)
L, nx, Fnum, dt = su['L'], su['nx'], su['fnum'], su['dt']
n = 50000
opv, opia = oracle.OPV(n), oracle.OPIA(nx, 1)
oracle.sample_equal_weight_grid(oracle.Rng.seq(1234), (L, nx), opv, opia, 1, AR, su['nrho'], su['T_init'], Fnum)
n = int(opia.n_total[0])
pv = mb.ParticleVector(int(1.3 * n), ctx); pv.set_logical(1, opv.logical(1, n))
pia = mb.ParticleIndexerArray(nx, 1, ctx); pia.upload(opia.indexer.copy(), opia.n_total.copy(), opia.contiguous.copy())
grid = mb.Grid1DUniform(L, nx); walls = mb.MaxwellWalls1D(300., 300., -500., 500., 1., 1.)
it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
cf = mb.CollisionFactors(nx, mb.estimate_sigma_g_w_max(it, AR, AR, 300., 300., Fnum), ctx)
pp, avg = mb.PhysProps(nx, 1, ndens_not_Np=True, ctx=ctx), mb.PhysProps(nx, 1, ndens_not_Np=True, ctx=ctx)
pxy=[]; pxx=[]
n_t, n_avg = int(sys.argv[1]), int(sys.argv[2]); surf = np.zeros((2, 11)); ncoll = 0
mb.sort_particles(None, grid, pv, pia, 1)
for t in range(1, n_t + 1):
    r = mb.PhiloxRng(t)
    mb.ntc_equal_weight(r, cf, None, it, pv, pia, (1, nx), 1, dt, L / nx)
    av = t > n_t - n_avg
    s = mb.convect_particles(r, grid, walls, pv, pia, 1, AR, dt, surf_props=av)
    mb.sort_particles(None, grid, pv, pia, 1)
    if av:
        surf += s / n_avg
        mb.compute_props_sorted([pv], pia, [AR], pp, grid); mb.avg_props(avg, pp, n_avg)
        if t % 200 == 0:
            rows = pv.logical(1, int(pia.n_total[0])); cidx = np.floor(rows[:,4]*grid.inv_dx).astype(int)
            uy = np.bincount(cidx, rows[:,2], minlength=nx)/np.bincount(cidx, minlength=nx)
            pxy.append(np.bincount(cidx, rows[:,0]*AR*rows[:,1]*(rows[:,2]-uy[cidx]), minlength=nx)/(L/nx))
            pxx.append(np.bincount(cidx, rows[:,0]*AR*rows[:,1]*rows[:,1], minlength=nx)/(L/nx))
        if t % 100 == 0: ncoll += cf.download()['n_coll_performed'].sum() / (n_avg / 100)
d = avg.download()
print('T', d['T'][0][:5], sp[:5, 1]); print('T mid', d['T'][0][23:27], sp[23:27, 1])
print('v', d['v'][0, :5, 1], sp[:5, 6]); print('n', d['n'][0][:3], sp[:3, 4])
print('surf L', surf[0]); print('surf R', surf[1]); print('sparta', g['boundary'])
P=np.mean(pxy,0); print('Pxy cells', P[:4], P[23:27], P[-4:], 'mean', P.mean()); print('Pxx mean', np.mean(pxx,0).mean())
print('coll per step', ncoll, 'sgwm', cf.download()['sigma_g_w_max'][:3])
