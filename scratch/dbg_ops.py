"""locate the op that raises the bad-cell flag in bench_ops at 2e7 particles"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "merzbild.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import merzbild_b200 as mb
from bench_ops import population, AR, DX, NDENS, DT

ctx = mb.Context(0, 1234)
it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
ppc = 150
nc = int(float(sys.argv[1]) * 0.6 // ppc)
a, ix, n = population(nc, ppc, 2, vw=True)
cap = int(n * 1.3)
pv, pia = mb.ParticleVector(cap, ctx), mb.ParticleIndexerArray(nc, 1, ctx)
grid = mb.Grid1DUniform(nc * DX, nc)
oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
pv.upload_soa(1, n, a)
pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))
x0 = a[4]
print("x range in", x0.min(), x0.max(), "L", grid.L, "max_x", grid.max_x, "cells", nc)
c_in = np.floor(x0 * grid.inv_dx).astype(np.int64)
print("host cells in: min", c_in.min(), "max", c_in.max())
def chk(name):
    try:
        ctx.sync(); print(name, "ok")
    except Exception as e:
        print(name, "FAILED", e)
mb.merge_octree_N2_based(mb.PhiloxRng(1), oc, pv, pia, (1, nc), 1, 100, grid, threshold=130); chk("merge")
mb.squash_pia(pv, pia, 1); chk("squash")
nt = int(pia.n_total[0])
xs = np.empty(nt); pv.download_soa(1, nt, [None]*4 + [xs, None, None])
cc = np.floor(xs * grid.inv_dx).astype(np.int64)
bad = np.where((cc < 0) | (cc >= nc) | ~np.isfinite(xs))[0]
print("after squash: n", nt, "bad", len(bad), xs[bad[:5]] if len(bad) else "", bad[:5])
okp, where = pia.check(1); print("pia check", okp, where)
mb.sort_particles(None, grid, pv, pia, 1); chk("sort")
