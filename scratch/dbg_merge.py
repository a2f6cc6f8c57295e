import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "merzbild.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import merzbild_b200 as mb
from oracle import oracle
from parity_util import *
ctx = mb.Context(0, 1234)
for n, target in ((40, 20), (300, 100), (3000, 400), (6000, 800)):
    rng = np.random.default_rng(17)
    rows = maxwellian_rows(rng, n, 1.0, vw=True, w=1e15)
    opv, opia = oracle_state(oracle, rows, 1)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    oc = oracle.Octree(1, 1, 1, 6000, 10)
    oracle.merge_octree_N2(oracle.Rng.philox(1234, 3, 1), oc, opv, opia, 1, 1, 1, target)
    mb.merge_octree_N2_based(mb.PhiloxRng(3, 1), mb.OctreeN2Merge(1, 1, 1, max_Nbins=6000), pv, pia, 1, 1, target)
    nt = int(pia.n_total[0]); nto = int(opia.n_total[0])
    a, b = pv.logical(1, nt), opv.logical(1, nto)
    print("n", n, "nt", nt, nto, "Nbins oracle", oc.Nbins)
    if nt == nto:
        bad = np.where(np.abs(a - b).max(axis=1) > 1e-9 * np.abs(b).max())[0]
        print("  mismatching rows", len(bad), bad[:10])
        sa = a[np.lexsort((a[:, 1], a[:, 0]))]; sb = b[np.lexsort((b[:, 1], b[:, 0]))]
        print("  multiset equal:", np.allclose(sa, sb, rtol=1e-9, atol=0))
        if len(bad):
            i = bad[0]
            print("  dev", a[i]); print("  ora", b[i])
            # is it a sign flip about the pair mean?
            j = i + 1 if i % 2 == 0 else i - 1
            print("  pair dev", a[j]); print("  pair ora", b[j])
