"""Helpers shared by the GPU parity tests: build the same state in the CPU oracle and on the device, compare."""
import numpy as np

AR = 66.3e-27
HE = 6.65e-27


def mirror_to_device(mb, ctx, opv, opia, capacity=None):
    """Device twins of an oracle (pv, pia): same logical particles, same indexer."""
    n = len(opv)
    cap = capacity or n
    pv = mb.ParticleVector(cap, ctx)
    rows = opv.logical(1, n)
    pv.set_logical(1, rows)
    pv.set_cell(1, opv.cell[:n].copy())  # pv.cell is indexed by logical position
    pia = mb.ParticleIndexerArray(opia.n_cells, opia.n_species, ctx)
    pia.upload(opia.indexer.copy(), opia.n_total.copy(), opia.contiguous.copy())
    return pv, pia


def oracle_state(oracle, rows, n_cells, capacity=None, cell_of_all=1):
    """Oracle pv with `rows` at logical 1..n, all indexed from one cell (like the reference tests do before a sort)."""
    n = rows.shape[0]
    cap = capacity or n
    opv = oracle.OPV(cap)
    opv.fill_identity(rows) if n <= 2000 else _fill_fast(opv, rows)
    opia = oracle.OPIA(n_cells, 1)
    if n > 0:
        opia.indexer[0, cell_of_all - 1] = (n, 1, n, n, 0, -1, 0)
    opia.n_total[0] = n
    return opv, opia


def _fill_fast(opv, rows):
    n = rows.shape[0]
    # ParticleVector(np): index = 1:np, buffer = np:-1:1; add_particle!(pv, i, ...) for i = 1..n consumes buffer from the end,
    # i.e. physical slot i for logical i (particles.jl:210-212, :311-315, :739-742)
    opv.particles[:n] = rows
    opv.index[:n] = np.arange(1, n + 1)
    opv.nbuffer = len(opv) - n


def assert_same_pia(opia, pia, species=None):
    ix, nt, ct = pia.download()
    sl = slice(None) if species is None else slice(species - 1, species)
    np.testing.assert_array_equal(ix[sl], opia.indexer[sl])
    np.testing.assert_array_equal(nt[sl], opia.n_total[sl])
    np.testing.assert_array_equal(ct[sl], opia.contiguous[sl])


def assert_rows_close(a, b, rtol=1e-12, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    scale = np.maximum(np.abs(b).max(axis=0), 1e-300)
    err = (np.abs(a - b) / scale).max()
    assert err <= rtol, (what, err)


def maxwellian_rows(rng, n, L, T=300.0, m=AR, w=1.0, vw=False):
    rows = np.zeros((n, 7))
    rows[:, 0] = w * (rng.uniform(0.5, 2.0, n) if vw else 1.0)
    rows[:, 1:4] = rng.normal(0.0, np.sqrt(1.380649e-23 * T / m), (n, 3))
    rows[:, 4] = rng.uniform(0.0, L, n)
    rows[:, 5:7] = rng.uniform(0.0, 1.0, (n, 2))
    return rows
