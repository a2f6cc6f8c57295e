"""CPU (gloo, world_size 2) cover of the host-side multi-rank logic: rendezvous of the 128-byte communicator id the way
bench.py does it, the ChunkSplitters slab rule of mb_grid1d_slab, and the ownership / conservation bookkeeping of the slab
exchange on a numpy model (the executable specification the CUDA + NCCL path is tested against on GPUs)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "merzbild.jl_b200"))
    import merzbild_b200 as mb

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = [bytes(range(128)) if rank == 0 else None]  # stands in for mb_comm_unique_id (NCCL needs a GPU)
    dist.broadcast_object_list(uid, src=0)
    assert uid[0] == bytes(range(128))
    nx = 11
    G = mb.Grid1DUniform(nx * 1e-5, nx)
    slab = G.slab(rank, world)
    rng = np.random.default_rng(5)
    n = 2000
    x = rng.uniform(0, G.L, n)
    w = 1.0 + np.arange(n)
    cell = np.floor(x * G.inv_dx).astype(np.int64)
    own = (cell >= slab.cell_offset) & (cell < slab.cell_offset + slab.n_cells)
    xs, ws = x[own], w[own]
    for step in range(10):
        xs = np.clip(xs + rng.normal(0, 0.8e-5, xs.shape), G.min_x, G.max_x)
        c = np.floor(xs * G.inv_dx).astype(np.int64) - slab.cell_offset
        left, right = c < 0, c >= slab.n_cells
        out = [None, None]
        # neighbour exchange: the model of ncclSend/ncclRecv of mb_exchange_slab
        send = {rank - 1: (xs[left], ws[left]), rank + 1: (xs[right], ws[right])}
        gathered = [None] * world
        dist.all_gather_object(gathered, send)
        keep = ~(left | right)
        arrivals = [g[rank] for g in gathered if rank in g and len(g[rank][0])]
        xs = np.concatenate([xs[keep]] + [a[0] for a in arrivals])
        ws = np.concatenate([ws[keep]] + [a[1] for a in arrivals])
        c = np.floor(xs * G.inv_dx).astype(np.int64) - slab.cell_offset
        assert c.min() >= 0 and c.max() < slab.n_cells  # one hop suffices: |dx| << slab width
    tot = [None] * world
    dist.all_gather_object(tot, ws)
    allw = np.sort(np.concatenate(tot))
    ok = np.array_equal(allw, w)
    q.put((rank, slab.n_cells, slab.cell_offset, ok))
    dist.destroy_process_group()


def test_slab_partition_and_exchange_model_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(6, 0), (5, 6)]  # first nx mod n slabs are one cell longer
    assert all(r[3] for r in res)  # every particle is owned by exactly one rank
