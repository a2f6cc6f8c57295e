// swpm! -- stochastic weighted particle method (collisions/collision_swpm.jl:201-287), CollisionFactorsSWPM (:17-23),
// compute_n_coll_single_species with w_max and G (:170-173).  Same skeleton as the NTC kernel (one thread per cell replays
// the reference's candidate loop with the cell's Philox stream), but every accepted pair sheds dw = min(w_i, w_k)/(1+G)
// into TWO new particles, and it is the children that are scattered (:261-283).  The per-cell tail window is therefore
// 2 * n_coll slots; windows are packed afterwards exactly as for the variable-weight NTC splits.
#include "mb_append.cuh"
#include "mb_common.cuh"
#include "mb_scan.cuh"

namespace mb {

struct SwpmArgs {
    SoA p;
    Indexer* ix;
    int64_t* n_total;
    int64_t cap;
    double* sgm;
    int64_t *n_coll, *n_perf;
    mb_interaction it;
    int64_t cell_lo, cell_hi;
    double G, dt, V;
    uint64_t seed;
    uint32_t timestep, substream;
    double* wmax;      // per cell of the range
    int32_t* ncoll32;  // 2 * n_coll (window size)
    int64_t* win;
    int32_t* nnew;
    int* flags;
    int single_cell_tail;
};

__device__ __forceinline__ double cell_wmax(const SoA& p, const Indexer& q) {  // :205-218
    double w = 0.0;
    for (int64_t i = q.start1 - 1; i < q.end1; i++) w = fmax(w, p.a[F_W][i]);
    if (q.n_group2 > 0)
        for (int64_t i = q.start2 - 1; i < q.end2; i++) w = fmax(w, p.a[F_W][i]);
    return w;
}
__device__ __forceinline__ int64_t swpm_ncoll(const SwpmArgs& a, double sgm, int64_t n, double w_max, double R) {
    return (int64_t)floor(0.5 * a.dt * (double)n * (double)(n - 1) * sgm * w_max * (a.G + 1) / a.V + R);
}

static __global__ void __launch_bounds__(128) k_swpm_prepass(SwpmArgs a) {
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const double wm = cell_wmax(a.p, q);
        a.wmax[r] = wm;
        PhiloxStream rng(a.seed, OP_SWPM, a.substream, a.timestep, (uint32_t)cell);
        int64_t nc = swpm_ncoll(a, a.sgm[cell - 1], q.n_local, wm, rng.rand());
        if (nc < 0) nc = 0;
        if (nc > 0x3fffffff) nc = 0x3fffffff;
        a.ncoll32[r] = (int32_t)(2 * nc);
    }
}

static __global__ void __launch_bounds__(128) k_swpm(SwpmArgs a) {
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const int64_t nt = *a.n_total;
    const int64_t wtot = a.win[nr];
    if (nt + wtot > a.cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicOr(&a.flags[0], DEVERR_CAPACITY);
            a.flags[1] = nt + wtot > 0x7fffffff ? 0x7fffffff : (int)(nt + wtot);
        }
        return;
    }
    const mb_interaction it = a.it;
    const double pw = 1.0 - 2 * it.vhs_o;
    const double wtf = 1.0 / (1.0 + a.G);
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = a.cell_lo + r;
        Indexer q = a.ix[cell - 1];
        const int64_t g2_before = q.n_group2;
        const int64_t win = nt + a.win[r];
        if (!(q.n_group2 == 0 || (a.single_cell_tail && q.end2 == nt))) {
            atomicOr(&a.flags[0], DEVERR_PRECONDITION);
            a.nnew[r] = 0;
            continue;
        }
        const double w_max = a.wmax[r];
        const double inv_w_max = 1.0 / w_max;
        PhiloxStream rng(a.seed, OP_SWPM, a.substream, a.timestep, (uint32_t)cell);
        double sgm = a.sgm[cell - 1];
        const int64_t n_coll = swpm_ncoll(a, sgm, q.n_local, w_max, rng.rand());
        int64_t n_perf = 0;
        for (int64_t c = 0; c < n_coll; c++) {
            int64_t i = (int64_t)floor(rng.rand() * (double)q.n_local);
            int64_t k = (int64_t)floor(rng.rand() * (double)q.n_local);
            while (i == k) k = (int64_t)floor(rng.rand() * (double)q.n_local);
            PRef pi, pk;
            load_p(a.p, map_cont(q, i), pi);
            load_p(a.p, map_cont(q, k), pk);
            const double gx = pi.vx - pk.vx, gy = pi.vy - pk.vy, gz = pi.vz - pk.vz;
            const double g = sqrt(gx * gx + gy * gy + gz * gz);
            if (!(g > EPS)) continue;
            const double sigma = it.vhs_factor * pow(g, pw);
            const double sg = sigma * g;
            sgm = fmax(sg, sgm);
            if (rng.rand() < sg * fmax(pi.w, pk.w) * inv_w_max / sgm) {
                n_perf += 1;
                const double cx = it.mu1 * pi.vx + it.mu2 * pk.vx, cy = it.mu1 * pi.vy + it.mu2 * pk.vy, cz = it.mu1 * pi.vz + it.mu2 * pk.vz;
                const double dw = fmin(pi.w, pk.w) * wtf;
                a.p.a[F_W][pi.pos] = pi.w - dw;
                a.p.a[F_W][pk.pos] = pk.w - dw;
                append_split(a.p, q, win, pi.pos, dw, pi.vx, pi.vy, pi.vz);
                const int64_t ci = q.end2 - 1;
                append_split(a.p, q, win, pk.pos, dw, pk.vx, pk.vy, pk.vz);
                const int64_t ck = q.end2 - 1;
                const double phi = twopi * rng.rand();
                double sphi, cphi;
                sincos(phi, &sphi, &cphi);
                const double ctheta = 2.0 * rng.rand() - 1.0;
                const double stheta = sqrt(1.0 - ctheta * ctheta);
                const double nx = g * (stheta * cphi), ny = g * (stheta * sphi), nz = g * ctheta;
                a.p.a[F_VX][ci] = cx + it.mu2 * nx; a.p.a[F_VY][ci] = cy + it.mu2 * ny; a.p.a[F_VZ][ci] = cz + it.mu2 * nz;
                a.p.a[F_VX][ck] = cx - it.mu1 * nx; a.p.a[F_VY][ck] = cy - it.mu1 * ny; a.p.a[F_VZ][ck] = cz - it.mu1 * nz;
            }
        }
        a.sgm[cell - 1] = sgm;
        a.n_coll[cell - 1] = n_coll;
        a.n_perf[cell - 1] = n_perf;
        a.ix[cell - 1] = q;
        a.nnew[r] = (int32_t)(q.n_group2 - g2_before);
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_swpm(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
                       double G, double dt, double V, uint32_t timestep, uint32_t substream) {
    MB_ARG(ctx && cf && it && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(cell_lo >= 1 && cell_hi <= pia->n_cells && cell_lo <= cell_hi, "cell range");
    MB_ARG(cf->n_cells == pia->n_cells && V > 0.0 && G >= 0.0, "swpm arguments");
    MB_CUDA(cudaSetDevice(ctx->device));
    int r = pv_ensure_alt(pv);
    if (r) return r;
    const int64_t nc = pia->n_cells, nr = cell_hi - cell_lo + 1, s = species - 1;
    ProfScope ps(ctx, PROF_NTC);
    ctx->state_gen++;
    SwpmArgs a;
    a.p = pv->cur;
    a.ix = pia->d_indexer + s * nc;
    a.n_total = pia->d_n_total + s;
    a.cap = pv->cap;
    a.sgm = cf->sigma_g_w_max;
    a.n_coll = cf->n_coll; a.n_perf = cf->n_coll_performed;
    a.it = *it;
    a.cell_lo = cell_lo; a.cell_hi = cell_hi;
    a.G = G; a.dt = dt; a.V = V;
    a.seed = stream_seed(ctx); a.timestep = timestep; a.substream = stream_substream(substream, species, species);
    a.flags = ctx->d_flags;
    a.single_cell_tail = nr == 1;
    int32_t* p32 = (int32_t*)ctx_scratch(ctx, 4, (size_t)(2 * nr) * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)3 * (nr + 1) + gs_partial_count(nr)) * 8);
    if (!p32 || !p64) return MB_ERR_CUDA;
    a.ncoll32 = p32; a.nnew = p32 + nr;
    a.win = p64;
    int64_t* packed = p64 + (nr + 1);
    a.wmax = (double*)(p64 + 2 * (nr + 1));
    int64_t* partial = p64 + 3 * (nr + 1);
    cudaStream_t st = ctx->stream;
    const int g = grid_for(nr, 128, 16);
    k_swpm_prepass<<<g, 128, 0, st>>>(a);
    MB_LAUNCH_CHECK(ctx);
    r = device_exclusive_scan(ctx, a.ncoll32, nr, a.win, partial);
    if (r) return r;
    MB_CUDA(cudaMemsetAsync(a.nnew, 0, (size_t)nr * 4, st));
    k_swpm<<<g, 128, 0, st>>>(a);
    MB_LAUNCH_CHECK(ctx);
    r = pack_windows(ctx, pv, a.ix, cell_lo, nr, a.win, a.nnew, packed, partial, a.n_total);
    if (r) return r;
    pia->sorted_layout[s] = 0;
    pia->n_bound[s] = pv->cap;
    pia->h_valid = false;
    return MB_OK;
}
