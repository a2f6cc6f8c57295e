// Context and device containers of libmerzbild_b200.so: mb_ctx, mb_pv (ParticleVector, particles.jl:194-212),
// mb_pia (ParticleIndexerArray, particles.jl:104-141), mb_cf (CollisionFactors, collision_ntc.jl:18-25),
// Grid1DUniform (grid_uniform1D.jl:49-86) and Interaction (collision_utils.jl:73-101) helpers.
#include <cstring>
#include <mutex>

#include <cstdlib>

#include "mb_common.cuh"

namespace mb {

static thread_local std::string g_err = "";
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return MB_ERR_NO_DEVICE;
    return MB_ERR_CUDA;
}

void* ctx_scratch(mb_ctx* ctx, int slot, size_t bytes) {
    if (ctx->scratch_bytes[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    if (ctx->scratch[slot]) {
        // stream-ordered free: earlier kernels that use the old block are still safe
        cudaFreeAsync(ctx->scratch[slot], ctx->stream);
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, want, ctx->stream);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMallocAsync(scratch)", __FILE__, __LINE__);
        return nullptr;
    }
    ctx->scratch[slot] = p;
    ctx->scratch_bytes[slot] = want;
    return p;
}

// sections may nest (the squash folded into the general sort path): every begin takes its own event pair, a stack pairs the ends
void prof_begin(mb_ctx* c, int section) {
    if (c->prof_used == c->prof_sec->size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        c->prof_ev->push_back(a);
        c->prof_ev->push_back(b);
        c->prof_sec->push_back(section);
    }
    const size_t i = c->prof_used++;
    (*c->prof_sec)[i] = section;
    c->prof_stack->push_back(i);
    cudaEventRecord((*c->prof_ev)[2 * i], c->stream);
}
void prof_end(mb_ctx* c) {
    if (c->prof_stack->empty()) return;
    const size_t i = c->prof_stack->back();
    c->prof_stack->pop_back();
    cudaEventRecord((*c->prof_ev)[2 * i + 1], c->stream);
}

static int alloc_soa(SoA& s, int64_t cap) {
    for (int f = 0; f < 7; f++) s.a[f] = nullptr;
    if (cap <= 0) return MB_OK;
    // one allocation, 256-byte aligned slices so 128-bit vector accesses are always legal
    const size_t stride = (((size_t)cap * 8 + 255) / 256) * 256;
    char* base = nullptr;
    MB_CUDA(cudaMalloc(&base, stride * 7));
    for (int f = 0; f < 7; f++) s.a[f] = (double*)(base + stride * f);
    return MB_OK;
}
static void free_soa(SoA& s) {
    if (s.a[0]) cudaFree(s.a[0]);
    for (int f = 0; f < 7; f++) s.a[f] = nullptr;
}
int pv_ensure_alt(mb_pv* pv) {
    if (pv->has_alt) return MB_OK;
    int r = alloc_soa(pv->alt, pv->cap);
    if (r) return r;
    pv->has_alt = true;
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

const char* mb_last_error_string(void) { return g_err.c_str(); }
int mb_version(void) { return 100; }

int mb_ctx_create(int device, uint64_t seed, mb_ctx** out) {
    MB_ARG(out != nullptr, "out == NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(std::string("no CUDA device available (libmerzbild_b200 has no CPU fallback): ") + cudaGetErrorString(e));
        cudaGetLastError();
        return MB_ERR_NO_DEVICE;
    }
    MB_ARG(device >= 0 && device < ndev, "device out of range");
    MB_CUDA(cudaSetDevice(device));
    mb_ctx* c = new mb_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->seed = seed;
    c->band_w = 2;
    c->state_gen = 1;
    {   // L2 fetch granularity (experiment knob): the NTC gathers use 8 bytes of every sector they touch
        const char* e = getenv("MB_L2_FETCH");
        if (e) {
            cudaError_t le = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            fprintf(stderr, "[mb] cudaLimitMaxL2FetchGranularity <- %s: %s, now %zu\n", e, cudaGetErrorString(le), got);
            cudaGetLastError();
        }
    }
    MB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    MB_CUDA(cudaMalloc(&c->d_flags, 16 * sizeof(int)));
    MB_CUDA(cudaMemset(c->d_flags, 0, 16 * sizeof(int)));
    MB_CUDA(cudaHostAlloc(&c->h_flags, 16 * sizeof(int), cudaHostAllocDefault));
    MB_CUDA(cudaEventCreate(&c->ev0));
    MB_CUDA(cudaEventCreate(&c->ev1));
    MB_CUDA(cudaMalloc(&c->d_xch_counts, 12 * sizeof(int64_t)));  // [8..9]: leavers packed by the edge exchange, [10]: particles the sort dropped
    MB_CUDA(cudaMemset(c->d_xch_counts, 0, 12 * sizeof(int64_t)));
    MB_CUDA(cudaHostAlloc(&c->h_xch_counts, 8 * sizeof(int64_t), cudaHostAllocDefault));
    c->nranks = 1;
    c->prof_ev = new std::vector<cudaEvent_t>();
    c->prof_sec = new std::vector<int>();
    c->prof_stack = new std::vector<size_t>();
    *out = c;
    return MB_OK;
}

int mb_ctx_destroy(mb_ctx* c) {
    if (!c) return MB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 24; i++)
        if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->l2_scratch) cudaFree(c->l2_scratch);
    for (int i = 0; i < 2; i++) {
        if (c->xch_send[i]) cudaFree(c->xch_send[i]);
        if (c->xch_recv[i]) cudaFree(c->xch_recv[i]);
    }
    cudaFree(c->d_flags);
    cudaFreeHost(c->h_flags);
    cudaFree(c->d_xch_counts);
    cudaFreeHost(c->h_xch_counts);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (auto e : *c->prof_ev) cudaEventDestroy(e);
    delete c->prof_ev;
    delete c->prof_sec;
    delete c->prof_stack;
    cudaStreamDestroy(c->stream);
    delete c;
    return MB_OK;
}

int mb_sync(mb_ctx* c) {
    MB_ARG(c != nullptr, "ctx == NULL");
    MB_CUDA(cudaSetDevice(c->device));
    MB_CUDA(cudaMemcpyAsync(c->h_flags, c->d_flags, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    const int f = c->h_flags[0];
    if (f) {
        MB_CUDA(cudaMemsetAsync(c->d_flags, 0, 2 * sizeof(int), c->stream));
        char buf[256];
        snprintf(buf, sizeof buf, "device-side error flags 0x%x (1 capacity, 2 precondition, 4 band overflow, 8 bad cell, 16 octree); aux=%d", f,
                 c->h_flags[1]);
        set_error(buf);
        if (f & DEVERR_CAPACITY) return MB_ERR_CAPACITY;
        if (f & (DEVERR_PRECONDITION | DEVERR_BAD_CELL)) return MB_ERR_PRECONDITION;
        return MB_ERR_UNSUPPORTED;
    }
    return MB_OK;
}

void* mb_ctx_stream(mb_ctx* c) { return c ? (void*)c->stream : nullptr; }
int mb_ctx_set_seed(mb_ctx* c, uint64_t seed) {
    MB_ARG(c != nullptr, "ctx == NULL");
    c->seed = seed;
    return MB_OK;
}
int64_t mb_ctx_kernel_launches(mb_ctx* c) { return c ? c->n_launch : 0; }

int mb_timer_start(mb_ctx* c) {
    MB_ARG(c != nullptr, "ctx == NULL");
    MB_CUDA(cudaSetDevice(c->device));
    MB_CUDA(cudaEventRecord(c->ev0, c->stream));
    return MB_OK;
}
int mb_timer_stop(mb_ctx* c, double* ms) {
    MB_ARG(c != nullptr && ms != nullptr, "NULL");
    MB_CUDA(cudaEventRecord(c->ev1, c->stream));
    MB_CUDA(cudaEventSynchronize(c->ev1));
    float f = 0;
    MB_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = f;
    return MB_OK;
}
int mb_prof_enable(mb_ctx* c, int32_t on) {
    MB_ARG(c != nullptr, "ctx == NULL");
    MB_CUDA(cudaStreamSynchronize(c->stream));
    c->prof_on = on;
    c->prof_used = 0;
    c->prof_stack->clear();
    return MB_OK;
}
int mb_prof_read(mb_ctx* c, int32_t section, double* total_ms, int64_t* launches) {
    MB_ARG(c && total_ms && launches && section >= 0 && section < PROF_NSEC, "prof_read");
    MB_CUDA(cudaSetDevice(c->device));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    double t = 0;
    int64_t n = 0;
    for (size_t i = 0; i < c->prof_used; i++) {
        if ((*c->prof_sec)[i] != section) continue;
        float ms = 0;
        MB_CUDA(cudaEventElapsedTime(&ms, (*c->prof_ev)[2 * i], (*c->prof_ev)[2 * i + 1]));
        t += ms;
        n++;
        (*c->prof_sec)[i] = -1;
    }
    *total_ms = t;
    *launches = n;
    bool any = false;
    for (size_t i = 0; i < c->prof_used; i++) any |= (*c->prof_sec)[i] >= 0;
    if (!any) c->prof_used = 0;
    return MB_OK;
}
int mb_flush_l2(mb_ctx* c) {
    MB_ARG(c != nullptr, "ctx == NULL");
    MB_CUDA(cudaSetDevice(c->device));
    if (!c->l2_scratch) {
        c->l2_scratch_bytes = (size_t)256 << 20;
        MB_CUDA(cudaMalloc(&c->l2_scratch, c->l2_scratch_bytes));
    }
    MB_CUDA(cudaMemsetAsync(c->l2_scratch, 1, c->l2_scratch_bytes, c->stream));
    return MB_OK;
}

// ---- grid ----
int mb_grid1d_init(double L, int64_t nx, double wall_offset, mb_grid1d* g) {  // grid_uniform1D.jl:72-86
    MB_ARG(g != nullptr && nx > 0 && L > 0, "grid");
    g->L = L;
    g->n_cells = nx;
    g->dx = L / (double)nx;
    g->inv_dx = 1.0 / g->dx;
    g->min_x = g->dx * wall_offset;
    g->max_x = L - g->dx * wall_offset;
    g->cell_offset = 0;
    return MB_OK;
}
int mb_grid1d_slab(const mb_grid1d* G, int rank, int nranks, mb_grid1d* out) {
    MB_ARG(G && out && nranks > 0 && rank >= 0 && rank < nranks, "slab");
    const int64_t nx = G->n_cells, base = nx / nranks, rem = nx % nranks;
    const int64_t lo = rank * base + (rank < rem ? rank : rem);
    *out = *G;
    out->n_cells = base + (rank < rem ? 1 : 0);
    out->cell_offset = G->cell_offset + lo;
    return MB_OK;
}

// ---- ParticleVector ----
int mb_pv_create(mb_ctx* ctx, int64_t np, mb_pv** out) {
    MB_ARG(ctx && out && np >= 0, "pv_create");
    MB_CUDA(cudaSetDevice(ctx->device));
    mb_pv* p = new mb_pv();
    p->ctx = ctx;
    p->cap = np;
    p->has_alt = false;
    p->drop_oob = 0;
    p->n_arrivals = 0;
    p->arrivals_at_end = 0;
    p->cell = nullptr;
    p->d_n_arr = nullptr;
    for (int f = 0; f < 7; f++) p->alt.a[f] = nullptr;
    int r = alloc_soa(p->cur, np);
    if (r) { delete p; return r; }
    MB_CUDA(cudaMalloc(&p->d_n_arr, sizeof(int64_t)));
    MB_CUDA(cudaMemsetAsync(p->d_n_arr, 0, sizeof(int64_t), ctx->stream));
    if (np > 0) {
        MB_CUDA(cudaMalloc(&p->cell, (size_t)np * sizeof(int32_t)));
        // Particle(0, [0,0,0], [0,0,0]) everywhere (particles.jl:210)
        const size_t stride = (((size_t)np * 8 + 255) / 256) * 256;
        MB_CUDA(cudaMemsetAsync(p->cur.a[0], 0, stride * 7, ctx->stream));
        MB_CUDA(cudaMemsetAsync(p->cell, 0, (size_t)np * sizeof(int32_t), ctx->stream));
    }
    *out = p;
    return MB_OK;
}
int mb_pv_destroy(mb_pv* p) {
    if (!p) return MB_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    free_soa(p->cur);
    if (p->has_alt) free_soa(p->alt);
    if (p->cell) cudaFree(p->cell);
    if (p->d_n_arr) cudaFree(p->d_n_arr);
    delete p;
    return MB_OK;
}
int64_t mb_pv_length(mb_pv* p) { return p ? p->cap : -1; }

int mb_pv_resize(mb_pv* p, int64_t np) {  // particles.jl:269-298 (new slots hold zero-weight particles)
    MB_ARG(p && np >= p->cap, "resize can only grow");
    if (np == p->cap) return MB_OK;
    mb_ctx* ctx = p->ctx;
    ctx->state_gen++;
    MB_CUDA(cudaSetDevice(ctx->device));
    SoA ns;
    int r = alloc_soa(ns, np);
    if (r) return r;
    const size_t stride = (((size_t)np * 8 + 255) / 256) * 256;
    MB_CUDA(cudaMemsetAsync(ns.a[0], 0, stride * 7, ctx->stream));
    for (int f = 0; f < 7; f++)
        if (p->cap > 0) MB_CUDA(cudaMemcpyAsync(ns.a[f], p->cur.a[f], (size_t)p->cap * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    int32_t* nc = nullptr;
    MB_CUDA(cudaMalloc(&nc, (size_t)np * sizeof(int32_t)));
    MB_CUDA(cudaMemsetAsync(nc, 0, (size_t)np * sizeof(int32_t), ctx->stream));
    if (p->cap > 0) MB_CUDA(cudaMemcpyAsync(nc, p->cell, (size_t)p->cap * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    free_soa(p->cur);
    if (p->has_alt) { free_soa(p->alt); p->has_alt = false; }
    if (p->cell) cudaFree(p->cell);
    p->cur = ns;
    p->cell = nc;
    p->cap = np;
    return MB_OK;
}

static int check_range(mb_pv* p, int64_t lo, int64_t n) {
    MB_ARG(p != nullptr, "pv == NULL");
    MB_ARG(lo >= 1 && n >= 0 && lo - 1 + n <= p->cap, "logical range outside 1..length(pv)");
    return MB_OK;
}

int mb_pv_upload_soa(mb_pv* p, int64_t lo, int64_t n, const double* w, const double* vx, const double* vy, const double* vz, const double* x,
                     const double* y, const double* z) {
    int r = check_range(p, lo, n);
    if (r) return r;
    if (n == 0) return MB_OK;
    MB_CUDA(cudaSetDevice(p->ctx->device));
    p->ctx->state_gen++;
    const double* src[7] = {w, vx, vy, vz, x, y, z};
    for (int f = 0; f < 7; f++)
        if (src[f]) MB_CUDA(cudaMemcpyAsync(p->cur.a[f] + (lo - 1), src[f], (size_t)n * 8, cudaMemcpyHostToDevice, p->ctx->stream));
    return MB_OK;
}
int mb_pv_download_soa(mb_pv* p, int64_t lo, int64_t n, double* w, double* vx, double* vy, double* vz, double* x, double* y, double* z) {
    int r = check_range(p, lo, n);
    if (r) return r;
    MB_CUDA(cudaSetDevice(p->ctx->device));
    double* dst[7] = {w, vx, vy, vz, x, y, z};
    for (int f = 0; f < 7; f++)
        if (dst[f] && n > 0) MB_CUDA(cudaMemcpyAsync(dst[f], p->cur.a[f] + (lo - 1), (size_t)n * 8, cudaMemcpyDeviceToHost, p->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return MB_OK;
}
int mb_pv_upload_rows(mb_pv* p, int64_t lo, int64_t n, const double* rows) {
    int r = check_range(p, lo, n);
    if (r) return r;
    MB_ARG(rows != nullptr || n == 0, "rows == NULL");
    if (n == 0) return MB_OK;
    std::vector<double> t((size_t)n * 7);
    for (int64_t i = 0; i < n; i++)
        for (int f = 0; f < 7; f++) t[(size_t)f * n + i] = rows[(size_t)i * 7 + f];
    r = mb_pv_upload_soa(p, lo, n, &t[0], &t[(size_t)n], &t[(size_t)2 * n], &t[(size_t)3 * n], &t[(size_t)4 * n], &t[(size_t)5 * n], &t[(size_t)6 * n]);
    if (r) return r;
    MB_CUDA(cudaStreamSynchronize(p->ctx->stream));  // staging buffer goes out of scope
    return MB_OK;
}
int mb_pv_download_rows(mb_pv* p, int64_t lo, int64_t n, double* rows) {
    int r = check_range(p, lo, n);
    if (r) return r;
    if (n == 0) return MB_OK;
    std::vector<double> t((size_t)n * 7);
    r = mb_pv_download_soa(p, lo, n, &t[0], &t[(size_t)n], &t[(size_t)2 * n], &t[(size_t)3 * n], &t[(size_t)4 * n], &t[(size_t)5 * n], &t[(size_t)6 * n]);
    if (r) return r;
    for (int64_t i = 0; i < n; i++)
        for (int f = 0; f < 7; f++) rows[(size_t)i * 7 + f] = t[(size_t)f * n + i];
    return MB_OK;
}
int mb_pv_upload_cell(mb_pv* p, int64_t lo, int64_t n, const int64_t* cell) {
    int r = check_range(p, lo, n);
    if (r) return r;
    if (n == 0) return MB_OK;
    std::vector<int32_t> t((size_t)n);
    for (int64_t i = 0; i < n; i++) t[i] = (int32_t)cell[i];
    MB_CUDA(cudaSetDevice(p->ctx->device));
    p->ctx->state_gen++;
    MB_CUDA(cudaMemcpyAsync(p->cell + (lo - 1), t.data(), (size_t)n * 4, cudaMemcpyHostToDevice, p->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return MB_OK;
}
int mb_pv_download_cell(mb_pv* p, int64_t lo, int64_t n, int64_t* cell) {
    int r = check_range(p, lo, n);
    if (r) return r;
    if (n == 0) return MB_OK;
    std::vector<int32_t> t((size_t)n);
    MB_CUDA(cudaSetDevice(p->ctx->device));
    MB_CUDA(cudaMemcpyAsync(t.data(), p->cell + (lo - 1), (size_t)n * 4, cudaMemcpyDeviceToHost, p->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(p->ctx->stream));
    for (int64_t i = 0; i < n; i++) cell[i] = t[i];
    return MB_OK;
}
int mb_pv_device_ptrs(mb_pv* p, void** out7) {
    MB_ARG(p && out7, "NULL");
    for (int f = 0; f < 7; f++) out7[f] = p->cur.a[f];
    p->ctx->state_gen++;  // the caller may write through the pointers: cached moments / classification are void
    return MB_OK;
}

// ---- ParticleIndexerArray ----
__global__ void k_pia_init(Indexer* ix, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        ix[i] = Indexer{0, 0, -1, 0, 0, -1, 0};  // particles.jl:84
}

int mb_pia_create(mb_ctx* ctx, int64_t n_cells, int64_t n_species, mb_pia** out) {
    MB_ARG(ctx && out && n_cells > 0 && n_species > 0, "pia_create");
    MB_CUDA(cudaSetDevice(ctx->device));
    mb_pia* p = new mb_pia();
    p->ctx = ctx;
    p->n_cells = n_cells;
    p->n_species = n_species;
    MB_CUDA(cudaMalloc(&p->d_indexer, (size_t)n_cells * n_species * sizeof(Indexer)));
    MB_CUDA(cudaMalloc(&p->d_n_total, (size_t)n_species * 8));
    MB_CUDA(cudaMalloc(&p->d_holes, (size_t)n_species * sizeof(int)));
    MB_CUDA(cudaMemsetAsync(p->d_holes, 0, (size_t)n_species * sizeof(int), ctx->stream));
    MB_CUDA(cudaMemsetAsync(p->d_n_total, 0, (size_t)n_species * 8, ctx->stream));
    MB_CUDA(cudaHostAlloc(&p->h_n_total, (size_t)n_species * 8, cudaHostAllocDefault));
    for (int64_t s = 0; s < n_species; s++) p->h_n_total[s] = 0;
    p->h_valid = true;
    p->contiguous.assign(n_species, 1);
    p->sorted_layout.assign(n_species, 0);
    p->n_bound.assign(n_species, 0);
    p->contig_pending.assign(n_species, 0);
    k_pia_init<<<grid_for(n_cells * n_species, 256), 256, 0, ctx->stream>>>(p->d_indexer, n_cells * n_species);
    MB_LAUNCH_CHECK(ctx);
    *out = p;
    return MB_OK;
}
int mb_pia_destroy(mb_pia* p) {
    if (!p) return MB_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->d_indexer);
    cudaFree(p->d_n_total);
    cudaFree(p->d_holes);
    cudaFreeHost(p->h_n_total);
    delete p;
    return MB_OK;
}
int mb_pia_upload(mb_pia* p, const int64_t* indexer, const int64_t* n_total, const uint8_t* contiguous) {
    MB_ARG(p != nullptr, "pia == NULL");
    mb_ctx* ctx = p->ctx;
    ctx->state_gen++;
    MB_CUDA(cudaSetDevice(ctx->device));
    if (indexer) {
        MB_CUDA(cudaMemcpyAsync(p->d_indexer, indexer, (size_t)p->n_cells * p->n_species * sizeof(Indexer), cudaMemcpyHostToDevice, ctx->stream));
        for (auto& f : p->sorted_layout) f = 0;
    }
    if (n_total) {
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int64_t s = 0; s < p->n_species; s++) { p->h_n_total[s] = n_total[s]; p->n_bound[s] = n_total[s]; }
        MB_CUDA(cudaMemcpyAsync(p->d_n_total, p->h_n_total, (size_t)p->n_species * 8, cudaMemcpyHostToDevice, ctx->stream));
        p->h_valid = true;
    }
    if (contiguous)
        for (int64_t s = 0; s < p->n_species; s++) { p->contiguous[s] = contiguous[s]; p->contig_pending[s] = 0; }
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MB_OK;
}
static int pia_refresh_host(mb_pia* p) {
    if (p->h_valid) return MB_OK;
    MB_CUDA(cudaSetDevice(p->ctx->device));
    MB_CUDA(cudaMemcpyAsync(p->h_n_total, p->d_n_total, (size_t)p->n_species * 8, cudaMemcpyDeviceToHost, p->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(p->ctx->stream));
    p->h_valid = true;
    for (int64_t s = 0; s < p->n_species; s++) p->n_bound[s] = p->h_n_total[s];
    return MB_OK;
}
int mb_pia_download(mb_pia* p, int64_t* indexer, int64_t* n_total, uint8_t* contiguous) {
    MB_ARG(p != nullptr, "pia == NULL");
    MB_CUDA(cudaSetDevice(p->ctx->device));
    if (indexer)
        MB_CUDA(cudaMemcpyAsync(indexer, p->d_indexer, (size_t)p->n_cells * p->n_species * sizeof(Indexer), cudaMemcpyDeviceToHost, p->ctx->stream));
    p->h_valid = false;
    int r = pia_refresh_host(p);
    if (r) return r;
    {   // resolve the exact contiguous flag after merges (merging_octree_N2.jl:806-808): the pia's device flag of the species
        bool any = false;
        for (int64_t s = 0; s < p->n_species; s++) any |= p->contig_pending[s] != 0;
        if (any) {
            std::vector<int> f((size_t)p->n_species);
            MB_CUDA(cudaMemcpyAsync(f.data(), p->d_holes, (size_t)p->n_species * sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
            MB_CUDA(cudaStreamSynchronize(p->ctx->stream));
            for (int64_t s = 0; s < p->n_species; s++)
                if (p->contig_pending[s]) {  // only the species being resolved: the flags are per pia and per species
                    if (f[s] == 0) p->contiguous[s] = 1;  // every deletion fitted into group 2 of the last cell
                    p->contig_pending[s] = 0;
                    MB_CUDA(cudaMemsetAsync(p->d_holes + s, 0, sizeof(int), p->ctx->stream));
                }
        }
    }
    if (n_total)
        for (int64_t s = 0; s < p->n_species; s++) n_total[s] = p->h_n_total[s];
    if (contiguous)
        for (int64_t s = 0; s < p->n_species; s++) contiguous[s] = p->contiguous[s];
    return mb_sync(p->ctx);
}
int64_t mb_pia_n_total(mb_pia* p, int64_t species) {
    if (!p || species < 1 || species > p->n_species) return -1;
    if (pia_refresh_host(p)) return -1;
    return p->h_n_total[species - 1];
}

__global__ void k_check_pia(const Indexer* ix, int64_t n_cells, const int64_t* n_total, int* ok, unsigned long long* where, unsigned long long* sum) {
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_cells; i += (int64_t)gridDim.x * blockDim.x) {
        const Indexer q = ix[i];
        bool bad = q.n_local != q.n_group1 + q.n_group2;
        if (q.n_group1 > 0) bad |= q.n_group1 != q.end1 - q.start1 + 1;
        else bad |= (q.start1 != 0 || q.end1 != -1);
        if (q.n_group2 > 0) bad |= q.n_group2 != q.end2 - q.start2 + 1;
        else bad |= (q.start2 != 0 || q.end2 != -1);
        if (bad) { *ok = 0; atomicMin(where, (unsigned long long)(i + 1)); }
        local += (unsigned long long)q.n_local;
    }
    atomicAdd(sum, local);
}
int mb_check_pia(mb_pia* p, int64_t species, int32_t* ok, int64_t* where) {  // particles.jl:863-907
    MB_ARG(p && ok && where && species >= 1 && species <= p->n_species, "check_pia");
    mb_ctx* ctx = p->ctx;
    MB_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* d = (unsigned long long*)ctx_scratch(ctx, 7, 64);
    if (!d) return MB_ERR_CUDA;
    unsigned long long init[3] = {1ull, ~0ull, 0ull};
    MB_CUDA(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    k_check_pia<<<grid_for(p->n_cells, 256), 256, 0, ctx->stream>>>(p->d_indexer + (species - 1) * p->n_cells, p->n_cells, p->d_n_total + species - 1,
                                                                   (int*)d, d + 1, d + 2);
    MB_LAUNCH_CHECK(ctx);
    unsigned long long h[3];
    MB_CUDA(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    p->h_valid = false;
    const int64_t nt = mb_pia_n_total(p, species);
    if ((int)(h[0] & 0xffffffffu) == 0) { *ok = 0; *where = (int64_t)h[1]; }
    else { *where = 0; *ok = ((int64_t)h[2] == nt) ? 1 : 0; }
    return MB_OK;
}

// ---- Interaction / CollisionFactors ----
int mb_make_interaction(double m_i, double m_k, double d, double o, double Tref, mb_interaction* it) {  // collision_utils.jl:159-201
    MB_ARG(it != nullptr, "NULL");
    it->m_r = m_i * m_k / (m_i + m_k);
    it->mu1 = m_i / (m_i + m_k);
    it->mu2 = m_k / (m_i + m_k);
    it->vhs_d = d; it->vhs_o = o; it->vhs_Tref = Tref;
    const double m_half = 0.5 * (m_i + m_k);  // compute_mu_ref gets 0.5*(m_i+m_k), collision_utils.jl:188
    it->vhs_muref = 30.0 * std::sqrt(m_half * k_B * Tref) / (4.0 * std::sqrt(M_PI) * (5.0 - 2.0 * o) * (7.0 - 2.0 * o) * d * d);
    it->vhs_factor = M_PI * d * d * std::pow(2 * k_B * Tref / it->m_r, o - 0.5) / std::tgamma(2.5 - o);  // :98-101
    return MB_OK;
}
double mb_estimate_sigma_g_w_max(const mb_interaction* it, double m1, double m2, double T1, double T2, double Fnum, double mult) {  // :418-423
    const double g1 = std::sqrt(2 * T1 * k_B / m1), g2 = std::sqrt(2 * T2 * k_B / m2);
    const double g = 0.5 * (g1 + g2);
    return mult * (it->vhs_factor * std::pow(g, 1.0 - 2 * it->vhs_o)) * g * Fnum;
}

__global__ void k_fill_f64(double* p, int64_t n, double v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
int mb_cf_fill(mb_cf* cf, double v) {
    MB_ARG(cf != nullptr, "NULL");
    MB_CUDA(cudaSetDevice(cf->ctx->device));
    k_fill_f64<<<grid_for(cf->n_cells, 256), 256, 0, cf->ctx->stream>>>(cf->sigma_g_w_max, cf->n_cells, v);
    MB_LAUNCH_CHECK(cf->ctx);
    return MB_OK;
}
int mb_cf_create(mb_ctx* ctx, int64_t n_cells, double sgwm, mb_cf** out) {
    MB_ARG(ctx && out && n_cells > 0, "cf_create");
    MB_CUDA(cudaSetDevice(ctx->device));
    mb_cf* c = new mb_cf();
    c->ctx = ctx;
    c->n_cells = n_cells;
    MB_CUDA(cudaMalloc(&c->sigma_g_w_max, (size_t)n_cells * 8));
    MB_CUDA(cudaMalloc(&c->n_coll, (size_t)n_cells * 8 * 3));
    c->n_coll_performed = c->n_coll + n_cells;
    c->n_eq_w = c->n_coll + 2 * n_cells;
    MB_CUDA(cudaMemsetAsync(c->n_coll, 0, (size_t)n_cells * 8 * 3, ctx->stream));
    *out = c;
    return mb_cf_fill(c, sgwm);
}
int mb_cf_destroy(mb_cf* c) {
    if (!c) return MB_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    cudaFree(c->sigma_g_w_max);
    cudaFree(c->n_coll);
    delete c;
    return MB_OK;
}
int mb_cf_upload(mb_cf* c, const double* s) {
    MB_ARG(c && s, "NULL");
    MB_CUDA(cudaSetDevice(c->ctx->device));
    MB_CUDA(cudaMemcpyAsync(c->sigma_g_w_max, s, (size_t)c->n_cells * 8, cudaMemcpyHostToDevice, c->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return MB_OK;
}
int mb_cf_download(mb_cf* c, double* s, int64_t* n_coll, int64_t* n_perf, int64_t* n_eqw) {
    MB_ARG(c != nullptr, "NULL");
    MB_CUDA(cudaSetDevice(c->ctx->device));
    const size_t b = (size_t)c->n_cells * 8;
    if (s) MB_CUDA(cudaMemcpyAsync(s, c->sigma_g_w_max, b, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (n_coll) MB_CUDA(cudaMemcpyAsync(n_coll, c->n_coll, b, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (n_perf) MB_CUDA(cudaMemcpyAsync(n_perf, c->n_coll_performed, b, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (n_eqw) MB_CUDA(cudaMemcpyAsync(n_eqw, c->n_eq_w, b, cudaMemcpyDeviceToHost, c->ctx->stream));
    return mb_sync(c->ctx);
}

int mb_restore_particle_ordering(mb_ctx* ctx, mb_pv* pv) {  // particles.jl:1086-1137: index is the identity on the device
    MB_ARG(ctx && pv, "NULL");
    return MB_OK;
}

}  // extern "C"
