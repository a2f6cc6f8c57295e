// SurfProps on the device (properties/surface_props.jl:22-50 for the two walls of a 1-D grid and one species): the accumulators the
// convection kernels add to (update_surface_incident! / update_surface_reflected!, :77-131), surface_props_scale! (:144-160),
// clear_props! (:173-181), avg_props!(surf_avg, surf, n) (:202-222) and reduce_surf_props! (:232-252) -- all stream-ordered, no host
// synchronisation until mb_surf_download.  Layout: [wall][11] = np, flux_incident, flux_reflected, force[3], normal_pressure,
// shear_pressure[3], kinetic_energy_flux; wall 0 = left (x = 0), wall 1 = right (x = L).
#include "mb_common.cuh"

struct mb_surf {
    mb_ctx* ctx;
    double* d;  // 22 doubles
};

namespace mb {

int nccl_allreduce_sum_f64(mb_ctx* ctx, double* buf, size_t n);  // mb_exchange.cu

static __global__ void k_surf_scale(double* s, double factor) {  // surface_props_scale! (areas = 1); np is a count and is not scaled
    const int t = threadIdx.x;
    if (t < 22 && (t % 11) != 0) s[t] *= factor;
}
static __global__ void k_surf_avg(double* avg, const double* cur, double inv_n) {
    const int t = threadIdx.x;
    if (t < 22) avg[t] = avg[t] + cur[t] * inv_n;
}
static __global__ void k_surf_add(double* target, const double* src) {
    const int t = threadIdx.x;
    if (t < 22) target[t] += src[t];
}

double* surf_device_ptr(mb_surf* s) { return s->d; }
int surf_scale(mb_ctx* ctx, mb_surf* s, double factor) {
    k_surf_scale<<<1, 32, 0, ctx->stream>>>(s->d, factor);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_surf_create(mb_ctx* ctx, mb_surf** out) {
    MB_ARG(ctx && out, "NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    mb_surf* s = new mb_surf();
    s->ctx = ctx;
    s->d = nullptr;
    MB_CUDA(cudaMalloc(&s->d, 22 * sizeof(double)));
    MB_CUDA(cudaMemsetAsync(s->d, 0, 22 * sizeof(double), ctx->stream));
    *out = s;
    return MB_OK;
}
int mb_surf_destroy(mb_surf* s) {
    if (!s) return MB_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->d);
    delete s;
    return MB_OK;
}
int mb_surf_clear(mb_surf* s) {  // clear_props!(surf_props)
    MB_ARG(s != nullptr, "NULL");
    MB_CUDA(cudaSetDevice(s->ctx->device));
    MB_CUDA(cudaMemsetAsync(s->d, 0, 22 * sizeof(double), s->ctx->stream));
    return MB_OK;
}
int mb_surf_upload(mb_surf* s, const double* in22) {
    MB_ARG(s && in22, "NULL");
    MB_CUDA(cudaSetDevice(s->ctx->device));
    MB_CUDA(cudaMemcpyAsync(s->d, in22, 22 * sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return MB_OK;
}
int mb_surf_download(mb_surf* s, double* out22) {
    MB_ARG(s && out22, "NULL");
    MB_CUDA(cudaSetDevice(s->ctx->device));
    MB_CUDA(cudaMemcpyAsync(out22, s->d, 22 * sizeof(double), cudaMemcpyDeviceToHost, s->ctx->stream));
    return mb_sync(s->ctx);
}
int mb_surf_avg(mb_surf* avg, mb_surf* cur, int64_t n_avg_timesteps) {  // avg_props!(surf_props_avg, surf_props, n_avg_timesteps)
    MB_ARG(avg && cur && n_avg_timesteps > 0 && avg->ctx == cur->ctx, "surf_avg");
    MB_CUDA(cudaSetDevice(avg->ctx->device));
    k_surf_avg<<<1, 32, 0, avg->ctx->stream>>>(avg->d, cur->d, 1.0 / (double)n_avg_timesteps);
    MB_LAUNCH_CHECK(avg->ctx);
    return MB_OK;
}
/* reduce_surf_props!(surf_props_target, surf_props_chunks): target = sum of the chunks' accumulators (chunks of this process, in list
 * order like the reference's loop), then -- if the target's context belongs to a communicator of more than one rank and across_ranks
 * != 0 -- the sum over all ranks (ncclAllReduce of the 22 doubles). */
int mb_surf_reduce(mb_surf* target, mb_surf* const* chunks, int32_t n_chunks, int32_t across_ranks) {
    MB_ARG(target && (n_chunks == 0 || chunks) && n_chunks >= 0, "surf_reduce");
    mb_ctx* ctx = target->ctx;
    MB_CUDA(cudaSetDevice(ctx->device));
    for (int i = 0; i < n_chunks; i++) {
        MB_ARG(chunks[i] && chunks[i] != target, "surf_reduce: chunk");
        if (chunks[i]->ctx != ctx) {  // another chunk's stream: its accumulation must have finished
            MB_CUDA(cudaSetDevice(chunks[i]->ctx->device));
            MB_CUDA(cudaStreamSynchronize(chunks[i]->ctx->stream));
            MB_CUDA(cudaSetDevice(ctx->device));
        }
    }
    MB_CUDA(cudaMemsetAsync(target->d, 0, 22 * sizeof(double), ctx->stream));
    for (int i = 0; i < n_chunks; i++) {
        const double* src = chunks[i]->d;
        if (chunks[i]->ctx->device != ctx->device) {  // staged through the context's scratch
            double* tmp = (double*)ctx_scratch(ctx, 6, 22 * 8);
            if (!tmp) return MB_ERR_CUDA;
            MB_CUDA(cudaMemcpyAsync(tmp, src, 22 * 8, cudaMemcpyDefault, ctx->stream));
            src = tmp;
        }
        k_surf_add<<<1, 32, 0, ctx->stream>>>(target->d, src);
        MB_LAUNCH_CHECK(ctx);
    }
    if (across_ranks && ctx->nranks > 1) return nccl_allreduce_sum_f64(ctx, target->d, 22);
    return MB_OK;
}

}  // extern "C"
