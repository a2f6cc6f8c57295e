// Entry points declared in include/merzbild_b200.h whose kernels are not written yet: they fail loudly.
#include "mb_common.cuh"

using namespace mb;

extern "C" {

int mb_swpm(mb_ctx*, mb_cf*, const mb_interaction*, mb_pv*, mb_pia*, int64_t, int64_t, int64_t, double, double, double, uint32_t, uint32_t) {
    set_error("mb_swpm: not implemented yet");
    return MB_ERR_UNSUPPORTED;
}
int mb_fp_linear(mb_ctx*, const mb_interaction*, double, mb_pv*, mb_pia*, int64_t, int64_t, int64_t, double, double, uint32_t, uint32_t) {
    set_error("mb_fp_linear: not implemented yet");
    return MB_ERR_UNSUPPORTED;
}
int mb_merge_octree_N2(mb_ctx*, const mb_octree_params*, mb_pv*, mb_pia*, int64_t, int64_t, int64_t, int64_t, int64_t, const mb_grid1d*, uint32_t,
                       uint32_t) {
    set_error("mb_merge_octree_N2: not implemented yet");
    return MB_ERR_UNSUPPORTED;
}
int mb_comm_unique_id(void*) {
    set_error("mb_comm_unique_id: not implemented yet");
    return MB_ERR_UNSUPPORTED;
}
int mb_comm_init(mb_ctx*, const void*, int, int) {
    set_error("mb_comm_init: not implemented yet");
    return MB_ERR_UNSUPPORTED;
}
int mb_exchange_slab(mb_ctx*, const mb_grid1d*, mb_pv*, mb_pia*, int64_t, int64_t*, int64_t*) {
    set_error("mb_exchange_slab: not implemented yet");
    return MB_ERR_UNSUPPORTED;
}

}
