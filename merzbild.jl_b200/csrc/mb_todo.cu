// Entry points declared in include/merzbild_b200.h whose kernels are not written yet: they fail loudly.
#include "mb_common.cuh"

using namespace mb;

extern "C" {

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
