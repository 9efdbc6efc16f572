// Launcher of the device-side collation kernel (collate.cuh).
#include "collate.cuh"
#include "common.cuh"

namespace pamnet {

__global__ void __launch_bounds__(128) collate_kernel(const CollateArgs a) {
    pdl_wait();
    pdl_trigger();
    collate_body(a, (int)blockIdx.x);
}

int collate(const CollateArgs& a, cudaStream_t st) {
    PAMNET_CHECK_ARG(a.n_ids > 0 && a.n_ids < (1ll << 31), "collate: n_ids=%lld", (long long)a.n_ids);
    prof_begin(KC_MISC, 0.0, st);
    launch_pdl(collate_kernel, dim3((unsigned)a.n_ids), dim3(128), 0, st, a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

}  // namespace pamnet
