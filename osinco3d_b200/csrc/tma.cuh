// tma.cuh -- sm_100a TMA (cp.async.bulk.tensor) + mbarrier primitives as inline PTX, and the
// host-side tensor-map encoder (cuTensorMapEncodeTiled resolved through
// cudaGetDriverEntryPoint, so libo3d_b200.so does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace o3d {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async proxy (TMA unit)
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// one 3-D box: global (tensor map, element coordinates c0 fastest) -> shared, completion on mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- host ----------------------------------------------------------------------------------
// Tensor map of one padded field: dims (px, py, pz) doubles, box (bx, by, 1), no swizzle, zero
// fill out of bounds.  Returns 0 on success.
int make_field_tmap(CUtensorMap* out, const double* base, int px, int py, int pz, int box_x,
                    int box_y);

}  // namespace o3d
