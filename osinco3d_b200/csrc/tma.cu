// tma.cu -- host-side TMA tensor-map construction for padded fields.
#include "tma.cuh"

#include <cstdlib>

#include "o3d_common.cuh"

namespace o3d {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
                cudaSuccess ||
            q != cudaDriverEntryPointSuccess || !p) {
            (void)cudaGetLastError();
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_field_tmap(CUtensorMap* out, const double* base, int px, int py, int pz, int box_x,
                    int box_y) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return 1;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)px, (cuuint64_t)py, (cuuint64_t)pz};
    const cuuint64_t strides[2] = {(cuuint64_t)px * 8ull, (cuuint64_t)px * (cuuint64_t)py * 8ull};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // FLOAT64 is not a TMA element type on every toolkit: move 8-byte elements as INT64
    // L2 promotion = the granularity at which a TMA miss is filled from DRAM.  Tuning override:
    // O3D_TMA_L2PROMO = 0 | 64 | 128 | 256
    static int promo = -1;
    if (promo < 0) {
        const char* e = getenv("O3D_TMA_L2PROMO");
        promo = e ? atoi(e) : 128;
    }
    const CUtensorMapL2promotion pr = promo == 0     ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                      : promo == 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                      : promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                     : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_INT64, 3, const_cast<double*>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (px=%d py=%d pz=%d box=%dx%d)",
                  (int)r, px, py, pz, box_x, box_y);
        return 1;
    }
    return 0;
}

}  // namespace o3d
