// o3d_common.cuh -- shared device/host definitions of libo3d_b200 (sm_100a only).
//
// Arithmetic contract: every kernel evaluates the reference's expressions in the reference's
// order, in FP64, with FMA contraction disabled (nvcc -fmad=false), so stencil-only outputs
// are bit-identical to a gfortran -O3 x86-64 build of jojoledemago/osinco3d
// (src/derivation.f90, src/integration.f90 ...).  Ghost values carry their sign
// (x-(-y) == x+y bitwise), see DESIGN.md "Closures".
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace o3d {

// ---- boundary handling of one axis side -------------------------------------------------
enum : int {
    BM_WRAP = 0,    // periodic: f(p) = f(p -+ n)                 (der?_00)
    BM_MIRROR = 1,  // free-slip: f(1-k) = +-f(1+k)               (der?p_11 / der?i_11)
    BM_HALO = 2     // z only: planes -3..-1 / nz..nz+2 are stored ghost planes filled by the
                    // z-slab halo exchange (multi-GPU)
};

constexpr int R = 3;  // widest stencil radius (6th-order first derivative)

struct Dims {
    int nx, ny, nz;      // local extents (nz = planes owned by this rank)
    int bx, by;          // BM_WRAP | BM_MIRROR
    int bz_lo, bz_hi;    // BM_WRAP | BM_MIRROR | BM_HALO
    int sim2d;           // derz_2dsim / derzz_2dsim: z derivatives are zero
    int gz0, gnz;        // first owned global plane, global nz (red-black colouring)
};

// Map a line index q (may lie up to R outside [0,n)) to a stored index.
// `refl` reports a mirror reflection (sign flips for odd parity).
__host__ __device__ __forceinline__ int map_index(int q, int n, int mode_lo, int mode_hi,
                                                  bool& refl) {
    refl = false;
    if (q < 0) {
        if (mode_lo == BM_WRAP) return q + n;
        if (mode_lo == BM_MIRROR) {
            refl = true;
            return -q;
        }
        return q;  // BM_HALO: stored ghost plane
    }
    if (q >= n) {
        if (mode_hi == BM_WRAP) return q - n;
        if (mode_hi == BM_MIRROR) {
            refl = true;
            return 2 * (n - 1) - q;
        }
        return q;
    }
    return q;
}

// Coefficients exactly as the reference computes them.
struct Coef {
    double a1, b1, c1;  // src/derivation.f90:26-30   1,9,45 / (60 d)
    double a2, b2, c2;  // src/derivation.f90:517-521 1,16,30 / (12 d d)
};

inline Coef make_coef(double d) {
    Coef c;
    const double sixtyd = 60.0 * d;
    c.a1 = 1.0 / sixtyd;
    c.b1 = 9.0 / sixtyd;
    c.c1 = 45.0 / sixtyd;
    const double twelvedsq = 12.0 * d * d;
    c.a2 = 1.0 / twelvedsq;
    c.b2 = 16.0 / twelvedsq;
    c.c2 = 30.0 / twelvedsq;
    return c;
}

// interior expressions, src/derivation.f90:43-47 and :529-533 (evaluation order kept)
__device__ __forceinline__ double d1_expr(double a, double b, double c, double m3, double m2,
                                          double m1, double p1, double p2, double p3) {
    return a * (p3 - m3) - b * (p2 - m2) + c * (p1 - m1);
}
__device__ __forceinline__ double d2_expr(double a, double b, double c, double m2, double m1,
                                          double f0, double p1, double p2) {
    return -(a * (m2 + p2)) + b * (m1 + p1) - c * f0;
}

// ---- warp / block reductions ------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide max of a non-negative double; result valid in thread 0.  `red` = >= 32 doubles.
__device__ __forceinline__ double block_max(double v, double* red) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    v = warp_max(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (tid < 32) {
        r = (tid < (nthr + 31) / 32) ? red[tid] : 0.0;
        r = warp_max(r);
    }
    __syncthreads();
    return r;
}

// atomic max for NON-NEGATIVE doubles through their bit pattern (monotone for x >= +0)
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

}  // namespace o3d

#define O3D_CUDA_CHECK(call)                                                        \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) {                                                    \
            o3d::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                           cudaGetErrorString(_e));                                 \
            return O3D_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

namespace o3d {
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
}  // namespace o3d
