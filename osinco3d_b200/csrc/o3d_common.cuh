// o3d_common.cuh -- shared device/host definitions of libo3d_b200 (sm_100a only).
//
// Arithmetic contract: every kernel evaluates the reference's expressions in the reference's
// order, in FP64, with FMA contraction disabled (nvcc -fmad=false), so stencil-only outputs
// are bit-identical to a gfortran -O3 x86-64 build of jojoledemago/osinco3d
// (src/derivation.f90, src/integration.f90 ...).
//
// Memory layout (DESIGN.md "Data layout in HBM"): every field lives in a PADDED, PITCHED box
//     (px, py, nz + 6)   px = pitch (multiple of 16 doubles = 128 B), py = ny + 6
// with the interior point (0,0,0) at element (GX, GH, GH).  Ghost cells (3 layers per side,
// the widest stencil radius) hold the boundary closure of src/derivation.f90 as DATA:
//     periodic  *_00  : f(-g) = f(n-g),   f(n-1+g) = f(g-1)
//     even      *p_11 : f(-g) = +f(g),    f(n-1+g) = +f(n-1-g)
//     odd       *i_11 : f(-g) = -f(g),    f(n-1+g) = -f(n-1-g)
// so every stencil kernel is the reference's INTERIOR formula (src/derivation.f90:43-47,
// :529-533) at every point, branch-free.  x - (-y) == x + y bitwise, so this reproduces the
// explicitly written boundary planes (e.g. :137-159) exactly; the even first derivative on a
// wall plane evaluates to a*(x-x) - b*(y-y) + c*(z-z) = +0.0, the literal the reference
// assigns (:87,:105).  z ghost planes at a rank boundary are filled by the NCCL halo exchange.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>

namespace o3d {

enum : int {
    BM_WRAP = 0,    // periodic                                   (der?_00)
    BM_MIRROR = 1,  // free-slip: f(1-k) = +-f(1+k)               (der?p_11 / der?i_11)
    BM_HALO = 2     // z only: ghost planes come from the neighbouring rank (multi-GPU)
};

constexpr int R = 3;    // ghost width = widest stencil radius (6th-order first derivative)
constexpr int GH = R;   // ghost rows / planes before the interior in y and z
constexpr int GX = 16;  // doubles before the interior in x: interior rows start 128 B aligned

struct Geom {
    int nx, ny, nz;      // interior extents of this rank's slab
    int px, py;          // pitch in doubles (multiple of 16), padded rows per plane (ny + 6)
    long long sy, sz;    // element strides of j and k: px, px * py
    int bx, by;          // BM_WRAP | BM_MIRROR
    int bz_lo, bz_hi;    // BM_WRAP | BM_MIRROR | BM_HALO
    int sim2d;           // derz_2dsim / derzz_2dsim: z derivatives are zero
    int gz0, gnz;        // first owned global plane, global nz (red-black colouring)
    // zr_hi > zr_lo: restrict a z-marching launch / an x-y ghost fill to the planes [zr_lo, zr_hi)
    // (the chunks of the pipelined host-pointer procedures, pipeline.cu); 0, 0 = whole slab
    int zr_lo, zr_hi;
};

inline int pitch_for(int nx) { return ((GX + nx + R + 15) / 16) * 16; }
// doubles in one field allocation / offset of interior (0,0,0) from the allocation start
inline long long field_elems(const Geom& g) { return g.sz * (long long)(g.nz + 2 * GH); }
inline long long interior_offset(const Geom& g) { return GX + g.sy * GH + g.sz * GH; }

// Map a line index q (may lie up to n-1 outside [0,n)) to a stored interior index under a
// closure; `refl` reports a mirror reflection (sign flips for odd parity).
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

// Ghost images of an interior point: the element-index offsets (0 = none) of the up to two ghost
// cells of its line that hold a copy of point p under the closure (low side / high side).
// Producers use it to write the ghost cells together with the interior value, so that no
// separate ghost-fill pass is needed after a fused kernel.
struct Img2 {
    int lo, hi;
};
__host__ __device__ __forceinline__ Img2 image_offsets(int p, int n, int mode_lo, int mode_hi) {
    Img2 r;
    r.lo = 0, r.hi = 0;
    if (mode_lo == BM_MIRROR) {
        if (p >= 1 && p <= R) r.lo = -2 * p;               // ghost -p = +-f(p)
    } else if (mode_lo == BM_WRAP) {
        if (p >= n - R) r.lo = -n;                         // ghost p-n = f(p)
    }
    if (mode_hi == BM_MIRROR) {
        if (p >= n - 1 - R && p <= n - 2) r.hi = 2 * (n - 1 - p);  // ghost 2(n-1)-p = +-f(p)
    } else if (mode_hi == BM_WRAP) {
        if (p < R) r.hi = n;                               // ghost n+p = f(p)
    }
    return r;
}

// store v at element m and at every ghost image of the point (faces, edges, corners)
__device__ __forceinline__ void store_images(double* __restrict__ q, long long m, double v,
                                             const Img2& ix, long long ylo, long long yhi,
                                             long long zlo, long long zhi) {
    const long long xo[3] = {0, ix.lo, ix.hi};
    const long long yo[3] = {0, ylo, yhi};
    const long long zo[3] = {0, zlo, zhi};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c && !zo[c]) continue;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b && !yo[b]) continue;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (a && !xo[a]) continue;
                if (a | b | c) q[m + xo[a] + yo[b] + zo[c]] = v;
            }
        }
    }
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
// (__host__ too: tests/cpu/stencil_rules_test.cu evaluates them on ghost-extended lines on the CPU
// against the vectors generated from the reference source)
__host__ __device__ __forceinline__ double d1_expr(double a, double b, double c, double m3, double m2,
                                          double m1, double p1, double p2, double p3) {
    return a * (p3 - m3) - b * (p2 - m2) + c * (p1 - m1);
}
__host__ __device__ __forceinline__ double d2_expr(double a, double b, double c, double m2, double m1,
                                          double f0, double p1, double p2) {
    return -(a * (m2 + p2)) + b * (m1 + p1) - c * f0;
}

// ---- point-wise expressions of the velocity kernels (__host__ too: tests/cpu/step_rules_test.cu
// evaluates them on the CPU against the vectors generated from the reference source) ----------
// the nine first derivatives d[c][axis] of (ux, uy, uz)
struct Grad {
    double d[3][3];
};

// Smagorinsky viscosity, src/les_turbulence.f90:70-88
__host__ __device__ __forceinline__ double smagorinsky(const Grad& G, double csd2) {
    const double s11 = G.d[0][0], s22 = G.d[1][1], s33 = G.d[2][2];
    const double s12 = 0.5 * (G.d[0][1] + G.d[1][0]);
    const double s13 = 0.5 * (G.d[0][2] + G.d[2][0]);
    const double s23 = 0.5 * (G.d[1][2] + G.d[2][1]);
    const double smag = sqrt(2.0 * (s11 * s11 + s22 * s22 + s33 * s33 +
                                    2.0 * (s12 * s12 + s13 * s13 + s23 * s23)));
    return csd2 * smag;
}

// Q criterion, src/differential_operators.f90:103-104
__host__ __device__ __forceinline__ double q_criterion_expr(const Grad& G) {
    return -(0.5 * (G.d[0][0] * G.d[0][0] + G.d[1][1] * G.d[1][1] + G.d[2][2] * G.d[2][2])) -
           G.d[0][1] * G.d[1][0] - G.d[0][2] * G.d[2][0] - G.d[1][2] * G.d[2][1];
}

// f_c = nu_eff (d2x + d2y + d2z) u_c - (ux dx + uy dy + uz dz) u_c, src/integration.f90:129-134
__host__ __device__ __forceinline__ double rhs_expr(double nu_eff, double lx, double ly, double lz,
                                                    double u0, double u1, double u2, double g0,
                                                    double g1, double g2) {
    return nu_eff * (lx + ly + lz) - (u0 * g0 + u1 * g1 + u2 * g2);
}

// u_c* = u_c + adu f1 + bdu f2 + cdu f3, src/integration.f90:132-134
__host__ __device__ __forceinline__ double predictor_expr(double uc, double adu, double f1,
                                                          double bdu, double f2, double cdu,
                                                          double f3) {
    return uc + adu * f1 + bdu * f2 + cdu * f3;
}

// divergence (src/differential_operators.f90:35) [/ dt -> Poisson right-hand side,
// src/integration.f90:239] and the projection correction u = u* - dt dp (src/integration.f90:304-306)
__host__ __device__ __forceinline__ double div_expr(double dfx, double dfy, double dfz, int divide,
                                                    double dt) {
    double v = dfx + dfy + dfz;
    if (divide) v = v / dt;
    return v;
}
__host__ __device__ __forceinline__ double corr_expr(double ustar, double dt, double dp) {
    return ustar - dt * dp;
}

// scalar transport, src/integration.f90:403-409 (effective diffusivity), :422-423 (right-hand
// side), :436/:450 (clip to [0,1]), :443-447 (redistribution of the clipped excess)
__host__ __device__ __forceinline__ double transeq_alpha(double resc, double nut, double sc,
                                                         int iles) {
    return iles ? (1.0 / resc + nut / sc) : (1.0 / resc);
}
__host__ __device__ __forceinline__ double transeq_rhs_expr(double alpha_eff, double dx2, double dy2,
                                                            double dz2, double u0, double u1,
                                                            double u2, double dx1, double dy1,
                                                            double dz1, double src) {
    return alpha_eff * (dx2 + dy2 + dz2) - (u0 * dx1 + u1 * dy1 + u2 * dz1) + src;
}
__host__ __device__ __forceinline__ double clip01(double x) { return fmax(0.0, fmin(1.0, x)); }
__host__ __device__ __forceinline__ double transeq_weight(double pc) { return fmin(pc, 1.0 - pc); }
__host__ __device__ __forceinline__ double transeq_redistribute(double pc, double excess,
                                                                double sum_w) {
    double wgt = transeq_weight(pc);
    wgt = wgt / sum_w;
    return clip01(pc + excess * wgt);
}

// ---- warp / block reductions ------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
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
