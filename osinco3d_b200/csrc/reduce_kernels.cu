// reduce_kernels.cu -- device reductions over the interior of a padded field, for the per-step
// diagnostics the reference driver runs on full fields (function_stats src/functions.f90:27,
// minval/maxval prints src/IOfunctions.f90:322, compute_cfl src/utils.f90:178).
// Two-stage, fixed-shape reductions: results are deterministic run to run (no FP atomics).
#include "kernels.h"

namespace o3d {
namespace {

constexpr int RB = 256;      // threads per block
constexpr int RMAXB = 1184;  // 148 SMs x 8

__device__ __forceinline__ double red_op(double a, double b, int op) {
    if (op == RED_MIN) return fmin(a, b);
    if (op == RED_SUM) return a + b;
    return fmax(a, b);
}
__device__ __forceinline__ double red_init(int op) {
    if (op == RED_MIN) return 1.7976931348623157e308;
    if (op == RED_MAX) return -1.7976931348623157e308;
    return 0.0;
}

__device__ __forceinline__ double block_reduce(double v, int op, double* red) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op(v, __shfl_xor_sync(0xffffffffu, v, o), op);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = red_init(op);
    if (tid < 32) {
        r = (tid < RB / 32) ? red[tid] : red_init(op);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = red_op(r, __shfl_xor_sync(0xffffffffu, r, o), op);
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(RB) reduce_stage1(const Geom g, const double* __restrict__ f,
                                                     int op, double* partial) {
    __shared__ double red[32];
    double v = red_init(op);
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const long long base = (long long)k * g.sz + (long long)j * g.sy;
        for (int i = threadIdx.x; i < g.nx; i += RB) {
            double x = __ldg(f + base + i);
            if (op == RED_ABSMAX) x = fabs(x);
            v = red_op(v, x, op);
        }
    }
    v = block_reduce(v, op, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}

__global__ void __launch_bounds__(RB) reduce_stage2(const double* partial, int nparts, int op,
                                                     double* out) {
    __shared__ double red[32];
    double v = red_init(op);
    for (int m = threadIdx.x; m < nparts; m += RB) v = red_op(v, partial[m], op);
    v = block_reduce(v, op, red);
    if (threadIdx.x == 0) *out = v;
}

// ncomp independent sums over partial[c*nparts + b]
__global__ void __launch_bounds__(RB) sum_partials_kernel(const double* partial, int nparts,
                                                           double* out) {
    __shared__ double red[32];
    const double* p = partial + (long long)blockIdx.x * nparts;
    double v = 0.0;
    for (int m = threadIdx.x; m < nparts; m += RB) v += p[m];
    v = block_reduce(v, RED_SUM, red);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
}

// function_stats: min, max (first occurrence in i-fastest order), sum
struct StatAcc {
    double mn, mx, sum;
    long long imx;
};
__device__ __forceinline__ void stat_merge(StatAcc& a, const StatAcc& b) {
    a.mn = fmin(a.mn, b.mn);
    a.sum += b.sum;
    if (b.mx > a.mx || (b.mx == a.mx && b.imx < a.imx)) {
        a.mx = b.mx;
        a.imx = b.imx;
    }
}
__device__ __forceinline__ StatAcc stat_shfl(const StatAcc& a, int o) {
    StatAcc b;
    b.mn = __shfl_xor_sync(0xffffffffu, a.mn, o);
    b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
    b.sum = __shfl_xor_sync(0xffffffffu, a.sum, o);
    b.imx = __shfl_xor_sync(0xffffffffu, a.imx, o);
    return b;
}
__device__ __forceinline__ StatAcc stat_block(StatAcc v, StatAcc* red) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const StatAcc b = stat_shfl(v, o);
        stat_merge(v, b);
    }
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0)
        for (int q = 1; q < RB / 32; ++q) stat_merge(v, red[q]);
    __syncthreads();
    return v;
}
__device__ __forceinline__ StatAcc stat_init() {
    StatAcc a;
    a.mn = 1.7976931348623157e308;   // huge(), src/functions.f90:36
    a.mx = -1.7976931348623157e308;  // -huge(), :37
    a.sum = 0.0;
    a.imx = 0x7fffffffffffffffLL;
    return a;
}

__global__ void __launch_bounds__(RB) fstats_stage1(const Geom g, const double* __restrict__ f,
                                                     double* partial) {
    __shared__ StatAcc red[RB / 32];
    StatAcc v = stat_init();
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const long long base = (long long)k * g.sz + (long long)j * g.sy;
        for (int i = threadIdx.x; i < g.nx; i += RB) {
            const double x = __ldg(f + base + i);
            const long long lin = row * g.nx + i;  // position in the reference's array order
            v.mn = fmin(v.mn, x);
            v.sum += x;
            if (x > v.mx) {  // strict: first occurrence wins within a thread (lin ascending)
                v.mx = x;
                v.imx = lin;
            }
        }
    }
    v = stat_block(v, red);
    if (threadIdx.x == 0) {
        double* p = partial + 4ll * blockIdx.x;
        p[0] = v.mn, p[1] = v.mx, p[2] = v.sum, p[3] = __longlong_as_double(v.imx);
    }
}

__global__ void __launch_bounds__(RB) fstats_stage2(const double* partial, int nparts, int nx,
                                                     int ny, int nz, double* out6) {
    __shared__ StatAcc red[RB / 32];
    StatAcc v = stat_init();
    for (int b = threadIdx.x; b < nparts; b += RB) {
        StatAcc t;
        t.mn = partial[4ll * b], t.mx = partial[4ll * b + 1], t.sum = partial[4ll * b + 2];
        t.imx = __double_as_longlong(partial[4ll * b + 3]);
        stat_merge(v, t);
    }
    v = stat_block(v, red);
    if (threadIdx.x == 0) {
        out6[0] = v.mn;
        out6[1] = v.mx;
        out6[2] = v.sum / (double)((long long)nx * ny * nz);  // src/functions.f90:61
        long long m = v.imx;
        if (m == 0x7fffffffffffffffLL) m = 0;  // all values <= -huge: reference keeps (1,1,1)
        out6[3] = (double)(m % nx + 1);
        out6[4] = (double)((m / nx) % ny + 1);
        out6[5] = (double)(m / ((long long)nx * ny) + 1);
    }
}

// ---- calculate_residuals, src/utils.f90:93-160 ---------------------------------------------
struct ResAcc {
    double s[3], mx[3];
    long long im[3];
};
__device__ __forceinline__ void res_merge(ResAcc& a, const ResAcc& b) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a.s[c] += b.s[c];
        // the reference's second loop (:125-145) keeps the LAST point equal to the maximum
        if (b.mx[c] > a.mx[c] || (b.mx[c] == a.mx[c] && b.im[c] > a.im[c]))
            a.mx[c] = b.mx[c], a.im[c] = b.im[c];
    }
}
__device__ __forceinline__ ResAcc res_block(ResAcc v, ResAcc* red) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ResAcc b;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            b.s[c] = __shfl_xor_sync(0xffffffffu, v.s[c], o);
            b.mx[c] = __shfl_xor_sync(0xffffffffu, v.mx[c], o);
            b.im[c] = __shfl_xor_sync(0xffffffffu, v.im[c], o);
        }
        // xor butterflies must merge symmetrically to stay deterministic: order by lane
        if (tid & o) {
            ResAcc t = b;
            res_merge(t, v);
            v = t;
        } else {
            res_merge(v, b);
        }
    }
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0)
        for (int q = 1; q < RB / 32; ++q) res_merge(v, red[q]);
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(RB)
    resid_stage1(const Geom g, const double* __restrict__ u0, const double* __restrict__ u1,
                 const double* __restrict__ u2, const double* __restrict__ o0,
                 const double* __restrict__ o1, const double* __restrict__ o2, double two_dt,
                 double* partial) {
    __shared__ ResAcc red[RB / 32];
    ResAcc v;
#pragma unroll
    for (int c = 0; c < 3; ++c) v.s[c] = 0.0, v.mx[c] = 0.0, v.im[c] = -1;  // linf_? = 0.d0, :104
    const double* un[3] = {u0, u1, u2};
    const double* uo[3] = {o0, o1, o2};
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const int gk = g.gz0 + k;
        if (j < 1 || j > g.ny - 2 || gk < 1 || gk > g.gnz - 2) continue;  // do k = 2, nz-1 ...
        const long long base = (long long)k * g.sz + (long long)j * g.sy;
        for (int i = 1 + threadIdx.x; i < g.nx - 1; i += RB) {
            const long long lin = ((long long)gk * g.ny + j) * g.nx + i;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double a = fabs(__ldg(uo[c] + base + i) - __ldg(un[c] + base + i)) / two_dt;
                v.s[c] += a * a;
                if (a >= v.mx[c]) v.mx[c] = a, v.im[c] = lin;  // lin ascending within a thread
            }
        }
    }
    v = res_block(v, red);
    if (threadIdx.x == 0) {
        double* p = partial + 9ll * blockIdx.x;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            p[c] = v.s[c], p[3 + c] = v.mx[c], p[6 + c] = (double)v.im[c];
    }
}

__global__ void __launch_bounds__(RB) resid_stage2(const double* partial, int nparts,
                                                    double* out9) {
    __shared__ ResAcc red[RB / 32];
    ResAcc v;
#pragma unroll
    for (int c = 0; c < 3; ++c) v.s[c] = 0.0, v.mx[c] = 0.0, v.im[c] = -1;
    for (int b = threadIdx.x; b < nparts; b += RB) {
        ResAcc t;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            t.s[c] = partial[9ll * b + c], t.mx[c] = partial[9ll * b + 3 + c],
            t.im[c] = (long long)partial[9ll * b + 6 + c];
        res_merge(v, t);
    }
    v = res_block(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out9[c] = v.s[c], out9[3 + c] = v.mx[c], out9[6 + c] = (double)v.im[c];
    }
}

}  // namespace

int launch_residuals(cudaStream_t st, const Geom& g, const double* const* unew,
                     const double* const* uold, double two_dt, double* partial, double* out9) {
    int nb = reduce_blocks(g);
    if (nb > 296) nb = 296;
    resid_stage1<<<nb, RB, 0, st>>>(g, unew[0], unew[1], unew[2], uold[0], uold[1], uold[2],
                                    two_dt, partial);
    resid_stage2<<<1, RB, 0, st>>>(partial, nb, out9);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int reduce_blocks(const Geom& g) {
    long long b = (long long)g.ny * g.nz;
    if (b > RMAXB) b = RMAXB;
    if (b < 1) b = 1;
    return (int)b;
}

int launch_reduce(cudaStream_t st, const Geom& g, const double* f, int op, double* partial,
                  double* out) {
    const int nb = reduce_blocks(g);
    reduce_stage1<<<nb, RB, 0, st>>>(g, f, op, partial);
    reduce_stage2<<<1, RB, 0, st>>>(partial, nb, op == RED_ABSMAX ? RED_MAX : op, out);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_function_stats(cudaStream_t st, const Geom& g, const double* f, double* partial,
                          double* out6) {
    int nb = reduce_blocks(g);
    if (nb > RMAXB / 4) nb = RMAXB / 4;  // partial holds 4 doubles per block
    fstats_stage1<<<nb, RB, 0, st>>>(g, f, partial);
    fstats_stage2<<<1, RB, 0, st>>>(partial, nb, g.nx, g.ny, g.nz, out6);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_sum_partials(cudaStream_t st, const double* partial, int nparts, int ncomp,
                        double* out) {
    sum_partials_kernel<<<ncomp, RB, 0, st>>>(partial, nparts, out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
