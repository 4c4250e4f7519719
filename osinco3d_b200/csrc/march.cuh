// march.cuh -- the z-marching stencil engine shared by every fused stencil kernel (RHS +
// predictor, nu_t, divergence, projection correction, scalar transport, curl, Q, statistics).
//
// A CTA owns a 32 x 8 column of the (x,y) plane and marches over a chunk of z planes.
//   * Planes of the NF differentiated fields are staged in shared memory, WITH their 3-cell
//     x/y halos, by TMA (cp.async.bulk.tensor.3d, one 40 x 14 x 1 box per field and plane,
//     issued by one thread, completion on an mbarrier).  Because the fields are stored padded
//     with ghost cells that already hold the boundary closure (o3d_common.cuh), a box is a
//     plain rectangular read for interior and boundary tiles alike.
//   * The stages form a ring of 8 planes: k-3 .. k+3 are the z-stencil window of the plane
//     being computed, the eighth stage receives plane k+4 while plane k is computed, i.e. the
//     HBM latency of a plane is hidden behind a whole plane of compute without holding a single
//     register (the FP64 z-window of three fields would cost 42 registers per thread).
//   * One __syncthreads per plane retires the oldest stage; streamed operands of the epilogue
//     (AB history, u*, ...) are prefetched one plane ahead into registers.
// Box columns start at i0-4 (not i0-3) so that every box row is 10 full 32-byte sectors.
#pragma once
#include "kernels.h"
#include "tma.cuh"

namespace o3d {

constexpr int MTX = 32, MTY = 8, MNT = MTX * MTY;
constexpr int MXO = 4;                    // box starts 4 columns left of the tile
constexpr int MBX = MTX + 2 * MXO;        // 40 columns
constexpr int MBY = MTY + 2 * R;          // 14 rows
constexpr int MFIELD = MBX * MBY;         // doubles per staged field plane (4480 B = 35 x 128 B)
constexpr int MNST = 8;                   // ring stages

template <int NF>
struct alignas(64) MarchMaps {
    CUtensorMap m[NF];
};

struct MarchGeom {
    int nx, ny, nz;
    long long sy, sz;
    int zchunk;
    int sim2d;
};

template <int NF>
constexpr int march_smem_bytes() {
    return MNST * NF * MFIELD * 8 + MNST * 8;
}

// The staged window seen by one thread: p[m] points at this thread's cell of field 0 in plane
// k-3+m; field f is MFIELD doubles further.
template <int NF>
struct Ring {
    const double* p[7];
    __device__ __forceinline__ double c(int f) const { return p[3][f * MFIELD]; }
    __device__ __forceinline__ double x(int f, int d) const { return p[3][f * MFIELD + d]; }
    __device__ __forceinline__ double y(int f, int d) const { return p[3][f * MFIELD + d * MBX]; }
    __device__ __forceinline__ double z(int f, int d) const { return p[3 + d][f * MFIELD]; }
    // src/derivation.f90:43-47 / :529-533 along each axis
    __device__ __forceinline__ double d1x(int f, const Coef& q) const {
        return d1_expr(q.a1, q.b1, q.c1, x(f, -3), x(f, -2), x(f, -1), x(f, 1), x(f, 2), x(f, 3));
    }
    __device__ __forceinline__ double d1y(int f, const Coef& q) const {
        return d1_expr(q.a1, q.b1, q.c1, y(f, -3), y(f, -2), y(f, -1), y(f, 1), y(f, 2), y(f, 3));
    }
    __device__ __forceinline__ double d1z(int f, const Coef& q) const {
        return d1_expr(q.a1, q.b1, q.c1, z(f, -3), z(f, -2), z(f, -1), z(f, 1), z(f, 2), z(f, 3));
    }
    __device__ __forceinline__ double d2x(int f, const Coef& q) const {
        return d2_expr(q.a2, q.b2, q.c2, x(f, -2), x(f, -1), c(f), x(f, 1), x(f, 2));
    }
    __device__ __forceinline__ double d2y(int f, const Coef& q) const {
        return d2_expr(q.a2, q.b2, q.c2, y(f, -2), y(f, -1), c(f), y(f, 1), y(f, 2));
    }
    __device__ __forceinline__ double d2z(int f, const Coef& q) const {
        return d2_expr(q.a2, q.b2, q.c2, z(f, -2), z(f, -1), c(f), z(f, 1), z(f, 2));
    }
};

// Epilogue concept:
//   struct Pre;                                   streamed operands of one point
//   Pre  prefetch(long long m, bool ok) const;    issue their loads (m = element offset)
//   void apply(const Ring<NF>&, long long m, int i, int j, int k, const Pre&);
//   void finish(int tid, double* smem);           after the march (block reductions)
template <int NF, class Epi, int MINB>
__global__ void __launch_bounds__(MNT, MINB)
    march_kernel(const __grid_constant__ MarchMaps<NF> maps, const MarchGeom g, Epi epi) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + MNST * NF * MFIELD * 8);
    constexpr int STAGE = NF * MFIELD;                 // doubles
    constexpr uint32_t STAGE_BYTES = STAGE * 8;

    const int tid = threadIdx.x;
    const int tx = tid & (MTX - 1), ty = tid >> 5;
    const int i0 = blockIdx.x * MTX, j0 = blockIdx.y * MTY;
    const int i = i0 + tx, j = j0 + ty;
    const int kb = blockIdx.z * g.zchunk;
    const int ke = min(g.nz, kb + g.zchunk);
    const bool in_dom = (i < g.nx) && (j < g.ny);

    const uint32_t ring_s = smem_u32(ring);
    const uint32_t bars_s = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < MNST; ++s) mbar_init(bars_s + 8 * s, 1);
        fence_barrier_init();
    }
    __syncthreads();

    // box origin in tensor coordinates: element (GX + i0 - MXO, GH + j0 - R, GH + plane)
    const int cx = GX + i0 - MXO, cy = GH + j0 - R;
    auto issue = [&](int plane, int stage) {
        const uint32_t bar = bars_s + 8 * stage;
        mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int f = 0; f < NF; ++f)
            tma_load_3d(ring_s + (uint32_t)(stage * STAGE + f * MFIELD) * 8, &maps.m[f], bar, cx,
                        cy, GH + plane);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 7; ++s) issue(kb - R + s, s);
    }

    const long long m0 = (long long)j * g.sy + i;
    typename Epi::Pre cur = epi.prefetch(m0 + (long long)kb * g.sz, in_dom);
    const int cell = (ty + R) * MBX + tx + MXO;
#pragma unroll
    for (int s = 0; s < 6; ++s) mbar_wait(bars_s + 8 * s, 0);

    for (int k = kb; k < ke; ++k) {
        const int it = k - kb;
        // streamed operands of the next plane
        typename Epi::Pre nxt = epi.prefetch(m0 + (long long)(k + 1) * g.sz, in_dom && (k + 1 < ke));
        // plane k+3 has landed?
        mbar_wait(bars_s + 8 * ((it + 6) & 7), ((it + 6) >> 3) & 1);
        // every thread is done with plane k-1, so the stage of plane k-4 can be refilled
        __syncthreads();
        if (tid == 0 && k + 4 <= ke + 2) issue(k + 4, (it + 7) & 7);
        if (in_dom) {
            Ring<NF> r;
#pragma unroll
            for (int m = 0; m < 7; ++m) r.p[m] = ring + ((it + m) & 7) * STAGE + cell;
            epi.apply(r, m0 + (long long)k * g.sz, i, j, k, cur);
        }
        cur = nxt;
    }
    __syncthreads();
    epi.finish(tid, ring);
}

template <int NF, class Epi, int MINB>
int launch_march(cudaStream_t st, const Geom& g, const MarchMaps<NF>& maps, const Epi& epi) {
    static bool attr_set = false;
    auto kern = march_kernel<NF, Epi, MINB>;
    constexpr int smem = march_smem_bytes<NF>();
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return 1;
        attr_set = true;
    }
    MarchGeom mg;
    mg.nx = g.nx, mg.ny = g.ny, mg.nz = g.nz;
    mg.sy = g.sy, mg.sz = g.sz;
    mg.sim2d = g.sim2d;
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    mg.zchunk = pick_zchunk(gx * gy, g.nz);
    const int gz = (g.nz + mg.zchunk - 1) / mg.zchunk;
    kern<<<dim3(gx, gy, gz), dim3(MNT, 1, 1), smem, st>>>(maps, mg, epi);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
