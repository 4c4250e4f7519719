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
    // z range of this launch.  zmode 0/1: chunks of [zlo, zhi) (1 = interior of a split launch);
    // zmode 2: the two boundary chunks [0, zedge) and [nz - zedge, nz) of a split launch, which
    // run after the z-slab halo exchange that the interior launch overlaps
    int zmode, zlo, zhi, zedge;
    int sim2d;
    int bx, by, bz_lo, bz_hi;  // closures, for epilogues that also write ghost images
    // gated launches (NALT = 1): run only if the SOR control block says the solve has finished,
    // and read z-field 0 from the ping-pong buffer the last pass wrote (alternate tensor map when
    // the number of passes was odd).  Lets the projection correction be queued BEHIND a batch of
    // SOR passes without a host round trip in between.
    const SorCtrl* gate;
};

// NFZ fields need the 7-plane z window (ring of 7+P stages), NFC fields only the plane being
// computed, with its x/y halo (ring of 1+P stages), NFS "stream" fields only the point itself
// (32 x 8 boxes without halo, ring of 1+P stages: the streamed operands of an epilogue, fetched
// by TMA P planes ahead instead of through registers).  P = prefetch distance in planes.
// Fields are ordered z-fields, c-fields, s-fields in MarchMaps.
constexpr int MSFIELD = MTX * MTY;  // doubles per staged stream-field plane (2 KB)
template <int NFZ, int NFC, int P, int NFS = 0, int NFW = 0>
constexpr int march_smem_bytes() {
    return (NFZ * (7 + P) + NFC * (1 + P)) * MFIELD * 8 + NFS * (1 + P) * MSFIELD * 8 +
           NFW * (7 + P) * MSFIELD * 8 + (P + 1) * 8;
}

// The staged data seen by one thread: p[m] points at this thread's cell of z-field 0 in plane
// k-3+m (z-field f is MFIELD doubles further); q points at its cell of c-field 0 in plane k; s at
// its value of s-field 0 in plane k (s-field f is MSFIELD doubles further).
template <int NFZ, int NFC = 0>
struct Ring {
    const double* p[7];
    const double* q;
    const double* s;
    __device__ __forceinline__ double c(int f) const { return p[3][f * MFIELD]; }
    __device__ __forceinline__ double x(int f, int d) const { return p[3][f * MFIELD + d]; }
    __device__ __forceinline__ double y(int f, int d) const { return p[3][f * MFIELD + d * MBX]; }
    __device__ __forceinline__ double z(int f, int d) const { return p[3 + d][f * MFIELD]; }
    __device__ __forceinline__ double cx(int f, int d) const { return q[f * MFIELD + d]; }
    __device__ __forceinline__ double cy(int f, int d) const { return q[f * MFIELD + d * MBX]; }
    __device__ __forceinline__ double st(int f) const { return s[f * MSFIELD]; }
    // src/derivation.f90:43-47 / :529-533 along each axis
    __device__ __forceinline__ double d1x(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, x(f, -3), x(f, -2), x(f, -1), x(f, 1), x(f, 2), x(f, 3));
    }
    __device__ __forceinline__ double d1y(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, y(f, -3), y(f, -2), y(f, -1), y(f, 1), y(f, 2), y(f, 3));
    }
    __device__ __forceinline__ double d1z(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, z(f, -3), z(f, -2), z(f, -1), z(f, 1), z(f, 2), z(f, 3));
    }
    __device__ __forceinline__ double d2x(int f, const Coef& k) const {
        return d2_expr(k.a2, k.b2, k.c2, x(f, -2), x(f, -1), c(f), x(f, 1), x(f, 2));
    }
    __device__ __forceinline__ double d2y(int f, const Coef& k) const {
        return d2_expr(k.a2, k.b2, k.c2, y(f, -2), y(f, -1), c(f), y(f, 1), y(f, 2));
    }
    __device__ __forceinline__ double d2z(int f, const Coef& k) const {
        return d2_expr(k.a2, k.b2, k.c2, z(f, -2), z(f, -1), c(f), z(f, 1), z(f, 2));
    }
    // centre-only fields
    __device__ __forceinline__ double c_d1x(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, cx(f, -3), cx(f, -2), cx(f, -1), cx(f, 1), cx(f, 2),
                       cx(f, 3));
    }
    __device__ __forceinline__ double c_d1y(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, cy(f, -3), cy(f, -2), cy(f, -1), cy(f, 1), cy(f, 2),
                       cy(f, 3));
    }
};

// Split-ring view (NFW > 0): a field that is differentiated in all three directions is staged
// twice -- plane k WITH its x/y halo (c-ring, 1+P stages of 40 x 14 boxes) and planes k-3 .. k+3
// WITHOUT halo (w-ring, 7+P stages of 32 x 8 boxes) -- because the z stencil only ever reads the
// thread's own column; a field that is only differentiated in z needs the w-ring alone.  x / y /
// c index the c-fields, z the w-fields (RHS: the same three fields in both rings).  Per field the
// ring costs (7+P) * 2 KB + (1+P) * 4.4 KB instead of (7+P) * 4.4 KB, which buys prefetch depth
// (bytes in flight) inside the same shared-memory budget; the second read of a plane hits L2.
struct RingCW {
    const double* w[7];  // this thread's value in planes k-3 .. k+3 (field f is MSFIELD further)
    const double* q;     // its cell in the halo'd plane k (field f is MFIELD further)
    __device__ __forceinline__ double c(int f) const { return q[f * MFIELD]; }
    __device__ __forceinline__ double x(int f, int d) const { return q[f * MFIELD + d]; }
    __device__ __forceinline__ double y(int f, int d) const { return q[f * MFIELD + d * MBX]; }
    __device__ __forceinline__ double z(int f, int d) const { return w[3 + d][f * MSFIELD]; }
    __device__ __forceinline__ double d1x(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, x(f, -3), x(f, -2), x(f, -1), x(f, 1), x(f, 2), x(f, 3));
    }
    __device__ __forceinline__ double d1y(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, y(f, -3), y(f, -2), y(f, -1), y(f, 1), y(f, 2), y(f, 3));
    }
    __device__ __forceinline__ double d1z(int f, const Coef& k) const {
        return d1_expr(k.a1, k.b1, k.c1, z(f, -3), z(f, -2), z(f, -1), z(f, 1), z(f, 2), z(f, 3));
    }
    __device__ __forceinline__ double d2x(int f, const Coef& k) const {
        return d2_expr(k.a2, k.b2, k.c2, x(f, -2), x(f, -1), c(f), x(f, 1), x(f, 2));
    }
    __device__ __forceinline__ double d2y(int f, const Coef& k) const {
        return d2_expr(k.a2, k.b2, k.c2, y(f, -2), y(f, -1), c(f), y(f, 1), y(f, 2));
    }
    __device__ __forceinline__ double d2z(int f, const Coef& k) const {
        return d2_expr(k.a2, k.b2, k.c2, z(f, -2), z(f, -1), c(f), z(f, 1), z(f, 2));
    }
    // names of Ring<NFZ, NFC>'s centre-only accessors (c-field f)
    __device__ __forceinline__ double cx(int f, int d) const { return x(f, d); }
    __device__ __forceinline__ double c_d1x(int f, const Coef& k) const { return d1x(f, k); }
    __device__ __forceinline__ double c_d1y(int f, const Coef& k) const { return d1y(f, k); }
};

// Epilogue concept:
//   static constexpr int STREAMS;                 values moved per point (reads + writes): sizes
//                                                 the z chunks (pick_zchunk)
//   void setup(const MarchGeom&, int i, int j);   per-thread constants, before the march
//   struct Pre;                                   streamed operands of one point (register path)
//   Pre  prefetch(long long m, bool ok) const;    issue their loads (m = element offset)
//   void apply(const Ring<NFZ,NFC>&, long long m, int i, int j, int k, const Pre&);
//   void finish(int tid, double* smem);           after the march (block reductions)
// UNR > 1: the plane loop is unrolled UNR times, UNR a common multiple of the ring lengths, so
// that every ring position is a compile-time constant inside the body: shared-memory operands
// become [base + immediate] and the per-plane ring-pointer arithmetic disappears.
template <int NFZ, int NFC, int P, class Epi, int MINB, int NFS = 0, int NALT = 0, int UNR = 1,
          int NFW = 0>
__global__ void __launch_bounds__(MNT, MINB)
    march_kernel(const __grid_constant__ MarchMaps<NFZ + NFC + NFS + NALT + NFW> maps,
                 const MarchGeom g, Epi epi) {
    static_assert(NFW == 0 || (NFZ == 0 && NFS == 0 && NALT == 0 && UNR == 1),
                  "split-ring mode: z windows live in the w-ring only");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const CUtensorMap* map0 = &maps.m[0];
    if (NALT) {
        if (!*((volatile const int*)&g.gate->done)) return;  // solve not finished: do nothing
        if (*((volatile const int*)&g.gate->iter) & 1) map0 = &maps.m[NFZ + NFC + NFS];
    }
    constexpr int NZS = 7 + P, NCS = 1 + P, NB = P + 1;
    constexpr int ZSTAGE = NFZ * MFIELD, CSTAGE = NFC * MFIELD, SSTAGE = NFS * MSFIELD;  // doubles
    double* zring = reinterpret_cast<double*>(smem_raw);
    double* cring = zring + NZS * ZSTAGE;
    double* sring = cring + NCS * CSTAGE;
    constexpr int WSTAGE = NFW * MSFIELD;
    double* wring = sring + NCS * SSTAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + NZS * WSTAGE);
    constexpr uint32_t PLANE_BYTES = MFIELD * 8, SPLANE_BYTES = MSFIELD * 8;

    const int tid = threadIdx.x;
    const int tx = tid & (MTX - 1), ty = tid >> 5;
    const int i0 = blockIdx.x * MTX, j0 = blockIdx.y * MTY;
    const int i = i0 + tx, j = j0 + ty;
    int kb, ke;
    if (g.zmode == 2) {
        kb = (blockIdx.z == 0) ? 0 : g.nz - g.zedge;
        ke = kb + g.zedge;
    } else {
        kb = g.zlo + blockIdx.z * g.zchunk;
        ke = min(g.zhi, kb + g.zchunk);
    }
    const bool in_dom = (i < g.nx) && (j < g.ny);

    const uint32_t zring_s = smem_u32(zring), cring_s = smem_u32(cring), sring_s = smem_u32(sring);
    const uint32_t wring_s = smem_u32(wring);
    const uint32_t bars_s = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NB; ++s) mbar_init(bars_s + 8 * s, 1);
        fence_barrier_init();
    }
    __syncthreads();

    // box origin in tensor coordinates: element (GX + i0 - MXO, GH + j0 - R, GH + plane)
    const int cx = GX + i0 - MXO, cy = GH + j0 - R;
    // z-plane `plane` lives in z-stage (plane - (kb-3)) mod NZS, c/s-plane in (plane - kb) mod NCS
    auto issue_z = [&](int plane, uint32_t bar) {
        const unsigned st = (unsigned)(plane - (kb - R)) % NZS;
#pragma unroll
        for (int f = 0; f < NFZ; ++f)
            tma_load_3d(zring_s + (uint32_t)(st * ZSTAGE + f * MFIELD) * 8,
                        (NALT && f == 0) ? map0 : &maps.m[f], bar, cx, cy, GH + plane);
    };
    // w-plane `plane` (tile only) lives in w-stage (plane - (kb-3)) mod NZS; its tensor maps follow
    // the c-field maps
    auto issue_w = [&](int plane, uint32_t bar) {
        const unsigned st = (unsigned)(plane - (kb - R)) % NZS;
#pragma unroll
        for (int f = 0; f < NFW; ++f)
            tma_load_3d(wring_s + (uint32_t)(st * WSTAGE + f * MSFIELD) * 8,
                        &maps.m[NFZ + NFC + NFS + NALT + f], bar, GX + i0, GH + j0, GH + plane);
    };
    auto issue_c = [&](int plane, uint32_t bar) {
        const unsigned st = (unsigned)(plane - kb) % NCS;
#pragma unroll
        for (int f = 0; f < NFC; ++f)
            tma_load_3d(cring_s + (uint32_t)(st * CSTAGE + f * MFIELD) * 8, &maps.m[NFZ + f], bar,
                        cx, cy, GH + plane);
#pragma unroll
        for (int f = 0; f < NFS; ++f)
            tma_load_3d(sring_s + (uint32_t)(st * SSTAGE + f * MSFIELD) * 8,
                        &maps.m[NFZ + NFC + f], bar, GX + i0, GH + j0, GH + plane);
    };
    // "need group" n = what iteration n waits for: z-plane kb+n+3 and c/s-plane kb+n (group 0 also
    // carries z-planes kb-3 .. kb+2); its barrier is n mod NB
    const int niter = ke - kb;
    auto issue_group = [&](int n) {
        const uint32_t bar = bars_s + 8 * ((unsigned)n % NB);
        const int nz_planes = (n == 0) ? 7 : 1;
        mbar_expect_tx(bar, (uint32_t)(nz_planes * NFZ + NFC) * PLANE_BYTES +
                                (uint32_t)(NFS + nz_planes * NFW) * SPLANE_BYTES);
        if (n == 0) {
#pragma unroll
            for (int s = 0; s < 6; ++s) {
                issue_z(kb - R + s, bar);
                issue_w(kb - R + s, bar);
            }
        }
        issue_z(kb + n + R, bar);
        issue_w(kb + n + R, bar);
        issue_c(kb + n, bar);
    };
    if (tid == 0) {
        for (int n = 0; n < P && n < niter; ++n) issue_group(n);
    }

    const long long m0 = (long long)j * g.sy + i;
    epi.setup(g, i, j);
    typename Epi::Pre cur = epi.prefetch(m0 + (long long)kb * g.sz, in_dom);
    const int cell = (ty + R) * MBX + tx + MXO;

    long long m = m0 + (long long)kb * g.sz;
    if constexpr (UNR > 1) {
        static_assert(UNR % NZS == 0 && UNR % NCS == 0 && UNR % NB == 0,
                      "UNR must be a common multiple of the ring lengths");
        // barrier phase of the first round of an unrolled block: flips from block to block when a
        // block holds an odd number of barrier rounds
        uint32_t bp = 0;
        for (int k0 = kb; k0 < ke; k0 += UNR, bp ^= (uint32_t)((UNR / NB) & 1)) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int k = k0 + u;
                if (k < ke) {  // uniform over the CTA
                    const int it = k - kb;
                    typename Epi::Pre nxt = epi.prefetch(m + g.sz, in_dom && (k + 1 < ke));
                    __syncthreads();
                    if (tid == 0 && it + P < niter) issue_group(it + P);
                    mbar_wait(bars_s + 8 * (u % NB), bp ^ (uint32_t)((u / NB) & 1));
                    if (in_dom) {
                        Ring<NFZ, NFC> r;
#pragma unroll
                        for (int w = 0; w < 7; ++w) r.p[w] = zring + ((u + w) % NZS) * ZSTAGE + cell;
                        r.q = cring + (u % NCS) * CSTAGE + cell;
                        r.s = sring + (u % NCS) * SSTAGE + tid;
                        epi.apply(r, m, i, j, k, cur);
                    }
                    cur = nxt;
                    m += g.sz;
                }
            }
        }
    } else {
    // ring positions of the plane being computed, advanced incrementally (no modulo in the loop):
    // zb = stage of plane k-3, cb = stage of plane k, bi / bpar = barrier index and phase
    int zb = 0, cb = 0, bi = 0;
    uint32_t bpar = 0;
    for (int k = kb; k < ke; ++k) {
        const int it = k - kb;
        // streamed operands of the next plane (register path)
        typename Epi::Pre nxt = epi.prefetch(m + g.sz, in_dom && (k + 1 < ke));
        // every thread is done with plane k-1: its stages can be refilled with group it+P
        __syncthreads();
        if (tid == 0 && it + P < niter) issue_group(it + P);
        // group `it` has landed?
        mbar_wait(bars_s + 8 * bi, bpar);
        if (in_dom) {
            if constexpr (NFW > 0) {
                RingCW r;
#pragma unroll
                for (int w = 0; w < 7; ++w) {
                    int st = zb + w;
                    st = (st >= NZS) ? st - NZS : st;
                    r.w[w] = wring + st * WSTAGE + tid;
                }
                r.q = cring + cb * CSTAGE + cell;
                epi.apply(r, m, i, j, k, cur);
            } else {
                Ring<NFZ, NFC> r;
#pragma unroll
                for (int w = 0; w < 7; ++w) {
                    int st = zb + w;
                    st = (st >= NZS) ? st - NZS : st;
                    r.p[w] = zring + st * ZSTAGE + cell;
                }
                r.q = cring + cb * CSTAGE + cell;
                r.s = sring + cb * SSTAGE + tid;
                epi.apply(r, m, i, j, k, cur);
            }
        }
        cur = nxt;
        m += g.sz;
        zb = (zb + 1 == NZS) ? 0 : zb + 1;
        cb = (cb + 1 == NCS) ? 0 : cb + 1;
        if (++bi == NB) bi = 0, bpar ^= 1u;
    }
    }
    __syncthreads();
    epi.finish(tid, zring);
}

// zmode: ZFULL whole slab | ZINTERIOR planes [zedge, nz - zedge) | ZBOUNDARY the two end chunks
enum { ZFULL = 0, ZINTERIOR = 1, ZBOUNDARY = 2 };

template <int NFZ, int NFC, int P, class Epi, int MINB, int NFS = 0, int NALT = 0, int UNR = 1,
          int NFW = 0>
int launch_march(cudaStream_t st, const Geom& g,
                 const MarchMaps<NFZ + NFC + NFS + NALT + NFW>& maps, const Epi& epi,
                 int zmode = ZFULL, int zedge = 0, const SorCtrl* gate = nullptr) {
    static bool attr_set = false;
    auto kern = march_kernel<NFZ, NFC, P, Epi, MINB, NFS, NALT, UNR, NFW>;
    constexpr int smem = march_smem_bytes<NFZ, NFC, P, NFS, NFW>();
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
    mg.bx = g.bx, mg.by = g.by, mg.bz_lo = g.bz_lo, mg.bz_hi = g.bz_hi;
    mg.gate = gate;
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    mg.zmode = zmode, mg.zedge = zedge;
    mg.zlo = (zmode == ZINTERIOR) ? zedge : 0;
    mg.zhi = (zmode == ZINTERIOR) ? g.nz - zedge : g.nz;
    if (g.zr_hi > g.zr_lo) {  // explicit plane range (single rank: never combined with a split)
        if (zmode != ZFULL) return 1;
        mg.zlo = g.zr_lo, mg.zhi = g.zr_hi;
    }
    int gz;
    if (zmode == ZBOUNDARY) {
        mg.zchunk = zedge;
        gz = 2;
    } else {
        const int span = mg.zhi - mg.zlo;
        if (span <= 0) return 0;
        mg.zchunk = pick_zchunk(gx * gy, span, MINB, NFZ + NFW, Epi::STREAMS);
        gz = (span + mg.zchunk - 1) / mg.zchunk;
    }
    kern<<<dim3(gx, gy, gz), dim3(MNT, 1, 1), smem, st>>>(maps, mg, epi);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}


// ---------------------------------------------------------------------------------------------
// Role-split variant of the engine: NROLE threads per grid point (blockDim = MNT * NROLE), each
// role running the SAME code on its own z-field (role r differentiates field r).  Used by the
// RHS kernel with one role per velocity component: a third of the work and of the live FP64
// state per thread, so 24 warps/SM are resident at <= 85 registers instead of 16 at 128, and
// the FP64 / shared-memory latency chains per plane are three times shorter.
//
// Epilogue concept (all threads call apply(); it may contain block-wide barriers):
//   void setup(const MarchGeom&, int i, int j, int role);
//   struct Pre;  Pre prefetch(long long m, bool ok) const;
//   void apply(const Ring<NFZ,0>&, long long m, int k, const Pre&, bool ok, int pt, double* xch);
// xch = NROLE * 3 * MNT doubles of shared scratch for exchanges between the roles.
template <int NFZ, int P, int NROLE, class Epi>
__global__ void __launch_bounds__(MNT* NROLE, 1)
    march_roles_kernel(const __grid_constant__ MarchMaps<NFZ> maps, const MarchGeom g, Epi epi) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NZS = 7 + P, NB = P + 1;
    constexpr int ZSTAGE = NFZ * MFIELD;  // doubles
    double* zring = reinterpret_cast<double*>(smem_raw);
    double* xch = zring + NZS * ZSTAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + NROLE * 3 * MNT);
    constexpr uint32_t PLANE_BYTES = MFIELD * 8;

    const int tid = threadIdx.x;
    const int pt = tid & (MNT - 1), role = tid / MNT;
    const int tx = pt & (MTX - 1), ty = pt >> 5;
    const int i0 = blockIdx.x * MTX, j0 = blockIdx.y * MTY;
    const int i = i0 + tx, j = j0 + ty;
    int kb, ke;
    if (g.zmode == 2) {
        kb = (blockIdx.z == 0) ? 0 : g.nz - g.zedge;
        ke = kb + g.zedge;
    } else {
        kb = g.zlo + blockIdx.z * g.zchunk;
        ke = min(g.zhi, kb + g.zchunk);
    }
    const bool in_dom = (i < g.nx) && (j < g.ny);

    const uint32_t zring_s = smem_u32(zring), bars_s = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NB; ++s) mbar_init(bars_s + 8 * s, 1);
        fence_barrier_init();
    }
    __syncthreads();

    const int cx = GX + i0 - MXO, cy = GH + j0 - R;
    auto issue_z = [&](int plane, uint32_t bar) {
        const int st = (plane - (kb - R)) % NZS;
#pragma unroll
        for (int f = 0; f < NFZ; ++f)
            tma_load_3d(zring_s + (uint32_t)(st * ZSTAGE + f * MFIELD) * 8, &maps.m[f], bar, cx, cy,
                        GH + plane);
    };
    const int niter = ke - kb;
    auto issue_group = [&](int n) {
        const uint32_t bar = bars_s + 8 * (n % NB);
        const int nz_planes = (n == 0) ? 7 : 1;
        mbar_expect_tx(bar, (uint32_t)(nz_planes * NFZ) * PLANE_BYTES);
        if (n == 0) {
#pragma unroll
            for (int s = 0; s < 6; ++s) issue_z(kb - R + s, bar);
        }
        issue_z(kb + n + R, bar);
    };
    if (tid == 0) {
        for (int n = 0; n < P && n < niter; ++n) issue_group(n);
    }

    const long long m0 = (long long)j * g.sy + i;
    epi.setup(g, i, j, role);
    typename Epi::Pre cur = epi.prefetch(m0 + (long long)kb * g.sz, in_dom);
    const int cell = (ty + R) * MBX + tx + MXO;

    for (int k = kb; k < ke; ++k) {
        const unsigned it = (unsigned)(k - kb);
        typename Epi::Pre nxt = epi.prefetch(m0 + (long long)(k + 1) * g.sz, in_dom && (k + 1 < ke));
        __syncthreads();
        if (tid == 0 && (int)it + P < niter) issue_group((int)it + P);
        mbar_wait(bars_s + 8 * (it % NB), (it / NB) & 1);
        Ring<NFZ, 0> r;
#pragma unroll
        for (int m = 0; m < 7; ++m) r.p[m] = zring + ((it + m) % NZS) * ZSTAGE + cell;
        r.q = nullptr;
        epi.apply(r, m0 + (long long)k * g.sz, k, cur, in_dom, pt, xch);
        cur = nxt;
    }
}

template <int NFZ, int P, int NROLE, class Epi>
int launch_march_roles(cudaStream_t st, const Geom& g, const MarchMaps<NFZ>& maps, const Epi& epi,
                       int zmode = ZFULL, int zedge = 0) {
    static bool attr_set = false;
    auto kern = march_roles_kernel<NFZ, P, NROLE, Epi>;
    constexpr int smem = NFZ * (7 + P) * MFIELD * 8 + NROLE * 3 * MNT * 8 + (P + 1) * 8;
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
    mg.bx = g.bx, mg.by = g.by, mg.bz_lo = g.bz_lo, mg.bz_hi = g.bz_hi;
    mg.gate = nullptr;
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    mg.zmode = zmode, mg.zedge = zedge;
    mg.zlo = (zmode == ZINTERIOR) ? zedge : 0;
    mg.zhi = (zmode == ZINTERIOR) ? g.nz - zedge : g.nz;
    if (g.zr_hi > g.zr_lo) {  // explicit plane range (single rank: never combined with a split)
        if (zmode != ZFULL) return 1;
        mg.zlo = g.zr_lo, mg.zhi = g.zr_hi;
    }
    int gz;
    if (zmode == ZBOUNDARY) {
        mg.zchunk = zedge;
        gz = 2;
    } else {
        const int span = mg.zhi - mg.zlo;
        if (span <= 0) return 0;
        // one CTA per SM: several waves of 148, chunks of >= 32 planes
        const int target_ctas = 148 * 6;
        int nchunks = (target_ctas + gx * gy - 1) / (gx * gy);
        int max_chunks = span / 32;
        if (max_chunks < 1) max_chunks = 1;
        if (nchunks > max_chunks) nchunks = max_chunks;
        if (nchunks < 1) nchunks = 1;
        mg.zchunk = (span + nchunks - 1) / nchunks;
        gz = (span + mg.zchunk - 1) / mg.zchunk;
    }
    kern<<<dim3(gx, gy, gz), dim3(MNT * NROLE, 1, 1), smem, st>>>(maps, mg, epi);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
