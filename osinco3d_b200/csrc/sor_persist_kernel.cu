// sor_persist_kernel.cu -- ALL iterations of one red-black SOR solve in ONE persistent launch
// (the pressure Poisson solve, src/poisson.f90:6-381; exit tests and dynamic omega :110-122).
//
// sor_tma_kernel.cu does one red+black iteration per launch, followed by a 1-thread control
// kernel (or the cooperative seam kernel) and a host poll every few iterations.  On the shipped
// grids (4.7 - 17 M points, 11 - 84 iterations per time step) that is launch- and poll-bound.
// Here a co-resident grid (cooperative launch, 3 CTAs per SM) keeps iterating:
//
//   iteration t:  every CTA sweeps its (tile, z-chunk) items -- the same TMA-staged, ping-pong
//                 red+black pass as sor_tma_kernel, bit for bit the same arithmetic;
//                 [odd periodic extents: grid barrier, odd red seam class, grid barrier, odd black
//                  seam class -- the in-place class sweeps of sor_kernels.cu on the new iterate]
//                 grid barrier; the LAST CTA to arrive evaluates dmax < eps, the stall exit,
//                 kmax and the dynamic-omega rule (sor_control_step) before it releases the
//                 others, which then read `done` / `omega` and go on or leave.
//
// No host round trip and no launch inside the solve: the host enqueues the kernel and (o3d_step)
// the projection correction gated on its outcome.  Same class order and arithmetic as the
// launch-per-pass path => bitwise equal iterates, iteration counts and omega history
// (tests/test_gpu_poisson.py::test_persistent_solve_equals_launch_per_pass_bitwise).
//
// Z slabs (nranks > 1): the three boundary planes per side of the new iterate (two for the next
// pass, the third for the projection correction behind the solve) are stored straight into the
// neighbour rank's ghost planes through peer-mapped pointers (CUDA IPC over NVLink) by the CTAs
// that compute them, followed by a system-scope release on a counter in the neighbour's memory; a
// CTA whose chunk touches a slab end acquires that counter before it stages ghost planes.  The
// residual maximum travels as flag-in-data words (32 data bits | 32-bit iteration tag) into a slot
// per source rank on every rank and is combined by the controlling CTA of each rank (identical
// inputs -> identical decisions).  No NCCL call and no host inside the solve.
#include <cstring>

#include "kernels.h"
#include "sor_common.cuh"
#include "tma.cuh"

namespace o3d {
namespace {

constexpr int GTX = 32, GTY = 16, GNT = 256;
constexpr int GBX = GTX + 4, GBY = GTY + 4, GPL = GBX * GBY;  // 36 x 20 = 720 cells, 5760 B
constexpr int GP = 2;                                         // planes prefetched ahead
constexpr int GNP = 5 + GP, GNR = 3 + GP, GNB = GP + 1;       // p stages, rhs stages, barriers
constexpr int GSMEM = (GNP + GNR) * GPL * 8 + GNB * 8 + 32 * 8 + 16;
// planes per side a rank stores into its neighbours' ghost planes: the pass reads 2, the projection
// correction queued behind the solve differentiates pp with radius 3 -- with the third plane delivered
// by the solve itself the correction needs no exchange at all
constexpr int PD = 3;
// every wait on another CTA / rank is bounded (a lost peer must not hang the GPU): ~20 s of SM
// clocks, far beyond any legitimate skew between ranks (host-side hiccups included)
constexpr long long SPIN_LIMIT = 40000000000ll;

struct PersistArgs {
    SorArgs s;            // geometry, operator, neighbour rule; s.pp / s.rhs unused by the pass
    double* p[2];         // interior origins of the two ping-pong buffers
    int bx, by, bz_lo, bz_hi;  // closures for the ghost images of the new iterate
    int tiles_x, tiles_y, nch, zchunk, zstagger;
    int first_src;        // buffer read by the first iteration of this launch
    int max_iters;        // iterations this launch may run
    double eps, factor;
    int kmax, idyn;
    int fixed;            // smoother mode: no exit tests, the control step only counts
    int dynamic;          // items handed out through an atomic ticket counter
    long long nxf, nyf, nzf;  // sizes of the x / y / z seam planes (SEAM)
    // ---- z slabs: peer-mapped neighbours (null: none on that side) ----
    PeerSync peer;
};

struct alignas(64) PersistMaps {
    CUtensorMap p[2], rhs;
};

template <int OFF>
__device__ __forceinline__ double lds(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ void sts(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// generic-proxy global writes <-> async-proxy (TMA) reads of the same addresses
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// exit tests + dynamic omega on a private copy (every field of the control block is read and
// written through volatile accesses: other SMs updated it in earlier iterations)
__device__ __forceinline__ void control_step_volatile(SorCtrl* ctrl, unsigned long long dmax_bits,
                                                      double eps, int kmax, int idyn, double factor,
                                                      int fixed) {
    volatile SorCtrl* v = ctrl;
    SorCtrl c;
    c.dmax_bits = dmax_bits;
    c.omega = v->omega, c.dmax_old = v->dmax_old, c.dmax_last = v->dmax_last;
    c.iter = v->iter, c.done = v->done;
    if (fixed) {
        c.iter += 1, c.dmax_bits = 0ull;
    } else {
        sor_control_step(&c, eps, kmax, idyn, factor);
    }
    v->dmax_bits = c.dmax_bits;
    v->omega = c.omega, v->dmax_old = c.dmax_old, v->dmax_last = c.dmax_last;
    v->iter = c.iter;
    v->done = c.done;
}

template <bool SEAM, bool MULTI>
__global__ void __launch_bounds__(GNT, 3)
    sor_persist_kernel(const __grid_constant__ PersistMaps maps, const PersistArgs a, SorCtrl* ctrl,
                       unsigned long long* sync) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sp = reinterpret_cast<double*>(smem_raw);
    double* sr = sp + GNP * GPL;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sr + GNR * GPL);
    double* red = reinterpret_cast<double*>(bars + GNB);
    int* ticket = reinterpret_cast<int*>(red + 32);
    const int tid = threadIdx.x;
    const uint32_t sp_s = smem_u32(sp), sr_s = smem_u32(sr), bars_s = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GNB; ++s) mbar_init(bars_s + 8 * s, 1);
        fence_barrier_init();
    }
    __syncthreads();

    const SorArgs& g = a.s;
    const unsigned G = gridDim.x;
    const int ntiles = a.tiles_x * a.tiles_y, nitems = ntiles * a.nch;
    unsigned long long epoch = 0;  // grid barriers passed since launch (thread 0)
    uint32_t gbase = 0;            // TMA groups issued (= waited for) by this CTA so far
    constexpr uint32_t PLB = GPL * 8;
    const uint32_t p_end = sp_s + GNP * PLB, r_end = sr_s + GNR * PLB;
    const int tx = tid & 31, typ = tid >> 5;
    const int own = (2 + 2 * typ) * GBX + 2 + tx;
    const uint32_t ownA8 = own * 8, ownB8 = (own + GBX) * 8;
    // ring-1 pairs (48 threads, warps 0 and 1): each holds exactly one red cell in every plane
    int rcell = -1, rstep = 0, rlx = 0, rly = 0;
    if (tid < 48) {
        if (tid < 16) rlx = 2 + 2 * tid, rly = 1, rstep = 1;                       // below the tile
        else if (tid < 32) rlx = 2 + 2 * (tid - 16), rly = GBY - 2, rstep = 1;     // above
        else if (tid < 40) rlx = 1, rly = 2 + 2 * (tid - 32), rstep = GBX;         // left
        else rlx = GBX - 2, rly = 2 + 2 * (tid - 40), rstep = GBX;                 // right
        rcell = rly * GBX + rlx;
    }
    const bool has_ring = rcell >= 0;
    const uint32_t ring0 = has_ring ? rcell * 8 : 0, ring1 = has_ring ? (rcell + rstep) * 8 : 0;
    constexpr bool multi = MULTI;

    // grid barrier; `last` runs in thread 0 of the last CTA to arrive, before the others go on
    auto grid_barrier = [&](auto&& last) {
        __syncthreads();
        if (tid == 0) {
            epoch += 1;
            if (MULTI) __threadfence_system();  // this CTA's stores into peer memory included
            else __threadfence();
            const unsigned long long t = atomicAdd(&sync[0], 1ull);
            if (t + 1 == epoch * G) {
                last();
                __threadfence();
                atomicExch(&sync[16], epoch);
            } else {
                const long long t0 = clock64();
                while (ld_acquire_gpu(&sync[16]) < epoch) {
                    if (clock64() - t0 > SPIN_LIMIT) {
                        atomicExch(&ctrl->done, 9);
                        break;
                    }
                }
            }
            __threadfence();
        }
        __syncthreads();
    };

    if (MULTI && a.peer.push_init) {
        // phase 0 (z slabs): the ghost planes of the INITIAL iterate (2 per side) and of the
        // right-hand side (1 per side; constant over the solve) are stored straight into the
        // neighbours' ghost planes -- whole padded planes, so that the x / y ghost cells travel
        // too -- by all CTAs together; the last CTA to finish releases one count per neighbour.
        // No NCCL call in front of the solve: interior items start at once, items that touch a
        // slab end acquire the neighbour's count first.
        const long long pl = g.sz, org = (long long)GX + (long long)GH * g.sy;  // plane start
        const double* ps = a.p[a.first_src] - org;
        const double* rs = g.rhs - org;
        const long long stride = (long long)G * GNT;
        const long long t0i = (long long)blockIdx.x * GNT + tid;
        if (a.peer.has_lo) {  // my planes 0, 1 (p) and 0 (rhs) -> lower neighbour's planes nz_lo + ..
            double* pd = a.peer.lo_p[a.first_src] - org + (long long)a.peer.lo_nz * pl;
            double* rd = a.peer.lo_rhs - org + (long long)a.peer.lo_nz * pl;
            for (long long q = t0i; q < 2 * pl; q += stride) pd[q] = ps[q];
            for (long long q = t0i; q < pl; q += stride) rd[q] = rs[q];
        }
        if (a.peer.has_hi) {  // my planes nz-2, nz-1 (p) and nz-1 (rhs) -> upper neighbour's -2, -1
            double* pd = a.peer.hi_p[a.first_src] - org - 2 * pl;
            double* rd = a.peer.hi_rhs - org - pl;
            const double* p2 = ps + (long long)(g.nz - 2) * pl;
            const double* r2 = rs + (long long)(g.nz - 1) * pl;
            for (long long q = t0i; q < 2 * pl; q += stride) pd[q] = p2[q];
            for (long long q = t0i; q < pl; q += stride) rd[q] = r2[q];
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            if (atomicAdd(&sync[8], 1ull) == (unsigned long long)(G - 1)) {
                __threadfence_system();
                if (a.peer.has_lo) red_release_sys_add(&a.peer.lo->init_cnt[1], 1ull);
                if (a.peer.has_hi) red_release_sys_add(&a.peer.hi->init_cnt[0], 1ull);
            }
        }
    }

    for (int it = 0; it < a.max_iters; ++it) {
        if (*((volatile int*)&ctrl->done)) break;  // uniform over the grid (read after a barrier)
        const double omega = *((volatile double*)&ctrl->omega);
        const double one_m_omega = 1.0 - omega;
        const int src = a.first_src ^ (it & 1);
        const CUtensorMap* pmap = &maps.p[src];
        double* const p_new = a.p[src ^ 1];
        // iterations completed by every rank before this one, over the whole session
        const unsigned long long T = a.peer.iter_base + (unsigned long long)it;
        double dmax = 0.0;

        // items of an iteration: static round-robin (CTA b takes b, b + G, ...) or, a.dynamic,
        // tickets from a counter per iteration parity (zeroed for the iteration after next by the
        // controlling CTA) -- the order in which the hardware would hand out CTAs
        int item = blockIdx.x;
        if (a.dynamic) {
            __syncthreads();
            if (tid == 0) *ticket = (int)atomicAdd(&sync[4 + (it & 1)], 1ull);
            __syncthreads();
            item = *ticket;
        }
        while (item < nitems) {
            const int tile = item % ntiles;
            int ch = item / ntiles;
            if (MULTI && a.nch > 2 && ch < 2) {
                // z slabs: the first wave sweeps chunk 1, chunk 0 follows -- by then the lower
                // neighbour's phase 0 has delivered its ghost planes, so nobody waits for it.
                // (Putting both slab-end chunks LAST was measured slower, 0.155 vs 0.149 ms per
                // iteration at 256^2 x 255 per GPU: their NVLink stores then sit in the tail.)
                ch ^= 1;
            }
            const int i0 = (tile % a.tiles_x) * GTX, j0 = (tile / a.tiles_x) * GTY;
            // (experiment knob, default 0: z-chunk boundaries of tiles of odd parity shifted by
            // `zstagger` planes, so that a tile runs a few planes ahead of its four neighbours)
            const int zsh = (((tile % a.tiles_x) + (tile / a.tiles_x)) & 1) ? a.zstagger : 0;
            const int kb = ch ? ch * a.zchunk + zsh : 0;
            const int ke = ch + 1 < a.nch ? (ch + 1) * a.zchunk + zsh : g.nz;
            const int niter = ke - kb;
            const int ngroups = niter > 1 ? niter - 1 : 1;  // group n feeds red(kb + n + 2)
            __syncthreads();  // the previous item's stages (and `red`) are no longer read

            // p plane q lives in stage (q - (kb-2)) mod GNP, rhs plane q in (q - (kb-1)) mod GNR
            const int cx = GX + i0 - 2, cy = GH + j0 - 2;
            auto issue_p = [&](int plane, uint32_t bar) {
                const unsigned st = (unsigned)(plane - (kb - 2)) % GNP;
                tma_load_3d(sp_s + st * PLB, pmap, bar, cx, cy, GH + plane);
            };
            auto issue_r = [&](int plane, uint32_t bar) {
                const unsigned st = (unsigned)(plane - (kb - 1)) % GNR;
                tma_load_3d(sr_s + st * PLB, &maps.rhs, bar, cx, cy, GH + plane);
            };
            auto issue_group = [&](int n) {
                const uint32_t bar = bars_s + 8 * ((gbase + n) % GNB);
                if (n == 0) {
                    mbar_expect_tx(bar, (6 + 4) * PLB);
#pragma unroll
                    for (int q = -2; q <= 3; ++q) issue_p(kb + q, bar);
#pragma unroll
                    for (int q = -1; q <= 2; ++q) issue_r(kb + q, bar);
                } else {
                    mbar_expect_tx(bar, 2 * PLB);
                    issue_p(kb + n + 3, bar);
                    issue_r(kb + n + 2, bar);
                }
            };
            auto wait_group = [&](int n) {
                const uint32_t q = gbase + n;
                mbar_wait(bars_s + 8 * (q % GNB), (q / GNB) & 1u);
            };
            if (tid == 0) {
                if (multi && it == 0 && a.peer.push_init) {
                    // the neighbours' phase 0 of THIS solve: one count per solve and side
                    const unsigned long long need = a.peer.solve_base + 1ull;
                    const long long t0 = clock64();
                    if (kb < 2 && a.peer.has_lo)
                        while (ld_acquire_sys(&a.peer.mine->init_cnt[0]) < need)
                            if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                    if (ke + 1 >= g.nz && a.peer.has_hi)
                        while (ld_acquire_sys(&a.peer.mine->init_cnt[1]) < need)
                            if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                }
                if (multi && it > 0) {
                    // ghost planes of the previous iterate come from the neighbours' CTAs: PD
                    // planes per tile and side, counted in OUR memory over the whole session
                    const unsigned long long need = (unsigned long long)PD * ntiles * T;
                    const long long t0 = clock64();
                    if (SEAM) {  // ... and both odd seam sweeps of that iteration
                        if (kb < 2 && a.peer.has_lo)
                            while (ld_acquire_sys(&a.peer.mine->seam_cnt[0]) < 2ull * T)
                                if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                        if (ke + 1 >= g.nz && a.peer.has_hi)
                            while (ld_acquire_sys(&a.peer.mine->seam_cnt[1]) < 2ull * T)
                                if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                    }
                    if (kb < 2 && a.peer.has_lo)
                        while (ld_acquire_sys(&a.peer.mine->halo_cnt[0]) < need)
                            if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                    if (ke + 1 >= g.nz && a.peer.has_hi)
                        while (ld_acquire_sys(&a.peer.mine->halo_cnt[1]) < need)
                            if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                }
                fence_proxy_async_global();
                for (int n = 0; n < GP && n < ngroups; ++n) issue_group(n);
            }

            // ---- the pass of sor_tma_kernel.cu over this item (same code path, same bits) ----
            // Own Y-PAIR: column 2+tx, rows 2+2*typ (member A) and 3+2*typ (member B): one red and
            // one black cell in every plane; pe = colour of member A in global plane 0.
            const int gi = i0 + tx, gj = j0 + 2 * typ;
            const bool inA = gi < g.nx && gj < g.ny, inB = gi < g.nx && gj + 1 < g.ny;
            const int pe = (gi + gj + g.gz0) & 1;
            const int rpar = has_ring ? ((i0 - 2 + rlx + j0 - 2 + rly + g.gz0) & 1) : 0;
            // SEAM: seam parity of the own pair in x,y (bit 0: member A, bit 1: member B) and of the
            // two cells of the ring pair (2 bits each: 0 / 1 = parity, 2 = ghost cell of an odd
            // periodic axis, never updated)
            int own_par = 0, ring_code = 0;
            if (SEAM) {
                const int sx = (g.seam_x && gi == g.nx - 1) ? 1 : 0;
                own_par = (sx ^ ((g.seam_y && gj == g.ny - 1) ? 1 : 0)) |
                          ((sx ^ ((g.seam_y && gj + 1 == g.ny - 1) ? 1 : 0)) << 1);
                if (has_ring) {
                    for (int m = 0; m < 2; ++m) {
                        const int c = rcell + m * rstep;
                        const int ri = i0 - 2 + c % GBX, rj = j0 - 2 + c / GBX;
                        int code = 0;
                        if (g.seam_x) code = (ri < 0 || ri >= g.nx) ? 2 : (ri == g.nx - 1);
                        if (g.seam_y && code != 2)
                            code = (rj < 0 || rj >= g.ny) ? 2 : (code ^ (rj == g.ny - 1 ? 1 : 0));
                        ring_code |= code << (2 * m);
                    }
                }
            }
            auto zseam = [&](int q) -> int {
                if (!SEAM || !g.seam_z) return 0;
                const int gk = g.gz0 + q;
                return (gk < 0 || gk >= g.gnz) ? 2 : (gk == g.gnz - 1 ? 1 : 0);
            };
            const Img2 ix = image_offsets(gi, g.nx, a.bx, a.bx);
            const Img2 iyA = image_offsets(gj, g.ny, a.by, a.by);
            const Img2 iyB = image_offsets(gj + 1, g.ny, a.by, a.by);
            const bool xy_img = (ix.lo | ix.hi | iyA.lo | iyA.hi | iyB.lo | iyB.hi) != 0;
            // planes whose points have z images (walls / periodic wrap handled by this rank)
            const int zimg_lo = (a.bz_lo == BM_MIRROR || a.bz_hi == BM_WRAP) ? R : -1;  // k <= zimg_lo
            const int zimg_hi = (a.bz_hi == BM_MIRROR || a.bz_lo == BM_WRAP) ? g.nz - 1 - R : g.nz;
            // z slabs: this pair's position in plane 0 of the neighbours' frames (my planes 0, 1 are
            // the lower neighbour's ghost planes nz_lo, nz_lo + 1; my planes nz-2, nz-1 the upper
            // neighbour's ghost planes -2, -1)
            double* peer_lo = nullptr;
            double* peer_hi = nullptr;
            if (MULTI) {
                const long long pair = (long long)gj * g.sy + gi;
                if (a.peer.has_lo) peer_lo = a.peer.lo_p[src ^ 1] + (long long)a.peer.lo_nz * g.sz + pair;
                if (a.peer.has_hi) peer_hi = a.peer.hi_p[src ^ 1] - (long long)g.nz * g.sz + pair;
            }

            double dloc = 0.0;
            auto nextp = [&](uint32_t x) { return x + PLB == p_end ? sp_s : x + PLB; };
            auto nextr = [&](uint32_t x) { return x + PLB == r_end ? sr_s : x + PLB; };
            // SOR update of the cell at byte offset c of the plane at a0 (am / ap = planes below /
            // above, ar = rhs plane): src/poisson.f90:95-102 with "/ A" as "* (1/A)"; returns the
            // relaxed value, d = |p_new - p_old|, pc = the old value
            auto update = [&](uint32_t am, uint32_t a0, uint32_t ap, uint32_t ar, uint32_t c,
                              double& d, double& pc) -> double {
                const uint32_t c0 = a0 + c;
                pc = lds<0>(c0);
                const double w = lds<-8>(c0), e = lds<8>(c0);
                const double sn = lds<-GBX * 8>(c0), nn = lds<GBX * 8>(c0);
                const double bb = lds<0>(am + c), tt = lds<0>(ap + c), rr = lds<0>(ar + c);
                const double pn =
                    (-(g.oneondx2 * (w + e)) - g.oneondy2 * (sn + nn) - g.oneondz2 * (bb + tt) + rr) *
                    g.invA;
                d = fabs(pn - pc);                     // :100
                return one_m_omega * pc + omega * pn;  // :102
            };
            const uint32_t cxor = ownA8 ^ ownB8, rxor = ring0 ^ ring1;
            // r = 1: member B of the pair is the red one in the current plane (member A otherwise);
            // c_red / c_blk = byte offsets of the red / black member; rc = the red cell of the ring
            // pair, rmm = which of its two cells that is.  All flip from plane to plane.
            int r = (pe + kb) & 1, rmm = (rpar + kb) & 1;
            uint32_t c_red = r ? ownB8 : ownA8, c_blk = r ? ownA8 : ownB8;
            uint32_t rc = rmm ? ring1 : ring0;
            auto flip = [&]() { r ^= 1, rmm ^= 1, c_red ^= cxor, c_blk ^= cxor, rc ^= rxor; };
            // red half-sweep of the plane at a0 (local plane q) over the own pair and the ring pair;
            // counted: the plane is owned by this chunk
            auto red_plane = [&](int q, uint32_t am, uint32_t a0, uint32_t ap, uint32_t ar,
                                 bool counted) {
                double d, pc;
                double v = update(am, a0, ap, ar, c_red, d, pc);
                bool ring_on = has_ring;
                if (SEAM) {
                    const int zs = zseam(q);
                    if (zs == 2 || (((own_par >> r) ^ zs) & 1)) v = pc, d = 0.0;  // odd class: keep
                    const int code = (ring_code >> (2 * rmm)) & 3;
                    ring_on = ring_on && zs != 2 && code != 2 && !((code ^ zs) & 1);
                }
                if (ring_on) {
                    double dr, pr;
                    const double vr = update(am, a0, ap, ar, rc, dr, pr);
                    sts(a0 + rc, vr);
                }
                sts(a0 + c_red, v);
                if (counted && (r ? inB : inA)) dloc = fmax(dloc, d);
            };

            wait_group(0);  // p planes kb-2 .. kb+3 in stages 0 .. 5, rhs kb-1 .. kb+2 in 0 .. 3
            flip();  // plane kb-1 has the other parity
            red_plane(kb - 1, sp_s, sp_s + PLB, sp_s + 2 * PLB, sr_s, false);
            flip();
            red_plane(kb, sp_s + PLB, sp_s + 2 * PLB, sp_s + 3 * PLB, sr_s + PLB, true);
            flip();
            red_plane(kb + 1, sp_s + 2 * PLB, sp_s + 3 * PLB, sp_s + 4 * PLB, sr_s + 2 * PLB,
                      kb + 1 < ke);
            flip();  // back to the parity of plane kb

            uint32_t a_m1 = sp_s + PLB, a_0 = sp_s + 2 * PLB, a_1 = sp_s + 3 * PLB,
                     a_2 = sp_s + 4 * PLB, a_3 = sp_s + 5 * PLB;
            uint32_t ar_0 = sr_s + PLB, ar_2 = sr_s + 3 * PLB;
            double* outp = p_new + (long long)kb * g.sz + (long long)gj * g.sy + gi;
            for (int k = kb; k < ke; ++k) {
                const int n = k - kb;
                __syncthreads();  // step k-1 done: its oldest stages may be refilled; red(k+1) visible
                if (tid == 0 && n + GP < ngroups) issue_group(n + GP);
                if (n >= 1 && n < ngroups) wait_group(n);
                // planes k+2 and k have the same parity: the same member is red in both
                if (k + 2 <= ke) red_plane(k + 2, a_1, a_2, a_3, ar_2, k + 2 < ke);
                {
                    // black member of the own pair in plane k: all six neighbours hold new red values
                    double d, pcb;
                    double vb = update(a_m1, a_0, a_1, ar_0, c_blk, d, pcb);
                    if (SEAM && (((own_par >> (1 - r)) ^ zseam(k)) & 1))
                        vb = pcb, d = 0.0;  // odd class: swept after the pass
                    const double vred = lds<0>(a_0 + c_red);
                    if (r ? inA : inB) dloc = fmax(dloc, d);
                    const double vA = r ? vb : vred, vB = r ? vred : vb;
                    if (inA) outp[0] = vA;
                    if (inB) outp[g.sy] = vB;
                    if (xy_img || k <= zimg_lo || k >= zimg_hi) {  // boundary-adjacent points only
                        const Img2 iz = image_offsets(k, g.nz, a.bz_lo, a.bz_hi);
                        if (inA)
                            store_images(outp, 0, vA, ix, iyA.lo * g.sy, iyA.hi * g.sy, iz.lo * g.sz,
                                         iz.hi * g.sz);
                        if (inB)
                            store_images(outp, g.sy, vB, ix, iyB.lo * g.sy, iyB.hi * g.sy,
                                         iz.lo * g.sz, iz.hi * g.sz);
                    }
                    if (MULTI) {
                        // the PD planes next to a rank boundary also go straight into the
                        // neighbour's ghost planes (value + x / y images: its TMA boxes read them)
                        double* o = nullptr;
                        if (k < PD && peer_lo) o = peer_lo + (long long)k * g.sz;
                        double* o2 = (k >= g.nz - PD && peer_hi) ? peer_hi + (long long)k * g.sz : nullptr;
                        if (!o) o = o2, o2 = nullptr;  // (slabs thinner than 2 PD planes: both sides)
                        for (; o; o = o2, o2 = nullptr) {
                            if (inA) {
                                o[0] = vA;
                                store_images(o, 0, vA, ix, iyA.lo * g.sy, iyA.hi * g.sy, 0, 0);
                            }
                            if (inB) {
                                o[g.sy] = vB;
                                store_images(o, g.sy, vB, ix, iyB.lo * g.sy, iyB.hi * g.sy, 0, 0);
                            }
                        }
                    }
                }
                outp += g.sz;
                flip();
                a_m1 = a_0, a_0 = a_1, a_1 = a_2, a_2 = a_3, a_3 = nextp(a_3);
                ar_0 = nextr(ar_0), ar_2 = nextr(ar_2);
            }
            gbase += ngroups;
            dmax = fmax(dmax, dloc);
            if (multi) {
                // this item's copies in the neighbours' ghost planes are complete: release them
                const int nlo = a.peer.has_lo ? max(0, min(ke, PD) - kb) : 0;
                const int nhi = a.peer.has_hi ? max(0, ke - max(kb, g.nz - PD)) : 0;
                if (nlo | nhi) {
                    __syncthreads();
                    if (tid == 0) {
                        if (nlo) red_release_sys_add(&a.peer.lo->halo_cnt[1], (unsigned long long)nlo);
                        if (nhi) red_release_sys_add(&a.peer.hi->halo_cnt[0], (unsigned long long)nhi);
                    }
                }
            }
            if (a.dynamic) {
                __syncthreads();
                if (tid == 0) *ticket = (int)atomicAdd(&sync[4 + (it & 1)], 1ull);
                __syncthreads();
                item = *ticket;
            } else {
                item += G;
            }
        }
        {
            const double bm = block_max(dmax, red);
            if (tid == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
        }
        fence_proxy_async_global();  // this iteration's stores -> the next iteration's TMA loads

        if (SEAM) {
            // the two thin odd classes, in place on the new iterate, rewriting the ghost images of
            // what they touch (sor_kernels.cu)
            SorArgs sa = g;
            sa.pp = p_new;
            const long long tot = a.nxf + a.nyf + a.nzf;
            const long long stride = (long long)G * GNT;
            double dm = 0.0;
            for (int colour = 0; colour < 2; ++colour) {
                grid_barrier([&] {
                    if (!MULTI) return;
                    // z slabs: an odd point next to a rank boundary reads the neighbour's plane.
                    // Before the red class that plane must hold the neighbour's PASS of this
                    // iteration (its epilogue stores: halo_cnt), before the black class its red
                    // odd sweep (seam_cnt); our own red sweep is released to the neighbours here.
                    const long long t0 = clock64();
                    unsigned long long* cnt[2] = {nullptr, nullptr};
                    unsigned long long need = 0ull;
                    if (colour == 0) {
                        cnt[0] = &a.peer.mine->halo_cnt[0], cnt[1] = &a.peer.mine->halo_cnt[1];
                        need = (unsigned long long)PD * ntiles * (T + 1ull);
                    } else {
                        if (a.peer.has_lo) red_release_sys_add(&a.peer.lo->seam_cnt[1], 1ull);
                        if (a.peer.has_hi) red_release_sys_add(&a.peer.hi->seam_cnt[0], 1ull);
                        cnt[0] = &a.peer.mine->seam_cnt[0], cnt[1] = &a.peer.mine->seam_cnt[1];
                        need = 2ull * T + 1ull;
                    }
                    for (int side = 0; side < 2; ++side) {
                        if (!(side ? a.peer.has_hi : a.peer.has_lo)) continue;
                        while (ld_acquire_sys(cnt[side]) < need)
                            if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                    }
                });
                for (long long t = (long long)blockIdx.x * GNT + tid; t < tot; t += stride) {
                    int i, j, kk;
                    bool ok;
                    seam_point_of(sa, t, a.nxf, a.nyf, a.nzf, i, j, kk, ok);
                    if (ok) {
                        const int gk = sa.gz0 + kk;
                        if (((i + j + gk) & 1) == colour && (seam_pop(sa, i, j, gk) & 1)) {
                            dm = fmax(dm, sor_point<true>(sa, i, j, kk, omega));
                            if (MULTI && (kk < PD || kk >= g.nz - PD)) {
                                // the point lies in a plane the neighbour keeps as a ghost plane
                                const long long m = (long long)j * g.sy + i;
                                const double v = p_new[(long long)kk * g.sz + m];
                                const Img2 jx = image_offsets(i, g.nx, a.bx, a.bx);
                                const Img2 jy = image_offsets(j, g.ny, a.by, a.by);
                                if (kk < PD && a.peer.has_lo) {
                                    double* o = a.peer.lo_p[src ^ 1] +
                                                (long long)(a.peer.lo_nz + kk) * g.sz + m;
                                    o[0] = v;
                                    store_images(o, 0, v, jx, jy.lo * g.sy, jy.hi * g.sy, 0, 0);
                                }
                                if (kk >= g.nz - PD && a.peer.has_hi) {
                                    double* o = a.peer.hi_p[src ^ 1] + (long long)(kk - g.nz) * g.sz + m;
                                    o[0] = v;
                                    store_images(o, 0, v, jx, jy.lo * g.sy, jy.hi * g.sy, 0, 0);
                                }
                            }
                        }
                    }
                }
            }
            const double bm = block_max(dm, red);
            if (tid == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
            fence_proxy_async_global();
        }

        grid_barrier([&] {
            // every CTA has drawn its last ticket of this iteration: re-arm the counter for it + 2
            if (a.dynamic) *((volatile unsigned long long*)&sync[4 + (it & 1)]) = 0ull;
            unsigned long long bits = *((volatile unsigned long long*)&ctrl->dmax_bits);
            if (MULTI && SEAM) {  // our black odd sweep is complete: 2 (T + 1) sweeps delivered
                if (a.peer.has_lo) red_release_sys_add(&a.peer.lo->seam_cnt[1], 1ull);
                if (a.peer.has_hi) red_release_sys_add(&a.peer.hi->seam_cnt[0], 1ull);
            }
            if (multi) {
                // all-to-all of the local maxima through peer memory, flag-in-data: the 64-bit
                // pattern travels as two 8-byte words {32 data bits, 32-bit tag = T + 1}, each a
                // single store that is its own arrival flag -- no fence, no separate flag round
                // trip.  Slots alternate with the iteration parity (a rank is at most one iteration
                // ahead).  Same inputs on every rank -> identical exit / omega decisions.
                const int me = a.peer.rank, P = a.peer.nranks;
                const unsigned long long tag = (T + 1ull) & 0xffffffffull;
                const unsigned long long w0 = (bits >> 32 << 32) | tag;
                const unsigned long long w1 = (bits << 32) | tag;
                for (int r = 0; r < P; ++r) {
                    volatile unsigned long long* q = &a.peer.all[r]->dmax_slot[T & 1][2 * me];
                    q[0] = w0, q[1] = w1;
                }
                const long long t0 = clock64();
                for (int r = 0; r < P; ++r) {
                    volatile unsigned long long* q = &a.peer.mine->dmax_slot[T & 1][2 * r];
                    unsigned long long v0, v1;
                    while (((v0 = q[0]) & 0xffffffffull) != tag || ((v1 = q[1]) & 0xffffffffull) != tag)
                        if (clock64() - t0 > SPIN_LIMIT) { atomicExch(&ctrl->done, 9); break; }
                    const unsigned long long b = (v0 >> 32 << 32) | (v1 >> 32);
                    bits = b > bits ? b : bits;
                }
            }
            if (*((volatile int*)&ctrl->done) != 9)
                control_step_volatile(ctrl, bits, a.eps, a.kmax, a.idyn, a.factor, a.fixed);
        });
    }
}

int g_persist_ctas = 0;  // co-resident CTAs (3 per SM x SMs), queried once

int persist_capacity() {
    if (g_persist_ctas) return g_persist_ctas;
    if (cudaFuncSetAttribute(sor_persist_kernel<false, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, GSMEM) != cudaSuccess ||
        cudaFuncSetAttribute(sor_persist_kernel<true, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, GSMEM) != cudaSuccess ||
        cudaFuncSetAttribute(sor_persist_kernel<false, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, GSMEM) != cudaSuccess ||
        cudaFuncSetAttribute(sor_persist_kernel<true, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, GSMEM) != cudaSuccess)
        return 0;
    int dev = 0, sms = 0, coop = 0, per_sm = 0, per_sm2 = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return 0;
    int per_sm3 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sor_persist_kernel<false, false>, GNT,
                                                  GSMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, sor_persist_kernel<true, false>, GNT,
                                                  GSMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, sor_persist_kernel<false, true>, GNT,
                                                  GSMEM);
    int per_sm4 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm4, sor_persist_kernel<true, true>, GNT,
                                                  GSMEM);
    if (per_sm2 < per_sm) per_sm = per_sm2;
    if (per_sm3 < per_sm) per_sm = per_sm3;
    if (per_sm4 < per_sm) per_sm = per_sm4;
    if (per_sm > 3) per_sm = 3;
    {   // tuning override: co-resident CTAs per SM
        const char* e = getenv("O3D_PERSIST_CTAS");
        if (e && atoi(e) > 0 && atoi(e) < per_sm) per_sm = atoi(e);
    }
    g_persist_ctas = per_sm * sms;
    return g_persist_ctas;
}

}  // namespace

// z chunks of the persistent pass: every CTA sweeps ceil(items / G) items per iteration, each
// costing its planes plus ~2.5 plane-equivalents of pipeline prologue (3 extra red planes, the
// fill latency); pick the chunk count that minimises waves x (chunk + 2.5), fewer chunks on a tie
int persist_pick_chunks(int ntiles, int nz, int G) {
    const char* e = getenv("O3D_NCH_P");
    if (e && atoi(e) > 0) return atoi(e) > nz / 2 ? (nz / 2 > 0 ? nz / 2 : 1) : atoi(e);
    int best = 1;
    double best_cost = 1e300;
    const int maxch = nz / 4 > 0 ? nz / 4 : 1;
    for (int nch = 1; nch <= maxch && nch <= 256; ++nch) {
        const int zc = (nz + nch - 1) / nch;
        const int real = (nz + zc - 1) / zc;
        const long long items = (long long)ntiles * real;
        const long long waves = (items + G - 1) / G;
        const double cost = (double)waves * (zc + 2.5);
        if (cost < best_cost - 1e-9) best_cost = cost, best = real;
    }
    return best;
}

int sor_persist_available() { return persist_capacity() > 0; }

int launch_sor_persist(cudaStream_t st, const SorArgs& a, const CUtensorMap* pmap0,
                       const CUtensorMap* pmap1, const CUtensorMap* rhs_map, double* p0, double* p1,
                       int first_src, int bx, int by, int bz_lo, int bz_hi, SorCtrl* ctrl,
                       unsigned long long* sync, int max_iters, double eps, int kmax, int idyn,
                       double factor, int fixed, const PeerSync* peer) {
    const int G_max = persist_capacity();
    if (G_max <= 0) return 2;
    PersistMaps maps;
    maps.p[0] = *pmap0, maps.p[1] = *pmap1, maps.rhs = *rhs_map;
    PersistArgs f;
    f.s = a;
    f.s.pp = nullptr;
    f.p[0] = p0, f.p[1] = p1;
    f.bx = bx, f.by = by, f.bz_lo = bz_lo, f.bz_hi = bz_hi;
    f.tiles_x = (a.nx + GTX - 1) / GTX, f.tiles_y = (a.ny + GTY - 1) / GTY;
    const int ntiles = f.tiles_x * f.tiles_y;
    {
        const char* e = getenv("O3D_PERSIST_DYN");
        const char* en = getenv("O3D_NCH_P");
        if ((e && e[0] == '0') || (en && atoi(en) > 0)) {
            f.nch = persist_pick_chunks(ntiles, a.nz, G_max);
            f.zchunk = (a.nz + f.nch - 1) / f.nch;
        } else {
            // tickets balance like the hardware's CTA scheduler: the chunk model of the
            // launch-per-pass kernel (throughput + half a chunk of tail) applies
            f.zchunk = pick_zchunk_slots(ntiles, a.nz, G_max, 3.3);
        }
        f.nch = (a.nz + f.zchunk - 1) / f.zchunk;
    }
    {
        // the last chunk of a shifted tile is `zstagger` planes shorter: keep it >= 2 planes
        // O3D_PERSIST_STAGGER=<planes>: measured 0 .. 16 planes without any effect on the DRAM
        // traffic of the static map (profiles/r2_sor_scheduling.txt).  Default: off on one GPU; 4
        // planes on z slabs -- no measurable effect there either, but it is the setting every
        // multi-GPU parity and scaling run of round 2 was recorded with (profiles/r2t_*, r2u_*)
        const char* e = getenv("O3D_PERSIST_STAGGER");
        int sh = e ? atoi(e) : ((peer && peer->nranks > 1) ? 4 : 0);
        const int last = a.nz - (f.nch - 1) * f.zchunk;
        if (f.nch < 2) sh = 0;
        if (sh > last - 2) sh = last - 2 > 0 ? last - 2 : 0;
        f.zstagger = sh;
    }
    f.first_src = first_src;
    f.max_iters = max_iters;
    f.eps = eps, f.factor = factor, f.kmax = kmax, f.idyn = idyn, f.fixed = fixed;
    const bool seams = a.seam_x || a.seam_y || a.seam_z;
    f.nxf = a.seam_x ? (long long)a.ny * a.nz : 0;
    f.nyf = a.seam_y ? (long long)a.nx * a.nz : 0;
    // the z seam is the last GLOBAL plane: only the rank that owns it sweeps it
    f.nzf = (a.seam_z && a.gz0 + a.nz == a.gnz) ? (long long)a.nx * a.ny : 0;
    if (peer) f.peer = *peer;
    else memset(&f.peer, 0, sizeof(f.peer));
    const long long items = (long long)ntiles * f.nch;
    const unsigned G = (unsigned)(items < G_max ? items : G_max);
    {
        // default: tickets.  CTAs that share an SM draw consecutive tickets, so neighbouring tiles
        // run on the same SM / GPC and the halo they share is found in L2; with the static
        // round-robin map (O3D_PERSIST_DYN=0) the same pass re-reads it from DRAM (measured at
        // 512^3: 3.6 GB instead of 2.4 GB per pass, profiles/r2_sor_scheduling.txt)
        const char* e = getenv("O3D_PERSIST_DYN");
        f.dynamic = (e && e[0] == '0') ? 0 : 1;
    }
    if (cudaMemsetAsync(sync, 0, 32 * sizeof(unsigned long long), st) != cudaSuccess) return 1;
    void* args[] = {(void*)&maps, (void*)&f, (void*)&ctrl, (void*)&sync};
    const bool mr = f.peer.nranks > 1;
    const void* fn = seams ? (mr ? (const void*)sor_persist_kernel<true, true>
                                 : (const void*)sor_persist_kernel<true, false>)
                           : (mr ? (const void*)sor_persist_kernel<false, true>
                                 : (const void*)sor_persist_kernel<false, false>);
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(G), dim3(GNT), args, GSMEM, st);
    count_launch();
    if (e != cudaSuccess) {
        set_error("persistent SOR launch failed: %s", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

}  // namespace o3d
