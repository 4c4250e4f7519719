// sor_tma_kernel.cu -- fused red+black SOR iteration, TMA-staged (the fast path of the pressure
// Poisson solve, src/poisson.f90:6-381, for 2-colourable grids).
//
// One pass over the grid per SOR iteration, ping-pong p_old -> p_new:
//   read p (8) + read rhs (8) + write p (8) = 24 B/pt per ITERATION
// (the reference streams pp, rhs and p_new: 32 B/pt, with a loop-carried dependence).
//
// The boundary rule of src/poisson.f90:57-92 (periodic wrap / mirror) lives in the GHOST CELLS
// of p_old and rhs (o3d_common.cuh): this kernel writes the ghost images of every point it
// stores (faces and edges), so the next pass -- and the projection correction after the last
// one -- find their closure in place and a plane is a plain rectangular TMA box for interior
// and boundary tiles alike: no index maps, no boundary branches.  A ghost cell's red update is
// recomputed from ghost data with the same arithmetic as the interior point it mirrors, so it is
// bit-identical to it.
//
// A CTA (256 threads, each owning a y-pair = one red + one black cell per plane) owns a 32 x 16
// column and marches in z.  Planes are staged with a 2-cell halo (36 x 20 boxes) in a ring of 7
// (p) + 5 (rhs) shared-memory stages, two planes prefetched ahead.  At march step k:
//   red(k+2): tile + its 1-cell ring (ring cells recomputed redundantly instead of waiting for
//             the neighbouring CTA: same inputs, same arithmetic -> same bits), from the OLD
//             black values of planes k+1, k+2, k+3;
//   black(k): tile, from the NEW red values of planes k-1, k, k+1; plane k of p_new is stored.
// The two phases touch disjoint data, so ONE __syncthreads per plane suffices.
// Result == a red half-sweep followed by a black half-sweep (sor_rb_kernel twice), bit for bit.
//
// Odd periodic extents (SEAM = true; every shipped periodic grid: 257 x 513 x 129, 241 x 241 x 81).
// Planes 0 and n-1 of such an axis are same-colour neighbours, so the grid is not 2-colourable;
// sor_kernels.cu splits the points into four independent classes (colour, parity of the number
// of seam planes n-1 the point lies on) swept in the order (red, even) (black, even) (red, odd)
// (black, odd).  This kernel does the first two -- all but ~3/n of the points -- in its single
// pass and copies the odd-parity points through unchanged; sor_seam_kernel then sweeps the two
// thin odd classes in place on p_new.  Why one pass still works: an even-parity point's
// neighbour ACROSS a seam always has odd parity, i.e. still holds its old value for the whole
// pass, which is exactly what the ghost cells of p_old on that axis contain; so those ghost
// cells (and the odd-parity ring cells) are simply never red-updated here.
#include "kernels.h"
#include "tma.cuh"

namespace o3d {
namespace {

constexpr int GTX = 32, GTY = 16, GNT = 256;
constexpr int GBX = GTX + 4, GBY = GTY + 4, GPL = GBX * GBY;  // 36 x 20 = 720 cells, 5760 B
constexpr int GP = 2;                                         // planes prefetched ahead
constexpr int GNP = 5 + GP, GNR = 3 + GP, GNB = GP + 1;       // p stages, rhs stages, barriers
constexpr int GSMEM = (GNP + GNR) * GPL * 8 + GNB * 8 + 32 * 8;

struct TmaSorArgs {
    double* p_new;
    double ox, oy, oz, invA;
    int nx, ny, nz;
    long long sy, sz;
    int bx, by, bz_lo, bz_hi;  // closures (BM_*), for the ghost images of p_new
    int gz0, gnz;
    int opx, opy, opz;  // odd periodic extent along the axis: plane n-1 is a seam (SEAM kernels)
    int zchunk, zmode, zlo, zhi, zedge;
};

struct alignas(64) TmaSorMaps {
    CUtensorMap p, rhs;
};

// shared-memory accesses as [register + immediate] through the 32-bit shared window
template <int OFF>
__device__ __forceinline__ double lds(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ void sts(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

template <bool SEAM>
__global__ void __launch_bounds__(GNT, 3)
    sor_tma_kernel(const __grid_constant__ TmaSorMaps maps, const TmaSorArgs a, SorCtrl* ctrl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sp = reinterpret_cast<double*>(smem_raw);
    double* sr = sp + GNP * GPL;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sr + GNR * GPL);
    double* red = reinterpret_cast<double*>(bars + GNB);
    if (*((volatile int*)&ctrl->done)) return;
    const double omega = *((volatile double*)&ctrl->omega);
    const double one_m_omega = 1.0 - omega;

    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * GTX, j0 = blockIdx.y * GTY;
    int kb, ke;
    if (a.zmode == 2) {
        kb = (blockIdx.z == 0) ? 0 : a.nz - a.zedge;
        ke = kb + a.zedge;
    } else {
        kb = a.zlo + blockIdx.z * a.zchunk;
        ke = min(a.zhi, kb + a.zchunk);
    }
    const int niter = ke - kb;
    const int ngroups = niter > 1 ? niter - 1 : 1;  // group n feeds red(kb + n + 2)

    const uint32_t sp_s = smem_u32(sp), sr_s = smem_u32(sr), bars_s = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GNB; ++s) mbar_init(bars_s + 8 * s, 1);
        fence_barrier_init();
    }
    __syncthreads();

    // p plane q lives in stage (q - (kb-2)) mod GNP, rhs plane q in (q - (kb-1)) mod GNR
    const int cx = GX + i0 - 2, cy = GH + j0 - 2;
    auto issue_p = [&](int plane, uint32_t bar) {
        const unsigned st = (unsigned)(plane - (kb - 2)) % GNP;
        tma_load_3d(sp_s + st * (GPL * 8), &maps.p, bar, cx, cy, GH + plane);
    };
    auto issue_r = [&](int plane, uint32_t bar) {
        const unsigned st = (unsigned)(plane - (kb - 1)) % GNR;
        tma_load_3d(sr_s + st * (GPL * 8), &maps.rhs, bar, cx, cy, GH + plane);
    };
    auto issue_group = [&](int n) {
        const uint32_t bar = bars_s + 8 * (n % GNB);
        if (n == 0) {
            mbar_expect_tx(bar, (6 + 4) * GPL * 8);
#pragma unroll
            for (int q = -2; q <= 3; ++q) issue_p(kb + q, bar);
#pragma unroll
            for (int q = -1; q <= 2; ++q) issue_r(kb + q, bar);
        } else {
            mbar_expect_tx(bar, 2 * GPL * 8);
            issue_p(kb + n + 3, bar);
            issue_r(kb + n + 2, bar);
        }
    };
    if (tid == 0) {
        for (int n = 0; n < GP && n < ngroups; ++n) issue_group(n);
    }

    // Own Y-PAIR: column 2+tx, rows 2+2*typ (member A) and 3+2*typ (member B): one red and one
    // black cell in every plane.  A warp's lanes address 32 consecutive doubles of two adjacent
    // rows in a checkerboard; with the even row pitch (36) the two rows fall on complementary
    // banks, so every shared-memory access of the sweep is conflict-free, and the global stores
    // are two fully coalesced rows.  pe = colour of member A in global plane 0.
    const int tx = tid & 31, typ = tid >> 5;
    const int own = (2 + 2 * typ) * GBX + 2 + tx;
    const int gi = i0 + tx, gj = j0 + 2 * typ;
    const bool inA = gi < a.nx && gj < a.ny, inB = gi < a.nx && gj + 1 < a.ny;
    const int pe = (gi + gj + a.gz0) & 1;
    // ring-1 pairs (48 threads): each pair holds exactly one red cell in every plane
    int rcell = -1, rstep = 0, rpar = 0;
    if (tid < 48) {
        int lx, ly;
        if (tid < 16) lx = 2 + 2 * tid, ly = 1, rstep = 1;                       // below the tile
        else if (tid < 32) lx = 2 + 2 * (tid - 16), ly = GBY - 2, rstep = 1;     // above
        else if (tid < 40) lx = 1, ly = 2 + 2 * (tid - 32), rstep = GBX;         // left
        else lx = GBX - 2, ly = 2 + 2 * (tid - 40), rstep = GBX;                 // right
        rcell = ly * GBX + lx;
        rpar = (i0 - 2 + lx + j0 - 2 + ly + a.gz0) & 1;
    }
    // SEAM: seam parity of the own pair in x,y (bit 0: member A, bit 1: member B) and of the two
    // cells of the ring pair (2 bits each: 0 / 1 = parity, 2 = ghost cell of an odd periodic axis,
    // never updated)
    int own_par = 0, ring_code = 0;
    if (SEAM) {
        const int sx = (a.opx && gi == a.nx - 1) ? 1 : 0;
        own_par = (sx ^ ((a.opy && gj == a.ny - 1) ? 1 : 0)) |
                  ((sx ^ ((a.opy && gj + 1 == a.ny - 1) ? 1 : 0)) << 1);
        if (rcell >= 0) {
            for (int m = 0; m < 2; ++m) {
                const int c = rcell + m * rstep;
                const int ri = i0 - 2 + c % GBX, rj = j0 - 2 + c / GBX;
                int code = 0;
                if (a.opx) code = (ri < 0 || ri >= a.nx) ? 2 : (ri == a.nx - 1);
                if (a.opy && code != 2)
                    code = (rj < 0 || rj >= a.ny) ? 2 : (code ^ (rj == a.ny - 1 ? 1 : 0));
                ring_code |= code << (2 * m);
            }
        }
    }
    // seam state of plane q: 0 / 1 = parity contribution, 2 = ghost plane of an odd periodic z
    auto zseam = [&](int q) -> int {
        if (!SEAM || !a.opz) return 0;
        const int gk = a.gz0 + q;
        return (gk < 0 || gk >= a.gnz) ? 2 : (gk == a.gnz - 1 ? 1 : 0);
    };
    // ghost images of the own points in x and y
    const Img2 ix = image_offsets(gi, a.nx, a.bx, a.bx);
    const Img2 iyA = image_offsets(gj, a.ny, a.by, a.by);
    const Img2 iyB = image_offsets(gj + 1, a.ny, a.by, a.by);
    const bool xy_img = (ix.lo | ix.hi | iyA.lo | iyA.hi | iyB.lo | iyB.hi) != 0;
    // planes whose points have z images (walls / periodic wrap handled by this rank)
    const int zimg_lo = (a.bz_lo == BM_MIRROR || a.bz_hi == BM_WRAP) ? R : -1;      // k <= zimg_lo
    const int zimg_hi = (a.bz_hi == BM_MIRROR || a.bz_lo == BM_WRAP) ? a.nz - 1 - R : a.nz;

    double dmax = 0.0;
    // ---- hot loop -------------------------------------------------------------------------
    // The pass is issue-bound (64 % of the issue slots at 3 CTAs/SM), so the loop is written for
    // instruction count: staged planes are addressed through 32-bit shared-window byte addresses
    // kept in registers and rotated (no index -> pointer arithmetic per access), every load is
    // ld.shared [reg + immediate], and the colour of the pair's members alternates from plane to
    // plane by XOR-ing precomputed offsets.
    constexpr uint32_t PLB = GPL * 8;  // bytes per staged plane
    const uint32_t p_end = sp_s + GNP * PLB, r_end = sr_s + GNR * PLB;
    auto nextp = [&](uint32_t x) { return x + PLB == p_end ? sp_s : x + PLB; };
    auto nextr = [&](uint32_t x) { return x + PLB == r_end ? sr_s : x + PLB; };
    // SOR update of the cell at byte offset c of the plane at a0 (am / ap = planes below / above,
    // ar = rhs plane): src/poisson.f90:95-102 with "/ A" as "* (1/A)" (this ordering is not the
    // bit-parity one); returns the relaxed value, d = |p_new - p_old|, pc = the old value
    auto update = [&](uint32_t am, uint32_t a0, uint32_t ap, uint32_t ar, uint32_t c, double& d,
                      double& pc) -> double {
        const uint32_t c0 = a0 + c;
        pc = lds<0>(c0);
        const double w = lds<-8>(c0), e = lds<8>(c0);
        const double sn = lds<-GBX * 8>(c0), nn = lds<GBX * 8>(c0);
        const double bb = lds<0>(am + c), tt = lds<0>(ap + c), rr = lds<0>(ar + c);
        const double p_new =
            (-(a.ox * (w + e)) - a.oy * (sn + nn) - a.oz * (bb + tt) + rr) * a.invA;
        d = fabs(p_new - pc);                     // :100
        return one_m_omega * pc + omega * p_new;  // :102
    };
    const uint32_t ownA8 = own * 8, ownB8 = (own + GBX) * 8, cxor = ownA8 ^ ownB8;
    const bool has_ring = rcell >= 0;  // warps 0 and 1 only
    const uint32_t ring0 = has_ring ? rcell * 8 : 0, ring1 = has_ring ? (rcell + rstep) * 8 : 0;
    const uint32_t rxor = ring0 ^ ring1;
    // r = 1: member B of the pair is the red one in the current plane (member A otherwise);
    // c_red / c_blk = byte offsets of the red / black member; rc = the red cell of the ring pair,
    // rmm = which of its two cells that is.  All flip from plane to plane.
    int r = (pe + kb) & 1, rmm = (rpar + kb) & 1;
    uint32_t c_red = r ? ownB8 : ownA8, c_blk = r ? ownA8 : ownB8;
    uint32_t rc = rmm ? ring1 : ring0;
    auto flip = [&]() { r ^= 1, rmm ^= 1, c_red ^= cxor, c_blk ^= cxor, rc ^= rxor; };
    // red half-sweep of the plane at a0 (global plane q) over the own pair and the ring pair;
    // counted: the plane is owned by this chunk
    auto red_plane = [&](int q, uint32_t am, uint32_t a0, uint32_t ap, uint32_t ar, bool counted) {
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
        // (a red cell has no red neighbour: every read above is of a black cell or of the cell's
        // own centre, every write of a red cell owned by exactly one thread)
        sts(a0 + c_red, v);
        if (counted && (r ? inB : inA)) dmax = fmax(dmax, d);
    };

    mbar_wait(bars_s, 0);  // group 0: p planes kb-2 .. kb+3 in stages 0 .. 5, rhs kb-1 .. kb+2 in 0 .. 3
    flip();  // plane kb-1 has the other parity
    red_plane(kb - 1, sp_s, sp_s + PLB, sp_s + 2 * PLB, sr_s, false);
    flip();
    red_plane(kb, sp_s + PLB, sp_s + 2 * PLB, sp_s + 3 * PLB, sr_s + PLB, true);
    flip();
    red_plane(kb + 1, sp_s + 2 * PLB, sp_s + 3 * PLB, sp_s + 4 * PLB, sr_s + 2 * PLB, kb + 1 < ke);
    flip();  // back to the parity of plane kb

    // addresses of the staged planes k-1 .. k+3 (p) and k, k+2 (rhs) for k = kb
    uint32_t a_m1 = sp_s + PLB, a_0 = sp_s + 2 * PLB, a_1 = sp_s + 3 * PLB, a_2 = sp_s + 4 * PLB,
             a_3 = sp_s + 5 * PLB;
    uint32_t ar_0 = sr_s + PLB, ar_2 = sr_s + 3 * PLB;
    double* outp = a.p_new + (long long)kb * a.sz + (long long)gj * a.sy + gi;
    int bi = 1 % GNB;  // barrier index / phase of group n = 1
    uint32_t bpar = 0;
    for (int k = kb; k < ke; ++k) {
        const int n = k - kb;
        __syncthreads();  // step k-1 done: its oldest stages may be refilled; red(k+1) visible
        if (tid == 0 && n + GP < ngroups) issue_group(n + GP);
        if (n >= 1) {
            if (n < ngroups) mbar_wait(bars_s + 8 * bi, bpar);
            if (++bi == GNB) bi = 0, bpar ^= 1u;
        }
        // planes k+2 and k have the same parity: the same member is red in both
        if (k + 2 <= ke) red_plane(k + 2, a_1, a_2, a_3, ar_2, k + 2 < ke);
        {
            // black member of the own pair in plane k: all six neighbours hold new red values
            double d, pcb;
            double vb = update(a_m1, a_0, a_1, ar_0, c_blk, d, pcb);
            if (SEAM && (((own_par >> (1 - r)) ^ zseam(k)) & 1))
                vb = pcb, d = 0.0;  // odd class: swept by sor_seam_kernel
            const double vred = lds<0>(a_0 + c_red);
            if (r ? inA : inB) dmax = fmax(dmax, d);
            const double vA = r ? vb : vred, vB = r ? vred : vb;
            if (inA) outp[0] = vA;
            if (inB) outp[a.sy] = vB;
            if (xy_img || k <= zimg_lo || k >= zimg_hi) {  // boundary-adjacent points only
                const Img2 iz = image_offsets(k, a.nz, a.bz_lo, a.bz_hi);
                if (inA)
                    store_images(outp, 0, vA, ix, iyA.lo * a.sy, iyA.hi * a.sy, iz.lo * a.sz,
                                 iz.hi * a.sz);
                if (inB)
                    store_images(outp, a.sy, vB, ix, iyB.lo * a.sy, iyB.hi * a.sy, iz.lo * a.sz,
                                 iz.hi * a.sz);
            }
        }
        outp += a.sz;
        flip();
        a_m1 = a_0, a_0 = a_1, a_1 = a_2, a_2 = a_3, a_3 = nextp(a_3);
        ar_0 = nextr(ar_0), ar_2 = nextr(ar_2);
    }
    const double bm = block_max(dmax, red);
    if (tid == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
}

}  // namespace

int sor_tma_box_x() { return GBX; }
int sor_tma_box_y() { return GBY; }

int launch_sor_tma(cudaStream_t st, const SorArgs& a, const CUtensorMap* p_old_map,
                   const CUtensorMap* rhs_map, double* p_new, int bx, int by, int bz_lo,
                   int bz_hi, SorCtrl* ctrl, int zmode, int zedge) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(sor_tma_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 GSMEM) != cudaSuccess ||
            cudaFuncSetAttribute(sor_tma_kernel<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 GSMEM) != cudaSuccess)
            return 1;
        attr_set = true;
    }
    TmaSorMaps maps;
    maps.p = *p_old_map, maps.rhs = *rhs_map;
    TmaSorArgs f;
    f.p_new = p_new;
    f.ox = a.oneondx2, f.oy = a.oneondy2, f.oz = a.oneondz2, f.invA = a.invA;
    f.nx = a.nx, f.ny = a.ny, f.nz = a.nz;
    f.sy = a.sy, f.sz = a.sz;
    f.bx = bx, f.by = by, f.bz_lo = bz_lo, f.bz_hi = bz_hi;
    f.gz0 = a.gz0, f.gnz = a.gnz;
    f.opx = a.seam_x, f.opy = a.seam_y, f.opz = a.seam_z;
    const int gx = (a.nx + GTX - 1) / GTX, gy = (a.ny + GTY - 1) / GTY;
    f.zmode = zmode, f.zedge = zedge;
    f.zlo = (zmode == 1) ? zedge : 0;
    f.zhi = (zmode == 1) ? a.nz - zedge : a.nz;
    int gz;
    if (zmode == 2) {
        f.zchunk = zedge;
        gz = 2;
    } else {
        const int span = f.zhi - f.zlo;
        if (span <= 0) return 0;
        // per chunk: 5 prologue planes of 2 of the 3 streams (+ their red updates), 3 CTAs per SM
        f.zchunk = pick_zchunk_slots(gx * gy, span, 148 * 3, 3.3);
        gz = (span + f.zchunk - 1) / f.zchunk;
    }
    if (f.opx || f.opy || f.opz)
        sor_tma_kernel<true><<<dim3(gx, gy, gz), GNT, GSMEM, st>>>(maps, f, ctrl);
    else
        sor_tma_kernel<false><<<dim3(gx, gy, gz), GNT, GSMEM, st>>>(maps, f, ctrl);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
