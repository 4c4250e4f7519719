// sor_kernels.cu -- pressure Poisson SOR (src/poisson.f90:6-381).
//
// The reference sweeps lexicographically (loop-carried dependence in i, j and k).  Two
// orderings are provided:
//   * RED_BLACK  (fast path): colour = (i+j+k) mod 2 with the GLOBAL k.  On an odd periodic
//     extent planes 0 and n-1 are same-colour neighbours, so the last plane of such an axis is a
//     "seam": points are further split by the parity of the number of seam planes they lie on.
//     Neighbouring points always differ in colour or in seam parity, so each of the (up to) four
//     classes is an independent set and the sweep is race-free and deterministic.
//   (pp lives in the padded layout of o3d_common.cuh, but the sweeps use the reference's
//   neighbour-index rule directly -- src/poisson.f90:57-92 -- because in-place updates would
//   otherwise have to keep ghost copies coherent between half-sweeps.)
//   * LEXI_WAVEFRONT (verification): hyperplanes i+j+k = h in ascending h; all points of a
//     hyperplane are independent and their neighbours are exactly as "old"/"new" as in the
//     lexicographic sweep, so the iterates are bit-identical to the reference's.
// Exit tests and the dynamic-omega rule (src/poisson.f90:110-122) run on the device in
// sor_control_kernel so the host only polls a flag every few iterations.
#include "kernels.h"
#include "sor_common.cuh"

namespace o3d {
namespace {

constexpr int SBX = 64, SBY = 4;

// bulk classes (seam parity 0)
__global__ void __launch_bounds__(SBX* SBY) sor_rb_kernel(const SorArgs a, int colour,
                                                          SorCtrl* ctrl, int zchunk) {
    __shared__ double red[32];
    if (*((volatile int*)&ctrl->done)) return;
    const double omega = *((volatile double*)&ctrl->omega);
    const int ii = blockIdx.x * SBX + threadIdx.x;
    const int j = blockIdx.y * SBY + threadIdx.y;
    const int kb = blockIdx.z * zchunk, ke = min(a.nz, kb + zchunk);
    double dmax = 0.0;
    if (j < a.ny) {
        for (int k = kb; k < ke; ++k) {
            const int gk = a.gz0 + k;
            const int i = 2 * ii + ((j + gk + colour) & 1);
            if (i < a.nx && !(seam_pop(a, i, j, gk) & 1))
                dmax = fmax(dmax, sor_point(a, i, j, k, omega));
        }
    }
    const double bm = block_max(dmax, red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
}

// seam classes (seam parity 1): the union of the seam planes, flattened
template <bool IMAGES>
__global__ void __launch_bounds__(256) sor_seam_kernel(const SorArgs a, int colour,
                                                       SorCtrl* ctrl, long long nxf,
                                                       long long nyf, long long nzf) {
    __shared__ double red[32];
    if (*((volatile int*)&ctrl->done)) return;
    const double omega = *((volatile double*)&ctrl->omega);
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double dmax = 0.0;
    int i = -1, j = -1, k = -1;
    bool ok = false;
    if (t < nxf) {  // x seam plane: (nx-1, j, k)
        i = a.nx - 1;
        j = (int)(t % a.ny);
        k = (int)(t / a.ny);
        ok = true;
    } else if ((t -= nxf) < nyf) {  // y seam plane, excluding points on the x seam
        j = a.ny - 1;
        i = (int)(t % a.nx);
        k = (int)(t / a.nx);
        ok = !(a.seam_x && i == a.nx - 1);
    } else if ((t -= nyf) < nzf) {  // z seam plane (owned by the last rank)
        k = a.nz - 1;
        i = (int)(t % a.nx);
        j = (int)(t / a.nx);
        ok = !(a.seam_x && i == a.nx - 1) && !(a.seam_y && j == a.ny - 1);
    }
    if (ok) {
        const int gk = a.gz0 + k;
        if (((i + j + gk) & 1) == colour && (seam_pop(a, i, j, gk) & 1))
            dmax = sor_point<IMAGES>(a, i, j, k, omega);
    }
    const double bm = block_max(dmax, red);
    if (threadIdx.x == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
}

// ------------------------------------------------------------------------------------------
// Fused red+black iteration: ONE pass over the grid per SOR iteration.
//
// A CTA owns a 32 x 8 column and marches in z.  Planes of the previous iterate live in a small
// shared-memory ring (36 x 12 cells: the tile plus a 2-cell halo).  At march step k the CTA
//   (1) red-updates plane k+1 over the tile plus a 1-cell ring (the ring cells are recomputed
//       redundantly instead of waiting for the neighbouring CTA -- same inputs, same arithmetic,
//       so bit-identical values), using only OLD black values of planes k, k+1, k+2;
//   (2) black-updates plane k over the tile, using only NEW red values of planes k-1, k, k+1;
//   (3) writes plane k of the new iterate.
// The iterate is ping-ponged between two buffers (read pp_old, write pp_new), so no CTA can
// observe a half-updated neighbour.  Result = exactly a red half-sweep followed by a black
// half-sweep, i.e. bit-identical to launching sor_rb_kernel twice, at
//   read pp (8) + read rhs (8) + write pp (8) = 24 B/pt per ITERATION
// instead of 48 B/pt for two in-place half-sweeps of colour-interleaved storage.
// Domain boundaries use the reference's neighbour rule as an index map when the halo cells are
// loaded (src/poisson.f90:57-92); requires 2-colourability (no odd periodic extent).
// Shared-memory planes are stored colour-split: cell (lx, ly) of the 36 x 20 staged plane sits
// at ly*36 + (lx&1)*18 + (lx>>1), so the cells of one colour in a row are contiguous (no bank
// conflicts) and a thread that owns the x-pair (2p, 2p+1) is active in both half-sweeps.
constexpr int FTX = 32, FTY = 16, FNT = 256;                  // tile; 16 x 16 threads own x-pairs
constexpr int FPX = FTX + 4, FPY = FTY + 4, FPL = FPX * FPY;  // staged plane, halo 2 (720 cells)
constexpr int FHALF = FPX / 2;                                // 18
constexpr int FNP = 5, FNR = 3;                               // ring slots: pp planes, rhs planes
constexpr int FNL = (FPL + FNT - 1) / FNT;                    // loader slots per thread (3)
constexpr int FRING = 2 * (FTX + 2) + 2 * FTY;                // 100 ring-1 cells

struct FusedArgs {
    const double* p_old;
    double* p_new;
    const double* rhs;
    double ox, oy, oz, invA;
    int mx, my, mz_lo, mz_hi;
    int nx, ny, nz;
    long long sy, sz;
    int gz0;
    int zchunk;
    int zmode, zlo, zhi, zedge;  // split launch, as MarchGeom
};

__device__ __forceinline__ int fmap(int q, int n, int mlo, int mhi) {
    bool refl;
    return map_index(q, n, mlo, mhi, refl);
}
__device__ __forceinline__ int fcell(int lx, int ly) {
    return ly * FPX + (lx & 1) * FHALF + (lx >> 1);
}

__global__ void __launch_bounds__(FNT, 3) sor_fused_kernel(const FusedArgs a, SorCtrl* ctrl) {
    __shared__ __align__(16) double sp[FNP][FPL];
    __shared__ __align__(16) double sr[FNR][FPL];
    __shared__ double red[32];
    if (*((volatile int*)&ctrl->done)) return;
    const double omega = *((volatile double*)&ctrl->omega);
    const double one_m_omega = 1.0 - omega;
    const int tid = threadIdx.x;
    const int px = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * FTX, j0 = blockIdx.y * FTY;
    int kb, ke;
    if (a.zmode == 2) {
        kb = (blockIdx.z == 0) ? 0 : a.nz - a.zedge;
        ke = kb + a.zedge;
    } else {
        kb = a.zlo + blockIdx.z * a.zchunk;
        ke = min(a.zhi, kb + a.zchunk);
    }

    // loader slots: global in-plane offset and shared index of the staged cells this thread
    // fetches (index map of src/poisson.f90:57-92 applied once, outside the march)
    long long poff[FNL];
    int pidx[FNL];
    bool rok[FNL];
#pragma unroll
    for (int q = 0; q < FNL; ++q) {
        const int c = tid + q * FNT;
        poff[q] = -1;
        pidx[q] = 0;
        rok[q] = false;
        if (c < FPL) {
            const int lx = c % FPX, ly = c / FPX;
            const int gi = i0 - 2 + lx, gj = j0 - 2 + ly;
            pidx[q] = fcell(lx, ly);
            if (gi < a.nx + 2 && gj < a.ny + 2) {
                poff[q] = fmap(gi, a.nx, a.mx, a.mx) + a.sy * fmap(gj, a.ny, a.my, a.my);
                rok[q] = lx >= 1 && lx <= FPX - 2 && ly >= 1 && ly <= FPY - 2;  // ring-1 region
            }
        }
    }
    // own x-pair (lx = 2+2px, 3+2px; ly = 2+ty): shared index of the EVEN member; the odd member
    // is FHALF further.  cpar: colour of the even member in local plane 0 (0 = red).
    const int own = (ty + 2) * FPX + 1 + px;
    const int gi = i0 + 2 * px, gj = j0 + ty;
    const bool in0 = gi < a.nx && gj < a.ny, in1 = gi + 1 < a.nx && gj < a.ny;
    const int cpar = (gi + gj + a.gz0) & 1;
    // one ring-1 cell for the first 100 threads
    int ring = -1, rpar = 0;
    if (tid < FRING) {
        int lx, ly;
        if (tid < FTX + 2) lx = 1 + tid, ly = 1;
        else if (tid < 2 * (FTX + 2)) lx = 1 + tid - (FTX + 2), ly = FPY - 2;
        else if (tid < 2 * (FTX + 2) + FTY) lx = 1, ly = 2 + tid - 2 * (FTX + 2);
        else lx = FPX - 2, ly = 2 + tid - 2 * (FTX + 2) - FTY;
        ring = fcell(lx, ly);
        rpar = (i0 - 2 + lx + j0 - 2 + ly + a.gz0) & 1;
        // neighbour offsets of a cell depend on which half it is in
        if (lx & 1) ring |= 0x10000;
    }

    double dmax = 0.0;
    // SOR update of the cell at shared index `cell` (odd: it sits in the odd half) of plane S
    auto update = [&](double* S, const double* Sm, const double* Sp, const double* Rr, int cell,
                      bool odd, bool store, bool count) -> double {
        const int other = odd ? cell - FHALF : cell + FHALF;  // same (lx>>1) in the other half
        const double west = odd ? S[other] : S[other - 1];
        const double east = odd ? S[other + 1] : S[other];
        const double pc = S[cell];
        // src/poisson.f90:95-98 with "/ A" as "* (1/A)" (see sor_point)
        const double p_new = (-(a.ox * (west + east)) - a.oy * (S[cell - FPX] + S[cell + FPX]) -
                              a.oz * (Sm[cell] + Sp[cell]) + Rr[cell]) * a.invA;
        const double v = one_m_omega * pc + omega * p_new;  // :102
        if (store) S[cell] = v;
        if (count) dmax = fmax(dmax, fabs(p_new - pc));  // :100
        return v;
    };

    int s_m1 = 0, s_0 = 1, s_p1 = 2, s_p2 = 3, s_free = 4;  // pp slots of planes k-1 .. k+2, free
    int r_0 = 0, r_p1 = 1, r_free = 2;                      // rhs slots of planes k, k+1, free
    double pv[FNL], rv[FNL];
    auto load = [&](int plane, bool want_p, bool want_r) {
        const long long z = a.sz * (long long)fmap(plane, a.nz, a.mz_lo, a.mz_hi);
#pragma unroll
        for (int q = 0; q < FNL; ++q) {
            if (want_p) pv[q] = (poff[q] >= 0) ? __ldg(a.p_old + z + poff[q]) : 0.0;
            if (want_r) rv[q] = rok[q] ? __ldg(a.rhs + z + poff[q]) : 0.0;
        }
    };
    auto stash = [&](double* d, const double* v) {
#pragma unroll
        for (int q = 0; q < FNL; ++q)
            if (tid + q * FNT < FPL) d[pidx[q]] = v[q];
    };
    auto red_plane = [&](int q, int sm, int s0, int sp1, int rs, bool owned) {
        // red = (i + j + global k) even.  Own pair: the even member is red iff cpar + q is even.
        const bool odd = (cpar + q) & 1;
        update(sp[s0], sp[sm], sp[sp1], sr[rs], own + (odd ? FHALF : 0), odd, true,
               owned && (odd ? in1 : in0));
        if (ring >= 0 && (((rpar + q) & 1) == 0))
            update(sp[s0], sp[sm], sp[sp1], sr[rs], ring & 0xffff, (ring >> 16) & 1, true, false);
    };

    // prologue: pp planes kb-2 .. kb+1, rhs planes kb-1, kb; red(kb-1), red(kb)
    load(kb - 2, true, false);
    stash(sp[4], pv);  // temporarily: plane kb-2 in the free slot
    load(kb - 1, true, true);
    stash(sp[s_m1], pv);
    stash(sr[r_free], rv);  // rhs(kb-1)
    load(kb, true, true);
    stash(sp[s_0], pv);
    stash(sr[r_0], rv);
    load(kb + 1, true, false);
    stash(sp[s_p1], pv);
    load(kb + 2, true, false);  // prefetch for the first march step
    {
        const long long z = a.sz * (long long)fmap(kb + 1, a.nz, a.mz_lo, a.mz_hi);
#pragma unroll
        for (int q = 0; q < FNL; ++q) rv[q] = rok[q] ? __ldg(a.rhs + z + poff[q]) : 0.0;
    }
    __syncthreads();
    red_plane(kb - 1, 4, s_m1, s_0, r_free, false);
    red_plane(kb, s_m1, s_0, s_p1, r_0, true);
    __syncthreads();  // red(kb-1) read slot 4 (plane kb-2): done before it is overwritten

    for (int k = kb; k < ke; ++k) {
        stash(sp[s_p2], pv);   // plane k+2
        stash(sr[r_p1], rv);   // rhs k+1
        if (k + 3 <= ke + 1) load(k + 3, true, false);
        if (k + 2 <= ke) {
            const long long z = a.sz * (long long)fmap(k + 2, a.nz, a.mz_lo, a.mz_hi);
#pragma unroll
            for (int q = 0; q < FNL; ++q) rv[q] = rok[q] ? __ldg(a.rhs + z + poff[q]) : 0.0;
        }
        __syncthreads();
        red_plane(k + 1, s_0, s_p1, s_p2, r_p1, k + 1 < ke);
        __syncthreads();
        {
            // black member of the own pair in plane k: all its neighbours are new red values
            const bool odd = ((cpar + k) & 1) == 0;  // even member red -> odd member is black
            const int bc = own + (odd ? FHALF : 0);
            const double vb = update(sp[s_0], sp[s_m1], sp[s_p1], sr[r_0], bc, odd, false,
                                     odd ? in1 : in0);
            const double vr = sp[s_0][own + (odd ? 0 : FHALF)];
            double* dst = a.p_new + (long long)k * a.sz + (long long)gj * a.sy + gi;
            const double v0 = odd ? vr : vb, v1 = odd ? vb : vr;
            if (in1) {
                *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            } else if (in0) {
                dst[0] = v0;
            }
        }
        // rotate the rings
        const int t = s_m1;
        s_m1 = s_0, s_0 = s_p1, s_p1 = s_p2, s_p2 = s_free, s_free = t;
        const int u = r_0;
        r_0 = r_p1, r_p1 = r_free, r_free = u;
    }
    const double bm = block_max(dmax, red);
    if (tid == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
}

// one hyperplane of the lexicographic sweep, reference arithmetic (division by A)
__global__ void __launch_bounds__(256) sor_wavefront_kernel(const SorArgs a, int h,
                                                            SorCtrl* ctrl) {
    __shared__ double red[32];
    if (*((volatile int*)&ctrl->done)) return;
    const double omega = *((volatile double*)&ctrl->omega);
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int k = blockIdx.y * 8 + threadIdx.y;
    double d = 0.0;
    const int i = h - j - k;
    if (j < a.ny && k < a.nz && i >= 0 && i < a.nx) {
        const long long sy = a.sy, sz = a.sz;
        int im1, ip1, jm1, jp1, km1, kp1;
        nbr_idx(i, a.nx, a.mx, a.mx, im1, ip1);
        nbr_idx(j, a.ny, a.my, a.my, jm1, jp1);
        nbr_idx(k, a.nz, a.mz_lo, a.mz_hi, km1, kp1);
        const long long m = (long long)k * sz + (long long)j * sy + i;
        const double pw = a.pp[(long long)k * sz + (long long)j * sy + im1];
        const double pe = a.pp[(long long)k * sz + (long long)j * sy + ip1];
        const double ps = a.pp[(long long)k * sz + (long long)jm1 * sy + i];
        const double pn = a.pp[(long long)k * sz + (long long)jp1 * sy + i];
        const double pb = a.pp[(long long)km1 * sz + (long long)j * sy + i];
        const double pt = a.pp[(long long)kp1 * sz + (long long)j * sy + i];
        const double pc = a.pp[m];
        const double p_new = sor_pnew_ref(a.oneondx2, a.oneondy2, a.oneondz2, pw, pe, ps, pn, pb,
                                          pt, a.rhs[m], a.A);  // src/poisson.f90:95-98
        d = fabs(p_new - pc);
        a.pp[m] = sor_relax(omega, pc, p_new);
    }
    const double bm = block_max(d, red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
}

__global__ void sor_control_kernel(SorCtrl* c, double eps, int kmax, int idyn, double factor) {
    sor_control_step(c, eps, kmax, idyn, factor);
}

// ------------------------------------------------------------------------------------------
// Both odd seam classes AND the end-of-iteration control in ONE launch (single rank): the seam
// sets are thin (~3/n of the grid), so three separate launches cost mostly launch latency --
// on the shipped 241 x 241 x 81 mixing layer (84 iterations per step) that is a third of the
// solve.  Cooperative launch (all CTAs co-resident), grid-stride loops, one hand-rolled grid
// barrier between the red and the black class; the last CTA to finish evaluates the exit tests
// and the dynamic omega (threadfence-reduction pattern).  sync[0]: barrier arrivals (monotone),
// sync[1]: finish arrivals (monotone).
__global__ void __launch_bounds__(256)
    sor_seam_fused_kernel(const SorArgs a, SorCtrl* ctrl, long long nxf, long long nyf,
                          long long nzf, unsigned long long* sync, double eps, int kmax, int idyn,
                          double factor) {
    __shared__ double red[32];
    __shared__ int last_s;
    if (*((volatile int*)&ctrl->done)) return;  // uniform over the grid: no CTA reaches a barrier
    const bool active = true;
    const double omega = *((volatile double*)&ctrl->omega);
    const long long tot = nxf + nyf + nzf;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double dmax = 0.0;
    for (int colour = 0; colour < 2; ++colour) {
        if (active) {
            for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot;
                 t += stride) {
                int i, j, k;
                bool ok;
                seam_point_of(a, t, nxf, nyf, nzf, i, j, k, ok);
                if (ok) {
                    const int gk = a.gz0 + k;
                    if (((i + j + gk) & 1) == colour && (seam_pop(a, i, j, gk) & 1))
                        dmax = fmax(dmax, sor_point<true>(a, i, j, k, omega));
                }
            }
        }
        if (colour == 0) {
            // grid barrier: the black class reads what the red class of other CTAs wrote
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                const unsigned long long t = atomicAdd(&sync[0], 1ull);
                const unsigned long long target = (t / gridDim.x + 1ull) * gridDim.x;
                while (*((volatile unsigned long long*)&sync[0]) < target) {
                }
                __threadfence();
            }
            __syncthreads();
        }
    }
    const double bm = block_max(dmax, red);
    if (threadIdx.x == 0) {
        if (bm > 0.0) atomic_max_nonneg(&ctrl->dmax_bits, bm);
        __threadfence();
        const unsigned long long t = atomicAdd(&sync[1], 1ull);
        last_s = (t % gridDim.x) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_s && threadIdx.x == 0) {
        __threadfence();
        sor_control_step(ctrl, eps, kmax, idyn, factor);
    }
}

}  // namespace

int launch_sor_rb(cudaStream_t st, const SorArgs& a, int colour, int seam_class, SorCtrl* ctrl,
                  int images) {
    if (seam_class == 0) {
        const int half = (a.nx + 1) / 2;
        const int gx = (half + SBX - 1) / SBX, gy = (a.ny + SBY - 1) / SBY;
        const int zchunk = pick_zchunk_slots(gx * gy, a.nz, 148 * 2, 0.5);
        sor_rb_kernel<<<dim3(gx, gy, (a.nz + zchunk - 1) / zchunk), dim3(SBX, SBY, 1), 0, st>>>(
            a, colour, ctrl, zchunk);
    } else {
        const long long nxf = a.seam_x ? (long long)a.ny * a.nz : 0;
        const long long nyf = a.seam_y ? (long long)a.nx * a.nz : 0;
        // the z seam is the last GLOBAL plane: only the rank that owns it sweeps it
        const bool own_z = a.seam_z && (a.gz0 + a.nz == a.gnz);
        const long long nzf = own_z ? (long long)a.nx * a.ny : 0;
        const long long tot = nxf + nyf + nzf;
        if (tot == 0) return 0;
        const unsigned nb = (unsigned)((tot + 255) / 256);
        if (images)
            sor_seam_kernel<true><<<nb, 256, 0, st>>>(a, colour, ctrl, nxf, nyf, nzf);
        else
            sor_seam_kernel<false><<<nb, 256, 0, st>>>(a, colour, ctrl, nxf, nyf, nzf);
    }
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_sor_seam_fused(cudaStream_t st, const SorArgs& a, SorCtrl* ctrl,
                          unsigned long long* sync, double eps, int kmax, int idyn,
                          double factor) {
    long long nxf = a.seam_x ? (long long)a.ny * a.nz : 0;
    long long nyf = a.seam_y ? (long long)a.nx * a.nz : 0;
    long long nzf = a.seam_z ? (long long)a.nx * a.ny : 0;
    const long long tot = nxf + nyf + nzf;
    // co-resident grid: at most 4 CTAs of 256 threads per SM
    static int max_ctas = 0;
    if (!max_ctas) {
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sor_seam_fused_kernel, 256, 0);
        if (per_sm > 4) per_sm = 4;
        max_ctas = per_sm * sms;
        if (max_ctas < 1) return 1;
    }
    long long nb = (tot + 255) / 256;
    if (nb > max_ctas) nb = max_ctas;
    if (nb < 1) nb = 1;
    SorArgs aa = a;
    void* args[] = {(void*)&aa,  (void*)&ctrl, (void*)&nxf,  (void*)&nyf,  (void*)&nzf,
                    (void*)&sync, (void*)&eps,  (void*)&kmax, (void*)&idyn, (void*)&factor};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)sor_seam_fused_kernel,
                                                      dim3((unsigned)nb), dim3(256), args, 0, st);
    count_launch();
    return e == cudaSuccess ? 0 : 1;
}

int launch_sor_fused(cudaStream_t st, const SorArgs& a, const double* p_old, double* p_new,
                     SorCtrl* ctrl, int zmode, int zedge) {
    FusedArgs f;
    f.p_old = p_old, f.p_new = p_new, f.rhs = a.rhs;
    f.ox = a.oneondx2, f.oy = a.oneondy2, f.oz = a.oneondz2, f.invA = a.invA;
    f.mx = a.mx, f.my = a.my, f.mz_lo = a.mz_lo, f.mz_hi = a.mz_hi;
    f.nx = a.nx, f.ny = a.ny, f.nz = a.nz;
    f.sy = a.sy, f.sz = a.sz;
    f.gz0 = a.gz0;
    const int gx = (a.nx + FTX - 1) / FTX, gy = (a.ny + FTY - 1) / FTY;
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
        // 4 extra planes per chunk: keep chunks long
        const int target = 148 * 6;
        int nch = (target + gx * gy - 1) / (gx * gy);
        int maxch = span / 48;
        if (maxch < 1) maxch = 1;
        if (nch > maxch) nch = maxch;
        if (nch < 1) nch = 1;
        f.zchunk = (span + nch - 1) / nch;
        gz = (span + f.zchunk - 1) / f.zchunk;
    }
    sor_fused_kernel<<<dim3(gx, gy, gz), FNT, 0, st>>>(f, ctrl);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_sor_control(cudaStream_t st, SorCtrl* ctrl, double eps, int kmax, int idyn,
                       double factor) {
    sor_control_kernel<<<1, 1, 0, st>>>(ctrl, eps, kmax, idyn, factor);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_sor_wavefront(cudaStream_t st, const SorArgs& a, int h, SorCtrl* ctrl) {
    sor_wavefront_kernel<<<dim3((a.ny + 31) / 32, (a.nz + 7) / 8), dim3(32, 8, 1), 0, st>>>(a, h,
                                                                                          ctrl);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
