// transeq_kernels.cu -- passive scalar transport, reference transeq
// (src/integration.f90:332-468): 6 derivative sweeps + 7 temporaries + 3 global sums + 5
// element-wise passes there; here
//   pass 1 (transeq_rhs_kernel): f1 = alpha_eff lap(phi) - u.grad(phi) + src, AB update into a
//           second phi buffer, and the three global sums of the conservative clipping
//           (sum(phi), sum(clip(phi)), sum(weight)) as per-CTA partials        72 B/pt
//   pass 2 (transeq_clip_kernel): clip, redistribute excess*weight/sum(weight), re-clip   16 B/pt
#include "kernels.h"
#include "stencil_tile.cuh"

namespace o3d {
namespace {

constexpr unsigned T_XMASK = 0x1, T_YMASK = 0x1;
static_assert(halo_slots(T_XMASK, T_YMASK) == 1, "one halo slot per thread");

__global__ void __launch_bounds__(NT, 3) transeq_rhs_kernel(const Dims g, const TranseqArgs a) {
    __shared__ double sm[2][SH * SW];
    __shared__ double red[3][NT / 32];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int kb = blockIdx.z * a.zchunk, ke = min(g.nz, kb + a.zchunk);
    const long long sz = (long long)g.nx * g.ny;

    const OwnCell oc = own_cell(g, i, j);  // phi is even along every axis (:412-419)
    HaloSlot hs = halo_slot<1>(g, i0, j0, tid, T_XMASK, T_YMASK);
    if (tid >= halo_cells(T_XMASK, T_YMASK)) hs.off = -1;

    double w[7];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        bool refl;
        const int pl = zplane(g, kb - R + m, refl);
        w[m] = oc.loadable ? __ldg(a.phi + (long long)pl * sz + oc.off) : 0.0;
    }
    double hreg = (hs.off >= 0) ? __ldg(a.phi + (long long)kb * sz + hs.off) : 0.0;
    const bool wallx = even_wall(i, g.nx, g.bx, g.bx, 0u, 0);
    const bool wally = even_wall(j, g.ny, g.by, g.by, 0u, 1);
    double s_old = 0.0, s_clip = 0.0, s_w = 0.0;

    for (int k = kb; k < ke; ++k) {
        const int buf = (k - kb) & 1;
        {
            bool refl;
            const int pl = zplane(g, k + R, refl);
            w[6] = oc.loadable ? __ldg(a.phi + (long long)pl * sz + oc.off) : 0.0;
        }
        double hn = 0.0;
        if (k + 1 < ke && hs.off >= 0) hn = __ldg(a.phi + (long long)(k + 1) * sz + hs.off);
        const long long m = (long long)k * sz + (long long)j * g.nx + i;
        double uv[3] = {0., 0., 0.}, nut = 0.0, f2v = 0.0, f3v = 0.0, srcv = 0.0;
        if (oc.in_dom) {
#pragma unroll
            for (int c = 0; c < 3; ++c) uv[c] = __ldg(a.u[c] + m);
            if (a.iles) nut = __ldg(a.nu_t + m);
            f2v = __ldg(a.f2 + m);
            f3v = a.f3[m];  // may alias f1
            if (a.src) srcv = __ldg(a.src + m);
        }
        if (oc.loadable) sm[buf][(ty + R) * SW + tx + R] = w[3];
        if (hs.off >= 0) sm[buf][hs.sm] = hreg;
        __syncthreads();
        if (oc.in_dom) {
            const double* t = &sm[buf][(ty + R) * SW + tx + R];
            const double f0 = w[3];
            const double dx1 = wallx ? 0.0
                                     : d1_expr(a.cx.a1, a.cx.b1, a.cx.c1, t[-3], t[-2], t[-1],
                                               t[1], t[2], t[3]);
            const double dy1 = wally ? 0.0
                                     : d1_expr(a.cy.a1, a.cy.b1, a.cy.c1, t[-3 * SW], t[-2 * SW],
                                               t[-SW], t[SW], t[2 * SW], t[3 * SW]);
            const double dx2 = d2_expr(a.cx.a2, a.cx.b2, a.cx.c2, t[-2], t[-1], f0, t[1], t[2]);
            const double dy2 = d2_expr(a.cy.a2, a.cy.b2, a.cy.c2, t[-2 * SW], t[-SW], f0, t[SW],
                                       t[2 * SW]);
            double dz1, dz2;
            if (g.sim2d) {
                dz1 = 0.0, dz2 = 0.0;
            } else {
                const bool wallz = (k == 0 && g.bz_lo == BM_MIRROR) ||
                                   (k == g.nz - 1 && g.bz_hi == BM_MIRROR);
                dz1 = wallz ? 0.0
                            : d1_expr(a.cz.a1, a.cz.b1, a.cz.c1, w[0], w[1], w[2], w[4], w[5],
                                      w[6]);
                dz2 = d2_expr(a.cz.a2, a.cz.b2, a.cz.c2, w[1], w[2], f0, w[4], w[5]);
            }
            // src/integration.f90:403-409
            const double alpha_eff = a.iles ? (1.0 / a.resc + nut / a.sc) : (1.0 / a.resc);
            // :422-423
            const double f1 = alpha_eff * (dx2 + dy2 + dz2) -
                              (uv[0] * dx1 + uv[1] * dy1 + uv[2] * dz1) + srcv;
            // :426
            const double pn = f0 + a.adu * f1 + a.bdu * f2v + a.cdu * f3v;
            a.f1[m] = f1;
            a.phi_new[m] = pn;
            // :433-444 partial sums
            s_old += pn;
            const double pc = fmax(0.0, fmin(1.0, pn));
            s_clip += pc;
            s_w += fmin(pc, 1.0 - pc);
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) w[q] = w[q + 1];
        hreg = hn;
    }
    s_old = warp_sum(s_old);
    s_clip = warp_sum(s_clip);
    s_w = warp_sum(s_w);
    if ((tid & 31) == 0) red[0][tid >> 5] = s_old, red[1][tid >> 5] = s_clip, red[2][tid >> 5] = s_w;
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
        for (int q = 0; q < NT / 32; ++q) t += red[tid][q];
        const int nblocks = gridDim.x * gridDim.y * gridDim.z;
        const int b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        a.partial[(long long)tid * nblocks + b] = t;
    }
}

__global__ void __launch_bounds__(256) transeq_clip_kernel(long long n,
                                                            const double* __restrict__ phi_new,
                                                            double* __restrict__ phi,
                                                            const double* __restrict__ sums,
                                                            double count) {
    const double phi_old_avg = sums[0] / count;   // src/integration.f90:433
    const double phi_new_avg = sums[1] / count;   // :439
    const double excess = phi_old_avg - phi_new_avg;  // :440
    const double sw = sums[2];
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < n;
         m += (long long)gridDim.x * blockDim.x) {
        const double pc = fmax(0.0, fmin(1.0, __ldg(phi_new + m)));  // :436
        double wgt = fmin(pc, 1.0 - pc);                             // :443
        wgt = wgt / sw;                                              // :444
        double p = pc + excess * wgt;                                // :447
        p = fmax(0.0, fmin(1.0, p));                                 // :450
        phi[m] = p;
    }
}

}  // namespace

int transeq_blocks(const Dims& g) {
    const int gx = (g.nx + TX - 1) / TX, gy = (g.ny + TY - 1) / TY;
    const int zc = pick_zchunk(gx * gy, g.nz);
    return gx * gy * ((g.nz + zc - 1) / zc);
}

int launch_transeq_rhs(cudaStream_t st, const Dims& g, const TranseqArgs& a_in) {
    TranseqArgs a = a_in;
    const int gx = (g.nx + TX - 1) / TX, gy = (g.ny + TY - 1) / TY;
    a.zchunk = pick_zchunk(gx * gy, g.nz);
    transeq_rhs_kernel<<<dim3(gx, gy, (g.nz + a.zchunk - 1) / a.zchunk), dim3(TX, TY, 1), 0, st>>>(
        g, a);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_transeq_clip(cudaStream_t st, long long n, const double* phi_new, double* phi,
                        const double* sums, double count) {
    long long b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    transeq_clip_kernel<<<(unsigned)b, 256, 0, st>>>(n, phi_new, phi, sums, count);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
