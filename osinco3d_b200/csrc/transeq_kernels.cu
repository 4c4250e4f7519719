// transeq_kernels.cu -- passive scalar transport, reference transeq
// (src/integration.f90:332-468): 6 derivative sweeps + 7 temporaries + 3 global sums + 5
// element-wise passes there; here
//   pass 1 (TranseqEpi on the march engine): f1 = alpha_eff lap(phi) - u.grad(phi) + src, AB
//           update into a second phi buffer, and the three global sums of the conservative
//           clipping (sum(phi), sum(clip(phi)), sum(weight)) as per-CTA partials       72 B/pt
//   pass 2 (transeq_clip_kernel): clip, redistribute excess*weight/sum(weight), re-clip 16 B/pt
#include "kernels.h"
#include "march.cuh"

namespace o3d {
namespace {

struct TranseqEpi {
    static constexpr int STREAMS = 9;
    const double* u[3];
    const double* nu_t;
    const double* src;
    const double* f2;
    const double* f3;
    double* f1;
    double* phi_new;
    double* partial;
    Coef cx, cy, cz;
    double resc, sc, adu, bdu, cdu;
    int iles, sim2d;
    double s_old, s_clip, s_w;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    struct Pre {
        double u[3], nut, f2v, f3v, srcv;
    };
    __device__ __forceinline__ Pre prefetch(long long m, bool ok) const {
        Pre p;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.u[c] = ok ? __ldg(u[c] + m) : 0.0;
        p.nut = (ok && iles) ? __ldg(nu_t + m) : 0.0;
        p.f2v = ok ? __ldg(f2 + m) : 0.0;
        p.f3v = ok ? f3[m] : 0.0;  // may alias f1
        p.srcv = (ok && src) ? __ldg(src + m) : 0.0;
        return p;
    }
    __device__ __forceinline__ void apply(const Ring<1>& r, long long m, int, int, int,
                                          const Pre& p) {
        // phi is even along every axis (src/integration.f90:412-419)
        const double dx1 = r.d1x(0, cx), dy1 = r.d1y(0, cy);
        const double dx2 = r.d2x(0, cx), dy2 = r.d2y(0, cy);
        const double dz1 = sim2d ? 0.0 : r.d1z(0, cz);
        const double dz2 = sim2d ? 0.0 : r.d2z(0, cz);
        // src/integration.f90:403-409
        const double alpha_eff = transeq_alpha(resc, p.nut, sc, iles);
        // :422-423
        const double f = transeq_rhs_expr(alpha_eff, dx2, dy2, dz2, p.u[0], p.u[1], p.u[2], dx1, dy1,
                                          dz1, p.srcv);
        // :426
        const double pn = predictor_expr(r.c(0), adu, f, bdu, p.f2v, cdu, p.f3v);
        f1[m] = f;
        phi_new[m] = pn;
        // :433-444 partial sums
        s_old += pn;
        const double pc = clip01(pn);
        s_clip += pc;
        s_w += transeq_weight(pc);
    }
    __device__ __forceinline__ void finish(int tid, double* smem) {
        const double a = warp_sum(s_old), b = warp_sum(s_clip), c = warp_sum(s_w);
        if ((tid & 31) == 0) {
            smem[tid >> 5] = a;
            smem[8 + (tid >> 5)] = b;
            smem[16 + (tid >> 5)] = c;
        }
        __syncthreads();
        if (tid < 3) {
            double t = 0.0;
            for (int w = 0; w < MNT / 32; ++w) t += smem[tid * 8 + w];
            const int nblocks = gridDim.x * gridDim.y * gridDim.z;
            const int blk = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
            partial[(long long)tid * nblocks + blk] = t;
        }
    }
};

__global__ void __launch_bounds__(256) transeq_clip_kernel(const Geom g,
                                                            const double* __restrict__ phi_new,
                                                            double* __restrict__ phi,
                                                            const double* __restrict__ sums,
                                                            double count) {
    const double phi_old_avg = sums[0] / count;       // src/integration.f90:433
    const double phi_new_avg = sums[1] / count;       // :439
    const double excess = phi_old_avg - phi_new_avg;  // :440
    const double sw = sums[2];
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const long long base = (long long)k * g.sz + (long long)j * g.sy;
        for (int i = threadIdx.x; i < g.nx; i += blockDim.x) {
            const long long m = base + i;
            const double pc = clip01(__ldg(phi_new + m));       // :436
            phi[m] = transeq_redistribute(pc, excess, sw);      // :443-450
        }
    }
}

}  // namespace

int transeq_blocks(const Geom& g) {
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    const int zc = pick_zchunk(gx * gy, g.nz, 3, 1, TranseqEpi::STREAMS);  // as launch_march<1,0,1,TranseqEpi,3>
    return gx * gy * ((g.nz + zc - 1) / zc);
}

int launch_transeq_rhs(cudaStream_t st, const Geom& g, const TranseqArgs& a) {
    TranseqEpi e;
    for (int c = 0; c < 3; ++c) e.u[c] = a.u[c];
    e.nu_t = a.nu_t, e.src = a.src, e.f2 = a.f2, e.f3 = a.f3, e.f1 = a.f1;
    e.phi_new = a.phi_new, e.partial = a.partial;
    e.cx = a.cx, e.cy = a.cy, e.cz = a.cz;
    e.resc = a.resc, e.sc = a.sc, e.adu = a.adu, e.bdu = a.bdu, e.cdu = a.cdu;
    e.iles = a.iles, e.sim2d = g.sim2d;
    e.s_old = e.s_clip = e.s_w = 0.0;
    MarchMaps<1> m;
    m.m[0] = *a.phi.tm;
    return launch_march<1, 0, 1, TranseqEpi, 3>(st, g, m, e);
}

int launch_transeq_clip(cudaStream_t st, const Geom& g, const double* phi_new, double* phi,
                        const double* sums, double count) {
    long long b = (long long)g.ny * g.nz;
    if (b > 148 * 16) b = 148 * 16;
    transeq_clip_kernel<<<(unsigned)b, 256, 0, st>>>(g, phi_new, phi, sums, count);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
