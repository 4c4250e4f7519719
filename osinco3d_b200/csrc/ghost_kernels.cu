// ghost_kernels.cu -- maintenance of the padded field layout (o3d_common.cuh):
//   fill_ghosts : write the boundary closure of src/derivation.f90 (periodic wrap, even / odd
//                 mirror) into the 3 ghost layers of up to 6 fields in one launch.  Only the
//                 face ghosts are filled: no stencil of the reference has mixed derivatives.
//   pack/unpack : contiguous Fortran array (nx,ny,nz) <-> padded interior (H2D / D2H staging).
// Surface work only: ~6 * 3 * n^2 cells per field against n^3 for a stencil pass.
#include "kernels.h"

namespace o3d {
namespace {

__global__ void __launch_bounds__(256) fill_ghosts_kernel(const Geom g, const GhostArgs a) {
    const int job = blockIdx.z / 3, axis = blockIdx.z % 3;
    if (job >= a.njobs) return;
    const GhostJob jb = a.job[job];
    if (!((jb.axes >> axis) & 1u)) return;
    const bool odd = (jb.par >> axis) & 1u;
    const bool dirichlet = (jb.par >> (axis + 4)) & 1u;  // GHOST_PAR_DIRICHLET(axis)
    double* __restrict__ p = jb.p;
    const int n = (axis == 0) ? g.nx : (axis == 1) ? g.ny : g.nz;
    const long long s = (axis == 0) ? 1 : (axis == 1) ? g.sy : g.sz;
    const int mlo = (axis == 0) ? g.bx : (axis == 1) ? g.by : g.bz_lo;
    const int mhi = (axis == 0) ? g.bx : (axis == 1) ? g.by : g.bz_hi;
    // the two other extents (a = fast, b = slow)
    const int na = (axis == 0) ? g.ny : g.nx;
    // x / y ghosts of the planes [zr_lo, zr_hi) only, when the geometry carries a plane range
    const bool ranged = (axis != 2) && (g.zr_hi > g.zr_lo);
    const int b0 = ranged ? g.zr_lo : 0;
    const int nb = (axis == 2) ? g.ny : (ranged ? g.zr_hi - g.zr_lo : g.nz);
    const long long sa = (axis == 0) ? g.sy : 1;
    const long long sb = (axis == 2) ? g.sy : g.sz;
    const long long total = (long long)na * nb * 6;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int ia, ib, g6;
        if (axis == 0) {  // 6 ghost cells of one row are handled by 6 neighbouring threads
            g6 = (int)(t % 6);
            const long long row = t / 6;
            ia = (int)(row % na), ib = (int)(row / na);
        } else {  // i fastest: coalesced rows
            ia = (int)(t % na);
            const long long rest = t / na;
            g6 = (int)(rest % 6), ib = (int)(rest / 6);
        }
        ib += b0;
        const int side = g6 / 3, gg = g6 % 3 + 1;
        const int mode = side ? mhi : mlo;
        if (mode == BM_HALO) continue;  // filled by the z-slab halo exchange
        // z ghosts of one side only (GHOST_Z_LO_ONLY / GHOST_Z_HI_ONLY)
        if (axis == 2 && (jb.axes & (side ? GHOST_Z_LO_ONLY : GHOST_Z_HI_ONLY))) continue;
        const int q = side ? (n - 1 + gg) : -gg;
        bool refl;
        const int src = map_index(q, n, mlo, mhi, refl);
        const long long base = (long long)ia * sa + (long long)ib * sb;
        const double v = p[base + (long long)src * s];
        double gv = (refl && odd) ? -v : v;
        if (refl && dirichlet) {
            // Dirichlet wall (NEW closure, no counterpart in the reference): the stored boundary
            // plane holds the prescribed value and the field is continued by odd reflection ABOUT
            // that value, f(-g) = 2 f_wall - f(+g); der?i_11 is its f_wall = 0 special case
            const double wall = p[base + (long long)(side ? n - 1 : 0) * s];
            gv = 2.0 * wall - v;
        }
        p[base + (long long)q * s] = gv;
    }
}

// One axis of the FULL closure (faces, edges, corners): the ghost cells of `axis` over the other
// two extents widened by `ea` / `eb` ghost cells, so that x, then y (over the x ghosts), then z
// (over the x and y ghosts) produces the images of images the SOR kernel's halo reads.
__global__ void __launch_bounds__(256) fill_axis_kernel(const Geom g, double* __restrict__ p,
                                                        int axis, int odd, int ea, int eb) {
    const int n = (axis == 0) ? g.nx : (axis == 1) ? g.ny : g.nz;
    const long long s = (axis == 0) ? 1 : (axis == 1) ? g.sy : g.sz;
    const int mlo = (axis == 0) ? g.bx : (axis == 1) ? g.by : g.bz_lo;
    const int mhi = (axis == 0) ? g.bx : (axis == 1) ? g.by : g.bz_hi;
    const int na = ((axis == 0) ? g.ny : g.nx) + 2 * ea;
    const int nb = ((axis == 2) ? g.ny : g.nz) + 2 * eb;
    const long long sa = (axis == 0) ? g.sy : 1;
    const long long sb = (axis == 2) ? g.sy : g.sz;
    const long long total = (long long)na * nb * 6;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int ia, ib, g6;
        if (axis == 0) {
            g6 = (int)(t % 6);
            const long long row = t / 6;
            ia = (int)(row % na), ib = (int)(row / na);
        } else {
            ia = (int)(t % na);
            const long long rest = t / na;
            g6 = (int)(rest % 6), ib = (int)(rest / 6);
        }
        const int side = g6 / 3, gg = g6 % 3 + 1;
        const int mode = side ? mhi : mlo;
        if (mode == BM_HALO) continue;
        const int q = side ? (n - 1 + gg) : -gg;
        bool refl;
        const int src = map_index(q, n, mlo, mhi, refl);
        const long long base = (long long)(ia - ea) * sa + (long long)(ib - eb) * sb;
        const double v = p[base + (long long)src * s];
        p[base + (long long)q * s] = (refl && odd) ? -v : v;
    }
}

// planes [k0, k0 + nk) of a padded field <-> a contiguous (nx, ny, nk) chunk of a host array
__global__ void __launch_bounds__(256) pack_planes_kernel(const Geom g, const double* __restrict__ src,
                                                          double* __restrict__ dst, int k0, int nk,
                                                          int to_padded) {
    const long long rows = (long long)g.ny * nk;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), kk = (int)(row / g.ny);
        const long long mp = (long long)(k0 + kk) * g.sz + (long long)j * g.sy;
        const long long mc = row * g.nx;
        if (to_padded)
            for (int i = threadIdx.x; i < g.nx; i += blockDim.x) dst[mp + i] = src[mc + i];
        else
            for (int i = threadIdx.x; i < g.nx; i += blockDim.x) dst[mc + i] = src[mp + i];
    }
}

__global__ void __launch_bounds__(256) pack_kernel(const Geom g, const double* __restrict__ src,
                                                    double* __restrict__ dst, int to_padded) {
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const long long mp = (long long)k * g.sz + (long long)j * g.sy;
        const long long mc = row * g.nx;
        if (to_padded)
            for (int i = threadIdx.x; i < g.nx; i += blockDim.x) dst[mp + i] = src[mc + i];
        else
            for (int i = threadIdx.x; i < g.nx; i += blockDim.x) dst[mc + i] = src[mp + i];
    }
}

}  // namespace

int launch_fill_ghosts(cudaStream_t st, const Geom& g, const GhostArgs& a) {
    if (a.njobs <= 0) return 0;
    const long long face = (long long)((g.nx > g.ny) ? g.nx : g.ny) * ((g.ny > g.nz) ? g.ny : g.nz) * 6;
    long long b = (face + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    fill_ghosts_kernel<<<dim3((unsigned)b, 1, 3 * a.njobs), 256, 0, st>>>(g, a);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_fill_ghosts_full(cudaStream_t st, const Geom& g, double* p, unsigned par) {
    for (int axis = 0; axis < 3; ++axis) {
        const int ea = (axis == 0) ? 0 : R;               // x ghosts exist once axis 0 is done
        const int eb = (axis == 2) ? R : 0;               // y ghosts exist once axis 1 is done
        const long long na = ((axis == 0) ? g.ny : g.nx) + 2 * ea;
        const long long nb = ((axis == 2) ? g.ny : g.nz) + 2 * eb;
        long long b = (na * nb * 6 + 255) / 256;
        if (b > 148 * 8) b = 148 * 8;
        if (b < 1) b = 1;
        fill_axis_kernel<<<(unsigned)b, 256, 0, st>>>(g, p, axis, (par >> axis) & 1u, ea, eb);
        count_launch();
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_pack(cudaStream_t st, const Geom& g, const double* contiguous, double* padded) {
    long long b = (long long)g.ny * g.nz;
    if (b > 148 * 16) b = 148 * 16;
    pack_kernel<<<(unsigned)b, 256, 0, st>>>(g, contiguous, padded, 1);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_pack_planes(cudaStream_t st, const Geom& g, const double* chunk, double* padded, int k0,
                       int nk) {
    long long b = (long long)g.ny * nk;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) return 0;
    pack_planes_kernel<<<(unsigned)b, 256, 0, st>>>(g, chunk, padded, k0, nk, 1);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_unpack_planes(cudaStream_t st, const Geom& g, const double* padded, double* chunk,
                         int k0, int nk) {
    long long b = (long long)g.ny * nk;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) return 0;
    pack_planes_kernel<<<(unsigned)b, 256, 0, st>>>(g, padded, chunk, k0, nk, 0);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_unpack(cudaStream_t st, const Geom& g, const double* padded, double* contiguous) {
    long long b = (long long)g.ny * g.nz;
    if (b > 148 * 16) b = 148 * 16;
    pack_kernel<<<(unsigned)b, 256, 0, st>>>(g, padded, contiguous, 0);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
