// stencil_tile.cuh -- 2.5-D z-marching tile machinery shared by the fused stencil kernels
// (RHS+predictor, divergence, projection correction, scalar transport, statistics).
//
// A CTA owns a TX x TY column of the (x,y) plane and marches in z over a chunk of planes.
//   * z neighbours of the CTA's own points live in a per-thread register window (7 planes);
//   * the centre plane is staged in shared memory as a cross-shaped tile: TX x TY interior,
//     +-3 columns in x, +-3 rows in y (no corners: no stencil of the reference has mixed
//     derivatives), double buffered so one __syncthreads per plane suffices;
//   * ghost values (periodic wrap / free-slip mirror with parity sign) are resolved when a
//     cell is LOADED, so the compute phase is the reference's interior formula everywhere
//     (src/derivation.f90:43-47, :529-533).  x - (-y) == x + y bitwise, so this reproduces the
//     explicitly written boundary planes (e.g. :137-159) exactly.
#pragma once
#include "o3d_common.cuh"

namespace o3d {

constexpr int TX = 32;
constexpr int TY = 8;
constexpr int NT = TX * TY;        // threads per CTA
constexpr int SW = TX + 2 * R;     // smem tile width
constexpr int SH = TY + 2 * R;     // smem tile height
constexpr int XH = 2 * R * TY;     // x-halo cells per field
constexpr int YH = 2 * R * TX;     // y-halo cells per field

// parity bits of a staged field: bit a set = odd along axis a
//   ux = 0b001, uy = 0b010, uz = 0b100, pp/phi = 0
__device__ __forceinline__ double psign(bool refl, unsigned par, int axis) {
    return (refl && ((par >> axis) & 1u)) ? -1.0 : 1.0;
}

// Addressing of the thread's own (i,j) cell.  Threads just outside a partial last tile act as
// ghost providers for the domain boundary.
struct OwnCell {
    long long off;   // offset inside a plane of the (mapped) cell
    bool in_dom;     // thread computes an output point
    bool loadable;   // thread loads a z-window / centre value
    bool rx, ry;     // own cell is a reflected x- / y-ghost
};

__device__ __forceinline__ OwnCell own_cell(const Dims& g, int i, int j) {
    OwnCell o;
    o.in_dom = (i < g.nx) && (j < g.ny);
    const bool xg = (i >= g.nx) && (i < g.nx + R) && (j < g.ny);
    const bool yg = (j >= g.ny) && (j < g.ny + R) && (i < g.nx);
    o.loadable = o.in_dom || xg || yg;
    bool rx = false, ry = false;
    int gi = 0, gj = 0;
    if (o.loadable) {
        gi = map_index(i, g.nx, g.bx, g.bx, rx);
        gj = map_index(j, g.ny, g.by, g.by, ry);
    }
    o.rx = rx;
    o.ry = ry;
    o.off = (long long)gi + (long long)g.nx * gj;
    return o;
}

// One halo slot: a cell of the cross that is not any thread's own cell.
struct HaloSlot {
    long long off;  // offset inside a plane, or -1 if the cell is not needed
    int sm;         // row * SW + col inside one field's smem tile
    int field;      // staged field index
    bool rx, ry;
};

// Decode halo cell number `idx` of a kernel that stages NF fields; xmask / ymask say which
// fields need x- / y-halos.
template <int NF>
__device__ __forceinline__ HaloSlot halo_slot(const Dims& g, int i0, int j0, int idx,
                                              unsigned xmask, unsigned ymask) {
    HaloSlot h;
    h.off = -1;
    h.sm = 0;
    h.field = 0;
    h.rx = h.ry = false;
    int rem = idx;
#pragma unroll
    for (int c = 0; c < NF; ++c) {
        if ((xmask >> c) & 1u) {
            if (rem >= 0 && rem < XH) {
                const int row = rem / (2 * R), col = rem % (2 * R);
                const int sc = (col < R) ? col : (TX + col);  // smem column
                const int gi = i0 - R + sc, gj = j0 + row;
                h.field = c;
                h.sm = (row + R) * SW + sc;
                if (gj < g.ny && gi < g.nx + R) {
                    bool rx;
                    const int mi = map_index(gi, g.nx, g.bx, g.bx, rx);
                    h.rx = rx;
                    h.off = (long long)mi + (long long)g.nx * gj;
                }
            }
            rem -= XH;
        }
        if ((ymask >> c) & 1u) {
            if (rem >= 0 && rem < YH) {
                const int row = rem / TX, col = rem % TX;
                const int sr = (row < R) ? row : (TY + row);  // smem row
                const int gi = i0 + col, gj = j0 - R + sr;
                h.field = c;
                h.sm = sr * SW + (col + R);
                if (gi < g.nx && gj < g.ny + R) {
                    bool ry;
                    const int mj = map_index(gj, g.ny, g.by, g.by, ry);
                    h.ry = ry;
                    h.off = (long long)g.nx * mj + gi;
                }
            }
            rem -= YH;
        }
    }
    return h;
}

__host__ __device__ constexpr int popc_c(unsigned v) {
    int n = 0;
    while (v) {
        n += v & 1u;
        v >>= 1;
    }
    return n;
}
__host__ __device__ constexpr int halo_cells(unsigned xmask, unsigned ymask) {
    return popc_c(xmask) * XH + popc_c(ymask) * YH;
}
__host__ __device__ constexpr int halo_slots(unsigned xmask, unsigned ymask) {
    return (halo_cells(xmask, ymask) + NT - 1) / NT;
}

// z plane lookup for plane q in [-R, nz+R): stored plane index and reflection flag.
__device__ __forceinline__ int zplane(const Dims& g, int q, bool& refl) {
    return map_index(q, g.nz, g.bz_lo, g.bz_hi, refl);
}

// "even first derivative is literally zero on the wall planes" (src/derivation.f90:87,:105)
__device__ __forceinline__ bool even_wall(int p, int n, int mode_lo, int mode_hi, unsigned par,
                                          int axis) {
    if ((par >> axis) & 1u) return false;
    return (p == 0 && mode_lo == BM_MIRROR) || (p == n - 1 && mode_hi == BM_MIRROR);
}

}  // namespace o3d
