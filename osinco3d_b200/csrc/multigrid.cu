// multigrid.cu -- host driver of the geometric multigrid Poisson solver behind the reference's
// solve_poisson_multigrid interface (src/poisson_multigrid.f90:10; called from
// src/integration.f90:244 with nlevels = kmax, npre = 5, npost = 4, tol = eps).
//
// The reference routine is undefined behaviour as called (phi declared (0:nx+1,...) but passed
// (nx,ny,nz), one V-cycle against zero-Dirichlet ghosts, tol unused): DESIGN.md section 6.  This
// solver keeps the interface and solves the SAME 7-point operator and neighbour rule as
// poisson_solver (src/poisson.f90:42-51,57-92) with V(npre,npost) cycles until
// max|rhs - L p| / |A| < tol, the quantity SOR's dmax converges to.
//
// Levels are not assumed nested (shipped extents are odd on periodic axes): per axis, 1-D
// linear-interpolation tables are built on the host (oracle/mg_model.py is the NumPy model of
// exactly this construction).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "session.h"

namespace o3d {

struct MgLevel {
    MgGrid g;
    int gn[3];  // global extents (level 0 of a z-slab run holds only nz_local planes)
    double d[3];
    double *p = nullptr, *rhs = nullptr, *res = nullptr;  // level 0: p/rhs alias pp/rhs fields
    MgTables t;                                           // to the next coarser level
    std::vector<void*> owned;                             // device allocations of this level
};

struct MgHierarchy {
    std::vector<MgLevel> lv;
    SorCtrl* ctrl = nullptr;  // smoother control block (omega = 1, never done)
    // cache key
    int n[3], variant, max_levels;
    double d[3];
    double* res0_base = nullptr;  // padded residual buffer of level 0
    // z-slab runs (nranks > 1): level 0 is distributed, every coarser level is REPLICATED on all
    // ranks.  A rank restricts the level-1 planes [ck0, ck0 + nck) whose centre tap it owns --
    // its residual carries 2 exchanged ghost planes per side, which covers every tap -- and the
    // planes are then replicated with grouped broadcasts.  Same taps, same order as on one GPU.
    // The coarse part of a V-cycle (levels >= 1: ~40 small launches per level, launch-latency
    // bound) is a fixed launch sequence: captured once into a CUDA graph, replayed per cycle.
    cudaGraphExec_t coarse_exec = nullptr;
    int graph_npre = -1, graph_npost = -1, graph_launches = 0;
    int ck0 = 0, nck = 0;
    const int* ridx_z_local = nullptr;        // z restriction taps as local plane offsets
    std::vector<long long> gfirst, gcount;    // element ranges of the ranks' level-1 chunks
};

namespace {

const int MG_MIN_N = 5;        // an axis with fewer points is not coarsened further
const int MG_MAX_CYCLES = 100;
const int MG_COARSE_SWEEPS = 40;

int coarse_extent(int n, int mode) {
    int nc;
    if (mode == BM_MIRROR)
        nc = (n % 2) ? (n + 1) / 2 : n / 2 + 1;
    else
        nc = (n % 2 == 0) ? n / 2 : (n + 1) / 2;
    if (n < MG_MIN_N || nc < 3) return 0;
    return nc;
}

struct AxisTab {
    std::vector<int> c0, ridx;
    std::vector<double> w, rw;
    double D;
};

// fine (n points, spacing d) <-> coarse (nc points); nc == 0: identity along this axis
AxisTab axis_tables(int n, double d, int mode, int nc) {
    AxisTab a;
    a.c0.resize(n), a.w.resize(n);
    if (nc == 0) {
        a.D = d;
        a.ridx.assign((size_t)n * 4, 0), a.rw.assign((size_t)n * 4, 0.0);
        for (int i = 0; i < n; ++i) {
            a.c0[i] = i, a.w[i] = 0.0;
            for (int t = 0; t < 4; ++t) a.ridx[4 * i + t] = i;
            a.rw[4 * i] = 1.0;
        }
        return a;
    }
    a.D = (mode == BM_MIRROR) ? (double)(n - 1) * d / (double)(nc - 1) : (double)n * d / (double)nc;
    for (int i = 0; i < n; ++i) {
        // integer arithmetic: the nested cases give weights of exactly 0 and 0.5
        const long long num = (mode == BM_MIRROR) ? (long long)i * (nc - 1) : (long long)i * nc;
        const long long den = (mode == BM_MIRROR) ? (n - 1) : n;
        long long c = num / den;
        double frac = (double)(num - c * den) / (double)den;
        if (mode == BM_MIRROR && c >= nc - 1) c = nc - 2, frac = 1.0;
        a.c0[i] = (int)c, a.w[i] = frac;
    }
    // restriction = transpose on the even / periodic extension, rows normalised to sum 1
    std::vector<std::vector<std::pair<int, double>>> rows(nc);
    auto add = [&](int c, int i, double v) {
        if (v == 0.0) return;
        for (auto& e : rows[c])
            if (e.first == i) {
                e.second += v;
                return;
            }
        rows[c].push_back({i, v});
    };
    for (int i = 0; i < n; ++i) {
        const int c = a.c0[i];
        int c1 = c + 1;
        if (mode == BM_WRAP) c1 %= nc;
        add(c, i, 1.0 - a.w[i]);
        add(c1, i, a.w[i]);
    }
    if (mode == BM_MIRROR) {
        // wall coarse nodes also collect the mirror images of the off-wall fine nodes
        for (auto& e : rows[0])
            if (e.first != 0) e.second *= 2.0;
        for (auto& e : rows[nc - 1])
            if (e.first != n - 1) e.second *= 2.0;
    }
    a.ridx.assign((size_t)nc * 4, 0), a.rw.assign((size_t)nc * 4, 0.0);
    for (int c = 0; c < nc; ++c) {
        auto& r = rows[c];
        // ascending fine index: fixed summation order
        for (size_t x = 0; x < r.size(); ++x)
            for (size_t y = x + 1; y < r.size(); ++y)
                if (r[y].first < r[x].first) std::swap(r[x], r[y]);
        double s = 0.0;
        for (auto& e : r) s += e.second;
        const size_t nt = r.size() < 4 ? r.size() : 4;
        for (size_t t = 0; t < 4; ++t) {
            a.ridx[4 * c + t] = r.empty() ? 0 : r[t < nt ? t : 0].first;
            a.rw[4 * c + t] = (t < nt) ? r[t].second / s : 0.0;
        }
    }
    return a;
}

template <class T>
int to_device(MgLevel& L, const std::vector<T>& h, const T** out) {
    T* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return 1;
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
        return 1;
    L.owned.push_back(d);
    *out = d;
    return 0;
}

void set_operator(MgGrid& g, const double* d) {
    // src/poisson.f90:42-51 with the level's spacings
    g.ox = 1.0 / (d[0] * d[0]);
    g.oy = 1.0 / (d[1] * d[1]);
    g.oz = 1.0 / (d[2] * d[2]);
    g.A = -(2.0 * g.ox + 2.0 * g.oy + 2.0 * g.oz);
    g.invA = 1.0 / g.A;
}

SorArgs smoother_args(const MgLevel& L) {
    SorArgs a;
    a.pp = L.p, a.rhs = L.rhs;
    a.oneondx2 = L.g.ox, a.oneondy2 = L.g.oy, a.oneondz2 = L.g.oz;
    a.A = L.g.A, a.invA = L.g.invA;
    a.mx = L.g.mx, a.my = L.g.my, a.mz_lo = L.g.mz_lo, a.mz_hi = L.g.mz_hi;
    a.nx = L.g.nx, a.ny = L.g.ny, a.nz = L.g.nz;
    a.sy = L.g.sy, a.sz = L.g.sz;
    a.gz0 = 0, a.gnz = L.g.nz;
    a.seam_x = (a.mx == BM_WRAP) && (a.nx & 1);
    a.seam_y = (a.my == BM_WRAP) && (a.ny & 1);
    a.seam_z = (L.g.mz_lo == BM_WRAP) && (a.nz & 1);
    return a;
}

// red-black Gauss-Seidel sweeps (the SOR half-sweep kernels with omega = 1).  dist: level 0 of
// a z-slab run -- global colouring / seam plane, one ghost plane of p exchanged before every
// class sweep (exactly the multi-rank in-place SOR of sor_solve)
int smooth(o3d_session* s, MgLevel& L, int sweeps, bool dist) {
    if (!dist && L.p == field(s, O3D_F_PP) && !(getenv("O3D_MG_SMOOTHER") &&
                                                 !strcmp(getenv("O3D_MG_SMOOTHER"), "inplace"))) {
        // level 0 of a single-rank run: the fused TMA pass of the SOR solver with omega = 1
        // (bitwise the same iterates as the in-place half-sweeps below, at 24 instead of 32 B/pt
        // and 1-3 instead of 2-4 launches per sweep).  The iterate may end up in the ping-pong
        // partner: O3D_F_PP is re-pointed, so refresh the level's pointer.
        const int rc = sor_fixed_sweeps(s, L.rhs, sweeps, s->mg->ctrl);
        if (rc == O3D_OK) {
            L.p = field(s, O3D_F_PP);
            return 0;
        }
        if (rc != O3D_ERR_UNSUPPORTED) return 1;
    }
    SorArgs a = smoother_args(L);
    double* pf[1] = {nullptr};
    int zwrap = 0;
    if (dist) {
        const SorArgs g = make_sor_args(s, L.p, L.rhs);
        a.mz_lo = g.mz_lo, a.mz_hi = g.mz_hi;
        a.gz0 = g.gz0, a.gnz = g.gnz, a.seam_z = g.seam_z;
        pf[0] = L.p - interior_offset(s->g);
        zwrap = (s->sor_variant != 2);
    }
    const bool seams = a.seam_x || a.seam_y || a.seam_z;
    for (int q = 0; q < sweeps; ++q) {
        for (int colour = 0; colour < 2; ++colour) {
            if (dist && comm_exchange(s, pf, 1, 1, zwrap)) return 1;
            if (launch_sor_rb(s->st, a, colour, 0, s->mg->ctrl)) return 1;
        }
        if (seams)
            for (int colour = 0; colour < 2; ++colour) {
                if (dist && comm_exchange(s, pf, 1, 1, zwrap)) return 1;
                if (launch_sor_rb(s->st, a, colour, 1, s->mg->ctrl)) return 1;
            }
    }
    return 0;
}

int build(o3d_session* s, int max_levels) {
    MgHierarchy* H = new MgHierarchy();
    s->mg = H;
    const int v = s->sor_variant;
    const int modes[3] = {(v == 2) ? BM_MIRROR : BM_WRAP, (v >= 1) ? BM_MIRROR : BM_WRAP,
                          (v == 2) ? BM_MIRROR : BM_WRAP};
    const bool multi = s->cfg.nranks > 1;
    H->n[0] = s->g.nx, H->n[1] = s->g.ny, H->n[2] = s->g.nz;
    H->d[0] = s->cfg.dx, H->d[1] = s->cfg.dy, H->d[2] = s->cfg.dz;
    H->variant = v, H->max_levels = max_levels;
    if (cudaMalloc(&H->ctrl, sizeof(SorCtrl)) != cudaSuccess) return 1;
    SorCtrl c0;
    c0.dmax_bits = 0ull, c0.omega = 1.0, c0.dmax_old = 0.0, c0.dmax_last = 0.0;
    c0.iter = 0, c0.done = 0, c0.pad0 = c0.pad1 = 0;
    if (cudaMemcpy(H->ctrl, &c0, sizeof(c0), cudaMemcpyHostToDevice) != cudaSuccess) return 1;

    MgLevel L0;
    L0.g.nx = s->g.nx, L0.g.ny = s->g.ny, L0.g.nz = s->g.nz;
    L0.g.sy = s->g.sy, L0.g.sz = s->g.sz;
    L0.g.mx = modes[0], L0.g.my = modes[1];
    {   // z rule of the Poisson VARIANT (src/initialization.f90:283-301: _0000 / _0011 wrap z,
        // _111111 mirrors it), not of the session's nbcz closures -- the two differ when nbcz is
        // free-slip under variant 0/1 or periodic under variant 2 -- with BM_HALO on the sides
        // that face another rank: exactly what sor_solve and the level-0 smoother use, and what
        // comm_exchange(zwrap = variant != 2) fills
        const SorArgs sa = make_sor_args(s, nullptr, nullptr);
        L0.g.mz_lo = sa.mz_lo;
        L0.g.mz_hi = sa.mz_hi;
    }
    L0.gn[0] = s->g.nx, L0.gn[1] = s->g.ny, L0.gn[2] = s->cfg.nz;
    for (int a = 0; a < 3; ++a) L0.d[a] = H->d[a];
    set_operator(L0.g, L0.d);
    if (cudaMalloc(&H->res0_base, (size_t)s->felems * sizeof(double)) != cudaSuccess) return 1;
    if (cudaMemset(H->res0_base, 0, (size_t)s->felems * sizeof(double)) != cudaSuccess) return 1;
    L0.res = H->res0_base + interior_offset(s->g);
    H->lv.push_back(L0);

    while ((int)H->lv.size() < max_levels) {
        MgLevel& F = H->lv.back();
        const int fn[3] = {F.gn[0], F.gn[1], F.gn[2]};
        const bool from_dist = multi && H->lv.size() == 1;
        int nc[3];
        bool any = false;
        for (int a = 0; a < 3; ++a) {
            nc[a] = coarse_extent(fn[a], modes[a]);
            any = any || nc[a];
        }
        if (!any) break;
        MgLevel C;
        int cn[3];
        for (int a = 0; a < 3; ++a) {
            const AxisTab t = axis_tables(fn[a], F.d[a], modes[a], nc[a]);
            cn[a] = nc[a] ? nc[a] : fn[a];
            C.d[a] = t.D;
            if (to_device(F, t.c0, &F.t.c0[a]) || to_device(F, t.w, &F.t.w[a]) ||
                to_device(F, t.ridx, &F.t.ridx[a]) || to_device(F, t.rw, &F.t.rw[a]))
                return 1;
            if (a == 2 && from_dist) {
                // owner of a level-1 plane = the rank that owns its heaviest (centre) fine tap;
                // every other tap must lie within the 2 exchanged ghost planes of that rank
                const int P = s->cfg.nranks, gnz = s->cfg.nz, ncz = cn[2];
                std::vector<int> owner(ncz), local(t.ridx.size(), 0);
                for (int ck = 0; ck < ncz; ++ck) {
                    int best = 0;
                    for (int q = 1; q < 4; ++q)
                        if (t.rw[4 * ck + q] > t.rw[4 * ck + best]) best = q;
                    const int centre = t.ridx[4 * ck + best];
                    int z0r = 0, nzr = 0, r = 0;
                    for (r = 0; r < P; ++r) {
                        o3d_slab_partition(gnz, P, r, &z0r, &nzr);
                        if (centre >= z0r && centre < z0r + nzr) break;
                    }
                    owner[ck] = r;
                    if (ck > 0 && owner[ck] < owner[ck - 1]) {
                        set_error("multigrid: level-1 plane ownership is not monotone in z");
                        return 2;
                    }
                    if (r != s->cfg.rank) continue;
                    for (int q = 0; q < 4; ++q) {
                        if (t.rw[4 * ck + q] == 0.0) continue;
                        int off = t.ridx[4 * ck + q] - s->z0;
                        if (modes[2] == BM_WRAP) {  // periodic image inside [-2, nz_local + 2)
                            off = ((off % gnz) + gnz) % gnz;
                            if (off >= s->nzl + 2) off -= gnz;
                        }
                        if (off < -2 || off >= s->nzl + 2) {
                            set_error("multigrid: restriction tap %d planes outside the slab", off);
                            return 2;
                        }
                        local[4 * ck + q] = off;
                    }
                }
                H->gfirst.assign(P, 0), H->gcount.assign(P, 0);
                const long long plane = (long long)cn[0] * cn[1];
                for (int r = 0; r < P; ++r) {
                    int a0 = -1, cnt = 0;
                    for (int ck = 0; ck < ncz; ++ck)
                        if (owner[ck] == r) {
                            if (a0 < 0) a0 = ck;
                            ++cnt;
                        }
                    H->gfirst[r] = (a0 < 0 ? 0 : a0) * plane, H->gcount[r] = cnt * plane;
                    if (r == s->cfg.rank) H->ck0 = a0 < 0 ? 0 : a0, H->nck = cnt;
                }
                if (to_device(F, local, &H->ridx_z_local)) return 1;
            }
        }
        C.g.nx = cn[0], C.g.ny = cn[1], C.g.nz = cn[2];
        C.gn[0] = cn[0], C.gn[1] = cn[1], C.gn[2] = cn[2];
        C.g.sy = cn[0], C.g.sz = (long long)cn[0] * cn[1];
        C.g.mx = modes[0], C.g.my = modes[1], C.g.mz_lo = C.g.mz_hi = modes[2];
        set_operator(C.g, C.d);
        const size_t n = (size_t)cn[0] * cn[1] * cn[2];
        double* buf = nullptr;
        if (cudaMalloc(&buf, 3 * n * sizeof(double)) != cudaSuccess) return 1;
        if (cudaMemset(buf, 0, 3 * n * sizeof(double)) != cudaSuccess) return 1;
        C.owned.push_back(buf);
        C.p = buf, C.rhs = buf + n, C.res = buf + 2 * n;
        H->lv.push_back(C);
    }
    return 0;
}

}  // namespace

void mg_destroy(o3d_session* s) {
    MgHierarchy* H = s->mg;
    if (!H) return;
    for (auto& L : H->lv)
        for (void* p : L.owned) cudaFree(p);
    if (H->coarse_exec) cudaGraphExecDestroy(H->coarse_exec);
    if (H->ctrl) cudaFree(H->ctrl);
    if (H->res0_base) cudaFree(H->res0_base);
    delete H;
    s->mg = nullptr;
}

int mg_solve(o3d_session* s, double* pp, const double* rhs, int nlevels, int npre, int npost,
             double tol, int* cycles, double* dmax) {
    if (s->sor_variant < 0) return O3D_ERR_BC;
    const bool multi = s->cfg.nranks > 1;
    if (npre < 0 || npost < 0 || npre + npost < 1) {
        set_error("multigrid needs npre + npost >= 1");
        return O3D_ERR_INVALID;
    }
    // the reference passes nlevels = kmax (src/integration.f90:244): treat it as a depth cap
    int max_levels = nlevels < 1 ? 1 : (nlevels > 32 ? 32 : nlevels);
    MgHierarchy* H = s->mg;
    if (H && (H->n[0] != s->g.nx || H->n[1] != s->g.ny || H->n[2] != s->g.nz ||
              H->d[0] != s->cfg.dx || H->d[1] != s->cfg.dy || H->d[2] != s->cfg.dz ||
              H->variant != s->sor_variant || H->max_levels != max_levels)) {
        mg_destroy(s);
        H = nullptr;
    }
    if (!H) {
        const int brc = build(s, max_levels);
        if (brc) {
            if (brc == 1)
                set_error("multigrid hierarchy allocation failed: %s",
                          cudaGetErrorString(cudaGetLastError()));
            mg_destroy(s);
            return brc == 1 ? O3D_ERR_CUDA : O3D_ERR_UNSUPPORTED;
        }
        H = s->mg;
    }
    H->lv[0].p = pp;
    H->lv[0].rhs = const_cast<double*>(rhs);
    const int nl = (int)H->lv.size();
    unsigned long long* maxbits = &H->ctrl->dmax_bits;
    const double absA = fabs(H->lv[0].g.A);
    double last = 0.0, prev = 1e300;
    int cyc = 0;
    double* ppf[1] = {pp - interior_offset(s->g)};
    double* resf[1] = {H->res0_base};
    const int zwrap = (s->sor_variant != 2);
    span_begin(s, ST_SOR);
    for (;; ++cyc) {
        // stopping test on the true residual, same measure as SOR's dmax (src/poisson.f90:100)
        O3D_CUDA_CHECK(cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long), s->st));
        if (multi && comm_exchange(s, ppf, 1, 1, zwrap)) return O3D_ERR_COMM;
        // (level 0's iterate may have moved to the ping-pong partner: always through lv[0].p)
        if (launch_mg_residual(s->st, H->lv[0].g, H->lv[0].p, rhs, nullptr, maxbits))
            return O3D_ERR_CUDA;
        if (multi && comm_allreduce(s, reinterpret_cast<double*>(maxbits), 1, RED_MAXBITS))
            return O3D_ERR_COMM;
        O3D_CUDA_CHECK(cudaMemcpyAsync(s->scal_h, maxbits, sizeof(double), cudaMemcpyDeviceToHost,
                                       s->st));
        O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
        poll_flag(s);
        last = s->scal_h[0] / absA;
        if (last < tol || cyc >= MG_MAX_CYCLES) break;
        if (cyc >= 2 && last > 0.9 * prev) break;  // stalled at the round-off / compatibility floor
        prev = last;
        // one V-cycle.  Level 0 (session fields; distributed over the z slabs when multi) ...
        if (nl == 1) {
            if (smooth(s, H->lv[0], npre + npost, multi)) return O3D_ERR_CUDA;
            continue;
        }
        {
            MgLevel& F = H->lv[0];
            MgLevel& C = H->lv[1];
            if (smooth(s, F, npre, multi)) return O3D_ERR_CUDA;
            if (multi && comm_exchange(s, ppf, 1, 1, zwrap)) return O3D_ERR_COMM;
            if (launch_mg_residual(s->st, F.g, F.p, F.rhs, F.res, nullptr)) return O3D_ERR_CUDA;
            if (multi) {
                // 2 ghost planes of the residual, restrict the owned level-1 planes, replicate
                if (comm_exchange(s, resf, 1, 2, zwrap)) return O3D_ERR_COMM;
                MgTables tl = F.t;
                tl.ridx[2] = H->ridx_z_local;
                const size_t nc = (size_t)C.g.nx * C.g.ny * C.g.nz;
                O3D_CUDA_CHECK(cudaMemsetAsync(C.p, 0, nc * sizeof(double), s->st));
                if (launch_mg_restrict(s->st, F.g, C.g, tl, F.res, C.rhs, C.p, H->ck0, H->nck))
                    return O3D_ERR_CUDA;
                if (comm_allgather_chunks(s, C.rhs, H->gfirst.data(), H->gcount.data()))
                    return O3D_ERR_COMM;
            } else if (launch_mg_restrict(s->st, F.g, C.g, F.t, F.res, C.rhs, C.p, 0, C.g.nz)) {
                return O3D_ERR_CUDA;
            }
        }
        // ... levels >= 1 (local / replicated buffers, no communication): fixed launch sequence
        auto coarse_part = [&]() -> int {
            for (int l = 1; l < nl - 1; ++l) {
                MgLevel& F = H->lv[l];
                MgLevel& C = H->lv[l + 1];
                if (smooth(s, F, npre, false)) return 1;
                if (launch_mg_residual(s->st, F.g, F.p, F.rhs, F.res, nullptr)) return 1;
                if (launch_mg_restrict(s->st, F.g, C.g, F.t, F.res, C.rhs, C.p, 0, C.g.nz)) return 1;
            }
            MgLevel& B = H->lv[nl - 1];
            if (launch_mg_coarse(s->st, B.g, B.p, B.rhs, MG_COARSE_SWEEPS)) return 1;
            for (int l = nl - 2; l >= 1; --l) {
                MgLevel& F = H->lv[l];
                MgLevel& C = H->lv[l + 1];
                if (launch_mg_prolong(s->st, F.g, C.g, F.t, C.p, F.p, 0)) return 1;
                if (smooth(s, F, npost, false)) return 1;
            }
            return 0;
        };
        const bool use_graph = !(getenv("O3D_MG_GRAPH") && atoi(getenv("O3D_MG_GRAPH")) == 0);
        if (use_graph && nl > 2) {
            if (H->coarse_exec && (H->graph_npre != npre || H->graph_npost != npost)) {
                cudaGraphExecDestroy(H->coarse_exec);
                H->coarse_exec = nullptr;
            }
            if (!H->coarse_exec) {
                const long long before = o3d_kernel_launches();
                cudaGraph_t graph = nullptr;
                O3D_CUDA_CHECK(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeRelaxed));
                const int crc = coarse_part();
                const cudaError_t ce = cudaStreamEndCapture(s->st, &graph);
                if (crc || ce != cudaSuccess || !graph) {
                    if (graph) cudaGraphDestroy(graph);
                    set_error("multigrid: capture of the coarse V-cycle failed: %s",
                              cudaGetErrorString(ce));
                    return O3D_ERR_CUDA;
                }
                H->graph_launches = (int)(o3d_kernel_launches() - before);
                count_launch(-H->graph_launches);  // captured, not run
                const cudaError_t ie = cudaGraphInstantiate(&H->coarse_exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) {
                    H->coarse_exec = nullptr;
                    set_error("multigrid: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
                    return O3D_ERR_CUDA;
                }
                H->graph_npre = npre, H->graph_npost = npost;
            }
            O3D_CUDA_CHECK(cudaGraphLaunch(H->coarse_exec, s->st));
            count_launch(H->graph_launches);
        } else if (coarse_part()) {
            return O3D_ERR_CUDA;
        }
        {   // ... and back up to level 0
            MgLevel& F = H->lv[0];
            MgLevel& C = H->lv[1];
            if (launch_mg_prolong(s->st, F.g, C.g, F.t, C.p, F.p, multi ? s->z0 : 0))
                return O3D_ERR_CUDA;
            touch(s, O3D_F_PP);  // interior changed: ghost images are stale
            if (smooth(s, F, npost, multi)) return O3D_ERR_CUDA;
        }
    }
    span_end(s, ST_SOR, cyc);
    touch(s, O3D_F_PP);
    if (cycles) *cycles = cyc;
    if (dmax) *dmax = last;
    return O3D_OK;
}

}  // namespace o3d
