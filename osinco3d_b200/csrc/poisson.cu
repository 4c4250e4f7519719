// poisson.cu -- host drivers of the pressure Poisson solvers.
//   sor_solve : reference poisson_solver_{0000,0011,111111} (src/poisson.f90), red-black fast
//               path or lexicographic-wavefront verification ordering; exit tests and dynamic
//               omega (src/poisson.f90:110-122) evaluated on the device.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <utility>

#include "session.h"

namespace o3d {

SorArgs make_sor_args(o3d_session* s, double* pp, const double* rhs) {
    SorArgs a;
    a.pp = pp;
    a.rhs = rhs;
    // src/poisson.f90:42-51
    const double dx2 = s->cfg.dx * s->cfg.dx, dy2 = s->cfg.dy * s->cfg.dy,
                 dz2 = s->cfg.dz * s->cfg.dz;
    a.oneondx2 = 1.0 / dx2;
    a.oneondy2 = 1.0 / dy2;
    a.oneondz2 = 1.0 / dz2;
    const double twoondx2 = 2.0 * a.oneondx2, twoondy2 = 2.0 * a.oneondy2,
                 twoondz2 = 2.0 * a.oneondz2;
    a.A = -(twoondx2 + twoondy2 + twoondz2);
    a.invA = 1.0 / a.A;
    // neighbour rule per variant: _0000 (periodic x3), _0011 (y mirrored), _111111 (all mirrored)
    const int v = s->sor_variant;
    a.mx = (v == 2) ? BM_MIRROR : BM_WRAP;
    a.my = (v >= 1) ? BM_MIRROR : BM_WRAP;
    const int mz = (v == 2) ? BM_MIRROR : BM_WRAP;
    const int nr = s->cfg.nranks > 1 ? s->cfg.nranks : 1;
    a.mz_lo = (nr > 1 && (s->cfg.rank > 0 || mz == BM_WRAP)) ? BM_HALO : mz;
    a.mz_hi = (nr > 1 && (s->cfg.rank < nr - 1 || mz == BM_WRAP)) ? BM_HALO : mz;
    a.nx = s->g.nx, a.ny = s->g.ny, a.nz = s->g.nz;
    a.sy = s->g.sy, a.sz = s->g.sz;
    a.gz0 = s->z0;
    a.gnz = s->cfg.nz;
    a.seam_x = (a.mx == BM_WRAP) && (a.nx & 1);
    a.seam_y = (a.my == BM_WRAP) && (a.ny & 1);
    a.seam_z = (mz == BM_WRAP) && (a.gnz & 1);
    return a;
}

static double sor_factor(const o3d_session* s) {
    return s->sor_variant == 1 ? 1.01 : 1.05;  // src/poisson.f90:35,:162,:287
}

int sor_solve(o3d_session* s, double* pp, const double* rhs, int* iters, double* dmax) {
    const o3d_config& c = s->cfg;
    SorArgs a = make_sor_args(s, pp, rhs);
    const bool multi = c.nranks > 1;
    if (multi && c.sor_order == O3D_SOR_LEXI_WAVEFRONT) {
        set_error("LEXI_WAVEFRONT ordering is a single-GPU verification mode");
        return O3D_ERR_UNSUPPORTED;
    }
    // reset control block: dmax_old = 1609, src/poisson.f90:52
    SorCtrl* h = s->ctrl_h;
    h->dmax_bits = 0ull;
    h->omega = s->omega;
    h->dmax_old = 1609.0;
    h->dmax_last = 0.0;
    h->iter = 0;
    h->done = (c.kmax < 1) ? 3 : 0;
    O3D_CUDA_CHECK(cudaMemcpyAsync(s->ctrl_d, h, sizeof(SorCtrl), cudaMemcpyHostToDevice, s->st));
    const bool seams = a.seam_x || a.seam_y || a.seam_z;
    const bool wavefront = (c.sor_order == O3D_SOR_LEXI_WAVEFRONT);
    // fast path: fused red+black pass with ping-pong buffers.  On a grid that is not
    // 2-colourable (odd periodic extent) the TMA pass sweeps the two even seam classes and
    // sor_seam_kernel follows with the two thin odd ones on the pass's output buffer -- the same
    // class order, hence the same bits, as the four in-place half-sweeps, which stay selectable
    // with O3D_SOR_SEAM=inplace.  O3D_SOR_FUSED=legacy selects the first-generation fused kernel
    // (index maps, register staging; 2-colourable grids only).
    // (read per solve, so that a test can switch paths inside one process)
    const char* e_fused = getenv("O3D_SOR_FUSED");
    const char* e_seam = getenv("O3D_SOR_SEAM");
    const bool legacy = e_fused && !strcmp(e_fused, "legacy");
    const bool seam_inplace = e_seam && !strcmp(e_seam, "inplace");
    const bool seam_split = e_seam && !strcmp(e_seam, "split");  // seam classes as 2 launches
    const bool fused_off = e_fused && !strcmp(e_fused, "off");  // in-place half-sweeps always
    const bool fused = !wavefront && !fused_off && (!seams || (!legacy && !seam_inplace));
    const bool tma = fused && !legacy;
    double* alt = nullptr;
    if (fused) {
        alt = field(s, O3D_F_PP2);
        if (!alt) return O3D_ERR_CUDA;
    }
    int id_pp = -1, id_rhs = -1;
    bool same_bc = false, fills_pending = false;
    if (tma) {
        // the TMA kernel reads the boundary rule from ghost cells: which session fields are these?
        for (int f = 0; f < O3D_F_COUNT; ++f) {
            if (s->base[f] && s->base[f] + interior_offset(s->g) == pp) id_pp = f;
            if (s->base[f] && s->base[f] + interior_offset(s->g) == rhs) id_rhs = f;
        }
        if (id_pp < 0 || id_rhs < 0) return O3D_ERR_INVALID;
        if (!sor_tmap(s, id_pp) || !sor_tmap(s, O3D_F_PP2) || !sor_tmap(s, id_rhs))
            return O3D_ERR_CUDA;
        // ghost cells = the solver variant's neighbour rule (it may differ from the session's
        // derivative closures when a stateless poisson_solver_xxxx call names another variant)
        same_bc = (a.mx == s->g.bx && a.my == s->g.by && a.mz_lo == s->g.bz_lo &&
                   a.mz_hi == s->g.bz_hi);
        if (same_bc) {
            int rc;
            // (a fill launched here rewrites x / y ghost cells inside the z ghost planes, which a
            // neighbour rank that is ahead may be storing into already: such a solve -- the first
            // after an upload -- gets its initial ghost planes through NCCL, in stream order)
            auto valid = [&](int id, bool edges) {
                const unsigned want = 0x1u | 0x2u | 0x8u | (edges ? 0x10u : 0u);
                return (s->gaxes[id] & want) == want && (s->gpar[id] & 0x7u) == 0u;
            };
            fills_pending = !valid(id_pp, true) || !valid(id_rhs, false);
            if ((rc = ensure_local_ghosts(s, id_pp, 0u, true))) return rc;
            if ((rc = ensure_local_ghosts(s, id_rhs, 0u, false))) return rc;
        } else {
            fills_pending = true;
            Geom gs = s->g;
            gs.bx = a.mx, gs.by = a.my, gs.bz_lo = a.mz_lo, gs.bz_hi = a.mz_hi;
            if (launch_fill_ghosts_full(s->st, gs, pp, 0u)) return O3D_ERR_CUDA;
            if (launch_fill_ghosts_full(s->st, gs, const_cast<double*>(rhs), 0u))
                return O3D_ERR_CUDA;
            touch(s, id_pp), touch(s, id_rhs);
        }
    }
    const double factor = sor_factor(s);
    s->last_sor_path = 0;
    // `same_bc` is a per-rank fact (an interior rank of a z-slab run sees BM_HALO on both sides
    // whatever the rules are), but everything that decides WHICH COLLECTIVE STEPS a rank takes
    // must be the same on every rank: the solver's z rule against the session's z closure,
    // globally.  They differ when nbcz is free-slip under poisson_solver_0000 / _0011 or periodic
    // under _111111 (src/initialization.f90:283-301 looks at the x / y flags only).
    const bool bc_match = a.mx == s->g.bx && a.my == s->g.by &&
                          ((s->sor_variant != 2) == (c.nbcz1 == O3D_PERIODIC));
    const bool speculate_p = s->spec_arm && tma && same_bc && bc_match && id_pp == O3D_F_PP;
    s->spec_state = 0;
    // ---- persistent path: the whole solve in one cooperative launch (sor_persist_kernel.cu) ----
    // O3D_SOR_PERSIST=0 or a forced host poll interval (sor_check_every) keep the launch-per-pass
    // loop below, which stays the bitwise reference of this path.
    {
        const char* e_p = getenv("O3D_SOR_PERSIST");
        bool persist = tma && !(e_p && e_p[0] == '0') && !seam_split && c.sor_check_every <= 0 &&
                       c.kmax >= 1 && id_pp == O3D_F_PP && sor_persist_available();
        PeerSync peer;
        memset(&peer, 0, sizeof(peer));
        if (persist && multi) {
            const char* e_peer = getenv("O3D_SOR_PEER");
            if ((e_peer && e_peer[0] == '0') || comm_peer_setup(s) ||
                comm_peer_args(s, &peer))
                persist = false;  // NCCL halos + all-reduce per sweep, below
        }
        if (persist) {
            if (multi && (id_rhs != O3D_F_RHS || !peer.push_init || fills_pending || !bc_match)) {
                // (a right-hand side that is not the session's O3D_F_RHS has no peer mapping:)
                // ghost planes of the initial iterate (2) and of the right-hand side (1; constant
                // over the solve) through one grouped NCCL exchange, stream-ordered before the
                // kernel.  Otherwise phase 0 of the kernel stores them into the neighbours' memory.
                peer.push_init = 0;
                const long long ioff0 = interior_offset(s->g);
                double* bases[2] = {pp - ioff0, const_cast<double*>(rhs) - ioff0};
                const int widths[2] = {2, 1};
                if (comm_exchange_async(s, bases, widths, 2, s->sor_variant != 2))
                    return O3D_ERR_COMM;
                const int rcw = comm_wait(s);
                if (rcw) return rcw;
            }
            span_begin(s, ST_SOR);
            const int lrc = launch_sor_persist(
                s->st, a, sor_tmap(s, O3D_F_PP), sor_tmap(s, O3D_F_PP2), sor_tmap(s, id_rhs), pp, alt,
                0, a.mx, a.my, a.mz_lo, a.mz_hi, s->ctrl_d, s->persist_sync_d, c.kmax, c.eps, c.kmax,
                c.idyn, factor, 0, multi ? &peer : nullptr);
            if (lrc == 1) return O3D_ERR_CUDA;
            if (lrc == 0) {
                span_end(s, ST_SOR, 0);
                s->last_sor_path = 1 | (multi ? 2 : 0);
                O3D_CUDA_CHECK(cudaMemcpyAsync(h, s->ctrl_d, sizeof(SorCtrl), cudaMemcpyDeviceToHost,
                                               s->st));
                O3D_CUDA_CHECK(cudaEventRecord(s->ev_ctrl, s->st));
                if (speculate_p) {
                    // the projection correction, gated on the solver's outcome, runs while the
                    // host wakes up and queues the next step
                    const int rc = spec_correct_launch(s);
                    if (rc) return rc;
                    s->spec_state = 2;
                }
                O3D_CUDA_CHECK(cudaEventSynchronize(s->ev_ctrl));
                // the guard of the PREVIOUS correction has landed by now (the copy queued just
                // above may or may not have: the flag is sticky, either reading is valid)
                if (s->flag_pending && *s->flag_h) s->diverged = 1;
                if (h->done == 9) {
                    set_error("persistent SOR: a grid barrier or a peer rank timed out");
                    return O3D_ERR_COMM;
                }
                if (multi) {
                    s->peer_iter_base += (unsigned long long)h->iter;
                    if (peer.push_init) s->peer_solves += 1ull;
                }
                goto finished;
            }
            // lrc == 2: not applicable here, fall through
        }
    }
    {
    int launched = 0;
    int batch = s->last_iters > 0 ? s->last_iters : 8;
    if (batch > 64) batch = 64;
    if (c.sor_check_every > 0) batch = c.sor_check_every;
    if (wavefront) batch = 1;
    const long long ioff = interior_offset(s->g);
    const int zwrap = (s->sor_variant != 2);  // _0000 / _0011 wrap in z, _111111 mirrors
    bool rhs_sent = false;
    // speculative projection: see session.h.  Needs the iterate's closure in the ghost cells of
    // whichever ping-pong buffer the last pass wrote (TMA passes do that) and no halo exchange.
    const bool speculate = s->spec_arm && tma && same_bc && !multi && id_pp == O3D_F_PP;
    s->spec_state = 0;
    while (true) {
        if (launched + batch > c.kmax) batch = c.kmax - launched;
        if (batch < 1) batch = 1;
        span_begin(s, ST_SOR);
        for (int b = 0; b < batch; ++b) {
            if (wavefront) {
                const int nh = a.nx + a.ny + a.nz - 2;
                for (int hpl = 0; hpl < nh; ++hpl)
                    if (launch_sor_wavefront(s->st, a, hpl, s->ctrl_d)) return O3D_ERR_CUDA;
            } else if (fused) {
                const int t = launched + b;  // iteration t reads src, writes dst
                double* src = (t & 1) ? alt : pp;
                double* dst = (t & 1) ? pp : alt;
                const int id_src = (t & 1) ? O3D_F_PP2 : id_pp;
                auto pass = [&](cudaStream_t q, int zm, int ze) {
                    if (tma)
                        return launch_sor_tma(q, a, sor_tmap(s, id_src), sor_tmap(s, id_rhs), dst,
                                              a.mx, a.my, a.mz_lo, a.mz_hi, s->ctrl_d, zm, ze);
                    return launch_sor_fused(q, a, src, dst, s->ctrl_d, zm, ze);
                };
                if (multi) {
                    // 2 ghost planes of the previous iterate; the first pass also ships the one
                    // rhs plane per side that the redundant red update of the ghost plane reads
                    // (constant over the solve).  One grouped NCCL call on the comm stream,
                    // overlapped with the interior chunks of this pass.
                    double* bases[2] = {src - ioff, const_cast<double*>(rhs) - ioff};
                    const int widths[2] = {2, 1};
                    if (comm_exchange_async(s, bases, widths, rhs_sent ? 1 : 2, zwrap))
                        return O3D_ERR_COMM;
                    rhs_sent = true;
                    const int rc = launch_overlapped(s, pass);
                    if (rc) return rc;
                } else if (pass(s->st, 0, 0)) {
                    return O3D_ERR_CUDA;
                }
                if (seams) {
                    // odd seam classes (red, then black) in place on the pass's output; they
                    // rewrite the ghost images of the points they touch
                    SorArgs sa = a;
                    sa.pp = dst;
                    if (!multi && !seam_split) {
                        // one cooperative launch: both classes + the end-of-iteration control
                        if (launch_sor_seam_fused(s->st, sa, s->ctrl_d, s->seam_sync_d, c.eps,
                                                  c.kmax, c.idyn, factor))
                            return O3D_ERR_CUDA;
                        continue;
                    }
                    double* dstf[1] = {dst - ioff};
                    for (int colour = 0; colour < 2; ++colour) {
                        if (multi && comm_exchange(s, dstf, 1, 1, zwrap)) return O3D_ERR_COMM;
                        if (launch_sor_rb(s->st, sa, colour, 1, s->ctrl_d, 1)) return O3D_ERR_CUDA;
                    }
                }
            } else {
                double* ppf[1] = {pp - ioff};
                for (int colour = 0; colour < 2; ++colour) {
                    if (multi && comm_exchange(s, ppf, 1, 1, zwrap)) return O3D_ERR_COMM;
                    if (launch_sor_rb(s->st, a, colour, 0, s->ctrl_d)) return O3D_ERR_CUDA;
                }
                for (int colour = 0; colour < 2; ++colour) {
                    if (multi && comm_exchange(s, ppf, 1, 1, zwrap)) return O3D_ERR_COMM;
                    if (launch_sor_rb(s->st, a, colour, 1, s->ctrl_d)) return O3D_ERR_CUDA;
                }
            }
            if (multi && !wavefront &&
                comm_allreduce(s, reinterpret_cast<double*>(&s->ctrl_d->dmax_bits), 1, RED_MAXBITS))
                return O3D_ERR_COMM;
            if (launch_sor_control(s->st, s->ctrl_d, c.eps, c.kmax, c.idyn, factor))
                return O3D_ERR_CUDA;
        }
        span_end(s, ST_SOR, 0);
        if (speculate && launched == 0) {
            const int rc = spec_correct_launch(s);
            if (rc) return rc;
            s->spec_state = 1;
        }
        launched += batch;
        O3D_CUDA_CHECK(
            cudaMemcpyAsync(h, s->ctrl_d, sizeof(SorCtrl), cudaMemcpyDeviceToHost, s->st));
        O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
        poll_flag(s);
        // the gated correction ran iff the solve had finished within the first batch
        if (s->spec_state == 1) s->spec_state = h->done ? 2 : 0;
        if (h->done || launched >= c.kmax) break;
        batch = 4;
        if (c.sor_check_every > 0) batch = c.sor_check_every;
        if (wavefront) batch = 1;
    }
    }
finished:
    if (fused && (h->iter & 1)) {
        // an odd number of ping-pong passes left the iterate in the alternate buffer: swap the
        // two physical fields (O(1), no copy)
        if (id_pp == O3D_F_PP) {
            swap_pp(s);
        } else {
            std::swap(s->base[id_pp], s->base[O3D_F_PP2]);
            std::swap(s->tmap[id_pp], s->tmap[O3D_F_PP2]);
            std::swap(s->tmap_sor[id_pp], s->tmap_sor[O3D_F_PP2]);
            std::swap(s->tmap_st[id_pp], s->tmap_st[O3D_F_PP2]);
            std::swap(s->tmap_sor_ok[id_pp], s->tmap_sor_ok[O3D_F_PP2]);
        }
    }
    touch(s, O3D_F_PP);
    touch(s, O3D_F_PP2);
    if (tma && same_bc && h->iter > 0) {
        // the last pass wrote the even ghost images (faces, edges) of the iterate it stored:
        // the projection correction and the next solve find the closure in place
        // (z slabs, peer-memory solve: the kernel stored 3 planes per side of every iterate into
        // the neighbours' ghost planes, so the rank-boundary halos are in place as well)
        const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO) &&
                           !((s->last_sor_path & 2) && bc_match);
        s->gaxes[O3D_F_PP] = 0x1u | 0x2u | 0x8u | 0x10u | (zhalo ? 0u : 0x4u);
        s->gpar[O3D_F_PP] = 0u;
    }
    s->t_cnt[ST_SOR] += h->iter;
    s->omega = h->omega;  // omega is intent(inout) and persists, src/integration.f90:222,247
    s->last_iters = h->iter;
    // Fortran `iter` after the loop: sweeps done, or kmax+1 when the loop ran out
    if (iters) *iters = (h->done == 3 || h->done == 0) ? c.kmax + 1 : h->iter;
    if (dmax) *dmax = h->dmax_last;
    return O3D_OK;
}

// `sweeps` red-black Gauss-Seidel / SOR iterations with the relaxation factor held in `ctrl`
// (never "done"), no exit tests: the multigrid smoother on level 0 of a single-rank run, through
// the same fused TMA pass (+ split odd seam classes) as sor_solve.  pp must be the session's
// O3D_F_PP field, rhs a session field; on return O3D_F_PP holds the result (the two ping-pong
// buffers are swapped when the number of passes was odd).  Returns O3D_ERR_UNSUPPORTED when the
// fused path does not apply (the caller then uses the in-place half-sweeps: same bits).
int sor_fixed_sweeps(o3d_session* s, const double* rhs, int sweeps, SorCtrl* ctrl) {
    if (sweeps <= 0) return O3D_OK;
    if (s->cfg.nranks > 1) return O3D_ERR_UNSUPPORTED;
    const char* e_fused = getenv("O3D_SOR_FUSED");
    if (e_fused && (!strcmp(e_fused, "off") || !strcmp(e_fused, "legacy"))) return O3D_ERR_UNSUPPORTED;
    double* pp = field(s, O3D_F_PP);
    double* alt = field(s, O3D_F_PP2);
    if (!pp || !alt) return O3D_ERR_CUDA;
    int id_rhs = -1;
    for (int f = 0; f < O3D_F_COUNT; ++f)
        if (s->base[f] && s->base[f] + interior_offset(s->g) == rhs) id_rhs = f;
    if (id_rhs < 0) return O3D_ERR_UNSUPPORTED;
    SorArgs a = make_sor_args(s, pp, rhs);
    if (!sor_tmap(s, O3D_F_PP) || !sor_tmap(s, O3D_F_PP2) || !sor_tmap(s, id_rhs))
        return O3D_ERR_CUDA;
    const bool same_bc = (a.mx == s->g.bx && a.my == s->g.by && a.mz_lo == s->g.bz_lo &&
                          a.mz_hi == s->g.bz_hi);
    int rc;
    if (same_bc) {
        if ((rc = ensure_local_ghosts(s, O3D_F_PP, 0u, true))) return rc;
        if ((rc = ensure_local_ghosts(s, id_rhs, 0u, false))) return rc;
    } else {
        Geom gs = s->g;
        gs.bx = a.mx, gs.by = a.my, gs.bz_lo = a.mz_lo, gs.bz_hi = a.mz_hi;
        if (launch_fill_ghosts_full(s->st, gs, pp, 0u)) return O3D_ERR_CUDA;
        if (launch_fill_ghosts_full(s->st, gs, const_cast<double*>(rhs), 0u)) return O3D_ERR_CUDA;
        touch(s, O3D_F_PP), touch(s, id_rhs);
    }
    const bool seams = a.seam_x || a.seam_y || a.seam_z;
    for (int t = 0; t < sweeps; ++t) {
        double* dst = (t & 1) ? pp : alt;
        const int id_src = (t & 1) ? O3D_F_PP2 : O3D_F_PP;
        if (launch_sor_tma(s->st, a, sor_tmap(s, id_src), sor_tmap(s, id_rhs), dst, a.mx, a.my,
                           a.mz_lo, a.mz_hi, ctrl, 0, 0))
            return O3D_ERR_CUDA;
        if (seams) {
            SorArgs sa = a;
            sa.pp = dst;
            for (int colour = 0; colour < 2; ++colour)
                if (launch_sor_rb(s->st, sa, colour, 1, ctrl, 1)) return O3D_ERR_CUDA;
        }
    }
    if (sweeps & 1) swap_pp(s);
    touch(s, O3D_F_PP);
    touch(s, O3D_F_PP2);
    if (same_bc) {  // the last pass (and the seam sweeps) wrote the ghost images of what they stored
        s->gaxes[O3D_F_PP] = 0x1u | 0x2u | 0x4u | 0x8u | 0x10u;
        s->gpar[O3D_F_PP] = 0u;
    }
    return O3D_OK;
}

}  // namespace o3d
