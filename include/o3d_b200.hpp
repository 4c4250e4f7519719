// o3d_b200.hpp -- C++ host-side mirror of the reference's Fortran module interfaces over the C ABI
// of libo3d_b200.so (include/o3d_b200.h).  Header only.
//
// The reference (jojoledemago/osinco3d) is compiled Fortran and the build image has no Fortran
// compiler, so next to the ISO_C_BINDING shims in fortran/ (which could not be compiled here) this
// is the compiled-language host side: one namespace per reference module, one function per module
// procedure, SAME names, argument order and meaning -- arrays are the reference's contiguous
// real(8) (nx,ny,nz) allocatables (i fastest) passed as pointers, the extents travel in a Shape
// where the Fortran routine reads them from size(f,.).  Where the reference prints and stops
// (src/integration.f90:103-104,310-325; src/initialization.f90:238-242) an o3d::Error carrying the
// C status code is thrown.  Nothing is computed on the host; without a CUDA device every call
// throws O3D_ERR_NO_DEVICE.
//
//   module derivation      src/derivation.f90            o3d::derivation::derx_00 ... derzz_2dsim
//   schemes() + pointers   src/initialization.f90:226    o3d::initialization::schemes, derxp ... derzzi
//   module diffoper        src/differential_operators.f90  o3d::diffoper::divergence, rotational, ...
//   module les_turbulence  src/les_turbulence.f90:10     o3d::les_turbulence::calculate_nu_t
//   module poisson         src/poisson.f90:6,132,257     o3d::poisson::poisson_solver_0000 ...
//   module poisson_multigrid  src/poisson_multigrid.f90:10  o3d::poisson_multigrid::solve_poisson_multigrid
//   module integration     src/integration.f90           o3d::integration::predict_velocity ...
//   main loop state        src/osinco3d_main.f90:97-128  o3d::Session (device-resident)
#pragma once
#include <stdexcept>
#include <string>

#include "o3d_b200.h"

namespace o3d {

struct Shape {
    int nx, ny, nz;
};

class Error : public std::runtime_error {
public:
    Error(int code, const std::string& where)
        : std::runtime_error(where + ": " + o3d_last_error()), code_(code) {}
    int code() const { return code_; }

private:
    int code_;
};

// status -> exception; `allow` passes one non-zero status through to the caller
inline int check(int rc, const char* where, int allow = O3D_OK) {
    if (rc != O3D_OK && rc != allow) throw Error(rc, where);
    return rc;
}

namespace initialization {
// schemes(), src/initialization.f90:226-304
inline void schemes(int nbcx1, int nbcxn, int nbcy1, int nbcyn, int nbcz1, int nbczn, int sim2d = 0) {
    check(o3d_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d), "schemes");
}
// the 12 procedure pointers schemes() binds, src/initialization.f90:104-109
#define O3D_HPP_DER(name)                                                     \
    inline void name(double* df, const double* f, double d, Shape s) {        \
        check(o3d_##name(df, f, d, s.nx, s.ny, s.nz), #name);                 \
    }
O3D_HPP_DER(derxp) O3D_HPP_DER(derxxp) O3D_HPP_DER(derxi) O3D_HPP_DER(derxxi)
O3D_HPP_DER(deryp) O3D_HPP_DER(deryyp) O3D_HPP_DER(deryi) O3D_HPP_DER(deryyi)
O3D_HPP_DER(derzp) O3D_HPP_DER(derzzp) O3D_HPP_DER(derzi) O3D_HPP_DER(derzzi)
// poi_type pointer, src/initialization.f90:110,283-301; returns the Fortran loop variable at exit
inline int poisson_solver(double* pp, const double* rhs, double dx, double dy, double dz, Shape s,
                          double& omega, double eps, int kmax, int idyn, double* dmax = nullptr) {
    int iters = 0;
    check(o3d_poisson_solver(pp, rhs, dx, dy, dz, s.nx, s.ny, s.nz, &omega, eps, kmax, idyn, &iters,
                             dmax),
          "poisson_solver");
    return iters;
}
}  // namespace initialization

namespace derivation {
// der_type: subroutine(df, f, d), src/initialization.f90:86-91; the 18 + 2 routines of
// src/derivation.f90:6-948
O3D_HPP_DER(derx_00) O3D_HPP_DER(derxp_11) O3D_HPP_DER(derxi_11)
O3D_HPP_DER(dery_00) O3D_HPP_DER(deryp_11) O3D_HPP_DER(deryi_11)
O3D_HPP_DER(derz_00) O3D_HPP_DER(derzp_11) O3D_HPP_DER(derzi_11)
O3D_HPP_DER(derxx_00) O3D_HPP_DER(derxxp_11) O3D_HPP_DER(derxxi_11)
O3D_HPP_DER(deryy_00) O3D_HPP_DER(deryyp_11) O3D_HPP_DER(deryyi_11)
O3D_HPP_DER(derzz_00) O3D_HPP_DER(derzzp_11) O3D_HPP_DER(derzzi_11)
O3D_HPP_DER(derz_2dsim) O3D_HPP_DER(derzz_2dsim)
}  // namespace derivation
#undef O3D_HPP_DER

namespace diffoper {
// src/differential_operators.f90:7, :40, :79
inline void divergence(double* divf, const double* fx, const double* fy, const double* fz, double dx,
                       double dy, double dz, Shape s, int odd) {
    check(o3d_divergence(divf, fx, fy, fz, dx, dy, dz, s.nx, s.ny, s.nz, odd), "divergence");
}
inline void rotational(double* rotx, double* roty, double* rotz, const double* ux, const double* uy,
                       const double* uz, double dx, double dy, double dz, Shape s) {
    check(o3d_rotational(rotx, roty, rotz, ux, uy, uz, dx, dy, dz, s.nx, s.ny, s.nz), "rotational");
}
inline void calculate_Q_criterion(double* Q, const double* ux, const double* uy, const double* uz,
                                  double dx, double dy, double dz, Shape s) {
    check(o3d_calculate_q_criterion(Q, ux, uy, uz, dx, dy, dz, s.nx, s.ny, s.nz),
          "calculate_Q_criterion");
}
}  // namespace diffoper

namespace les_turbulence {
// src/les_turbulence.f90:10; stats6 (optional) = function_stats(nu_t) the reference prints at :89-90
inline void calculate_nu_t(double* nu_t, const double* ux, const double* uy, const double* uz,
                           double dx, double dy, double dz, double cs, double delta, Shape s,
                           double* stats6 = nullptr) {
    check(o3d_calculate_nu_t(nu_t, ux, uy, uz, dx, dy, dz, cs, delta, s.nx, s.ny, s.nz, stats6),
          "calculate_nu_t");
}
}  // namespace les_turbulence

namespace poisson {
// src/poisson.f90:6, :132, :257 (poi_type, src/initialization.f90:93-102); pp and omega are inout;
// the return value is the Fortran loop variable `iter` after the loop
#define O3D_HPP_POI(name)                                                                       \
    inline int name(double* pp, const double* rhs, double dx, double dy, double dz, Shape s,    \
                    double& omega, double eps, int kmax, int idyn, double* dmax = nullptr) {    \
        int iters = 0;                                                                          \
        check(o3d_##name(pp, rhs, dx, dy, dz, s.nx, s.ny, s.nz, &omega, eps, kmax, idyn, &iters, \
                         dmax),                                                                 \
              #name);                                                                           \
        return iters;                                                                           \
    }
O3D_HPP_POI(poisson_solver_0000) O3D_HPP_POI(poisson_solver_0011) O3D_HPP_POI(poisson_solver_111111)
#undef O3D_HPP_POI
}  // namespace poisson

namespace poisson_multigrid {
// src/poisson_multigrid.f90:10 (same operator and boundary rule as poisson_solver);
// returns the number of V-cycles
inline int solve_poisson_multigrid(double* phi, const double* rhs, double dx, double dy, double dz,
                                   Shape s, int nlevels, int npre, int npost, double tol,
                                   double* dmax = nullptr) {
    int cycles = 0;
    check(o3d_solve_poisson_multigrid(phi, rhs, dx, dy, dz, s.nx, s.ny, s.nz, nlevels, npre, npost,
                                      tol, &cycles, dmax),
          "solve_poisson_multigrid");
    return cycles;
}
}  // namespace poisson_multigrid

namespace integration {
// src/integration.f90:14-16; fux / fuy / fuz are (nx,ny,nz,3) inout, adt / bdt / cdt the 3-vectors
// of src/initialization.f90:194-202
inline void predict_velocity(double* ux_pred, double* uy_pred, double* uz_pred, const double* ux,
                             const double* uy, const double* uz, double* fux, double* fuy,
                             double* fuz, double re, const double* adt, const double* bdt,
                             const double* cdt, int itime, int itscheme, double dx, double dy,
                             double dz, Shape s, int iles, double cs, double delta, double* nu_t) {
    check(o3d_predict_velocity(ux_pred, uy_pred, uz_pred, ux, uy, uz, fux, fuy, fuz, re, adt, bdt, cdt,
                               itime, itscheme, dx, dy, dz, s.nx, s.ny, s.nz, iles, cs, delta, nu_t),
          "predict_velocity");
}
// src/integration.f90:199-200; returns the solver's iteration (or V-cycle) count
inline int correct_pression(double* pp, const double* ux_pred, const double* uy_pred,
                            const double* uz_pred, double dx, double dy, double dz, Shape s,
                            double dt, double& omega, double eps, int kmax, int idyn, int multigrid,
                            double* dmax = nullptr) {
    int iters = 0;
    check(o3d_correct_pression(pp, ux_pred, uy_pred, uz_pred, dx, dy, dz, s.nx, s.ny, s.nz, dt, &omega,
                               eps, kmax, idyn, multigrid, &iters, dmax),
          "correct_pression");
    return iters;
}
// src/integration.f90:257-258; throws Error(O3D_ERR_DIVERGED) where the reference stops on a NaN
// or a value above 1000 (:309-325) -- the outputs are complete when it does
inline void correct_velocity(double* ux, double* uy, double* uz, const double* ux_pred,
                             const double* uy_pred, const double* uz_pred, const double* pp, double dt,
                             double dx, double dy, double dz, Shape s) {
    check(o3d_correct_velocity(ux, uy, uz, ux_pred, uy_pred, uz_pred, pp, dt, dx, dy, dz, s.nx, s.ny,
                               s.nz),
          "correct_velocity");
}
// src/integration.f90:332-333; src may be null (the reference never assigns it)
inline void transeq(double* phi, const double* ux, const double* uy, const double* uz,
                    const double* src, double* fphi, double re, double sc, const double* adt,
                    const double* bdt, const double* cdt, int itime, int itscheme, double dx, double dy,
                    double dz, Shape s, int iles, const double* nu_t) {
    check(o3d_transeq(phi, ux, uy, uz, src, fphi, re, sc, adt, bdt, cdt, itime, itscheme, dx, dy, dz,
                      s.nx, s.ny, s.nz, iles, nu_t),
          "transeq");
}
}  // namespace integration

// Device-resident state of the main loop (section B of o3d_b200.h): RAII over o3d_session.
class Session {
public:
    explicit Session(const o3d_config& cfg) : s_(nullptr) {
        check(o3d_session_create(&cfg, &s_), "o3d_session_create");
    }
    ~Session() {
        if (s_) o3d_session_destroy(s_);
    }
    Session(const Session&) = delete;
    Session& operator=(const Session&) = delete;
    o3d_session* get() const { return s_; }
    void upload(int field, const double* host) { check(o3d_upload(s_, field, host), "o3d_upload"); }
    void download(int field, double* host) { check(o3d_download(s_, field, host), "o3d_download"); }
    // one time step, src/osinco3d_main.f90:105-115; returns the Poisson iteration count
    int step(int itime, double* dmax = nullptr) {
        int iters = 0;
        check(o3d_step(s_, itime, &iters, dmax), "o3d_step");
        return iters;
    }
    void sync() { check(o3d_sync(s_), "o3d_sync"); }
    // ---- what the driver does around the step, src/osinco3d_main.f90:104,116-188 ----
    // everything :116-128 prints per step (include/o3d_b200.h: o3d_s_step_diagnostics)
    void step_diagnostics(double out23[23]) {
        check(o3d_s_step_diagnostics(s_, out23), "o3d_s_step_diagnostics");
    }
    void old_values() { check(o3d_s_old_values(s_), "o3d_s_old_values"); }               // :104
    void calculate_residuals(double dt, double t_ref, double u_ref, double out15[15]) {  // :167
        check(o3d_s_calculate_residuals(s_, dt, t_ref, u_ref, out15), "o3d_s_calculate_residuals");
    }
    void statistics_calc(double t, double out17[17]) {                                   // :178
        check(o3d_s_statistics(s_, t, out17), "o3d_s_statistics");
    }
    void write_all_data(const std::string& dir, int num) {                               // :147
        check(o3d_s_write_all_data(s_, dir.c_str(), num), "o3d_s_write_all_data");
    }
    void save_fields(const std::string& file, double time, const double* x, const double* y,
                     const double* z) {                                                  // :183
        check(o3d_s_save_fields(s_, file.c_str(), time, x, y, z), "o3d_s_save_fields");
    }
    void io_wait() { check(o3d_s_io_wait(s_), "o3d_s_io_wait"); }

private:
    o3d_session* s_;
};

}  // namespace o3d
