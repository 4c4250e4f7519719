// o3d_mainloop.cpp -- the reference's time loop (src/osinco3d_main.f90:97-188) replayed over the
// C++ mirror of the module interfaces (include/o3d_b200.hpp), state resident on the device:
//
//   every step          old_values (:104, only where the residual is evaluated), predict_velocity +
//                       correct_pression + correct_velocity [+ transeq] (:105-115), the divergence /
//                       velocity / CFL prints (:116-128) as device reductions
//   every nfre steps    write_all_data (:130-150)            -> <out>/outputs/ux_N.bin ...
//   every 25 steps      calculate_residuals (:167) and, when time > initstat, statistics_calc (:178)
//                       -> <out>/outputs/stats.dat, one '(17es21.12)' line (src/IOfunctions.f90:552)
//   every nsve steps    save_fields (:183)                   -> <out>/fields_NNNNNN.bin
//
// It is the harness SURVEY.md section 7 asks for: a driver-shaped caller that proves the pieces
// (o3d_step, o3d_s_step_diagnostics, o3d_s_statistics, the asynchronous writers) run together at
// the reference's cadence, and whose stats.dat can be diffed against the reference's shipped
// histories (tests/test_gpu_mainloop.py).  Initial conditions are the Taylor-Green vortex of
// src/initial_conditions.f90:141-153, generated on the host (input generation, not hot path).
//
//   o3d_mainloop --n 185 --steps 100 --out DIR [--re 1600] [--cfl 0.05 | --dt 5e-4]
//                [--omega 1.887] [--eps 1e-4] [--idyn 0] [--kmax 10000] [--les CS] [--nscr 1]
//                [--nfre 50] [--nsve 100] [--initstat -1] [--stats-at-start] [--wavefront]
//                [--quiet]
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/o3d_b200.hpp"

namespace {

struct Opt {
    int n = 64, steps = 50, nfre = 0, nsve = 0, idyn = 0, kmax = 10000, nscr = 0, quiet = 0;
    int stats_at_start = 0, wavefront = 0;
    double re = 1600.0, cfl = 0.05, dt = 0.0, omega = 1.887, eps = 1e-4, cs = 0.0, initstat = -1.0;
    std::string out = "o3d_run";
};

double arg_d(int& i, int argc, char** argv) {
    if (i + 1 >= argc) {
        fprintf(stderr, "missing value after %s\n", argv[i]);
        exit(2);
    }
    return atof(argv[++i]);
}

// the '(17es21.12)' record of write_statistics, src/IOfunctions.f90:552
void write_stats_row(const std::string& path, const double* v17) {
    FILE* f = fopen(path.c_str(), "a");
    if (!f) {
        printf(" Error: Unable to open the file 'outputs/stats.dat'\n");
        return;
    }
    for (int c = 0; c < 17; ++c) fprintf(f, "%21.12E", v17[c]);
    fputc('\n', f);
    fclose(f);
}

}  // namespace

int main(int argc, char** argv) {
    Opt o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--n") o.n = (int)arg_d(i, argc, argv);
        else if (a == "--steps") o.steps = (int)arg_d(i, argc, argv);
        else if (a == "--re") o.re = arg_d(i, argc, argv);
        else if (a == "--cfl") o.cfl = arg_d(i, argc, argv);
        else if (a == "--dt") o.dt = arg_d(i, argc, argv);
        else if (a == "--omega") o.omega = arg_d(i, argc, argv);
        else if (a == "--eps") o.eps = arg_d(i, argc, argv);
        else if (a == "--idyn") o.idyn = (int)arg_d(i, argc, argv);
        else if (a == "--kmax") o.kmax = (int)arg_d(i, argc, argv);
        else if (a == "--les") o.cs = arg_d(i, argc, argv);
        else if (a == "--nscr") o.nscr = (int)arg_d(i, argc, argv);
        else if (a == "--nfre") o.nfre = (int)arg_d(i, argc, argv);
        else if (a == "--nsve") o.nsve = (int)arg_d(i, argc, argv);
        else if (a == "--initstat") o.initstat = arg_d(i, argc, argv);
        else if (a == "--stats-at-start") o.stats_at_start = 1;
        else if (a == "--wavefront") o.wavefront = 1;
        else if (a == "--quiet") o.quiet = 1;
        else if (a == "--out") {
            if (i + 1 >= argc) return 2;
            o.out = argv[++i];
        } else {
            fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        }
    }
    const int n = o.n;
    const double pi = 3.141592653589793;
    const double d = pi / (n - 1);                    // dx = xlx/(nx-1), src/initialization.f90:182
    const double dt = o.dt > 0.0 ? o.dt : o.cfl * d;  // dt = cfl*dmin/u0
    mkdir(o.out.c_str(), 0755);
    const std::string outputs = o.out + "/outputs";
    mkdir(outputs.c_str(), 0755);
    const std::string stats_path = outputs + "/stats.dat";
    remove(stats_path.c_str());

    o3d_config c;
    memset(&c, 0, sizeof(c));
    c.nx = c.ny = c.nz = n;
    c.dx = c.dy = c.dz = d;
    c.nbcx1 = c.nbcxn = c.nbcy1 = c.nbcyn = c.nbcz1 = c.nbczn = O3D_FREE_SLIP;
    c.re = o.re, c.sc = 1.0, c.cs = o.cs;
    c.delta = std::cbrt(d * d * d);                   // src/initialization.f90:193
    c.dt = dt;
    // src/initialization.f90:194-202, written as the reference writes them
    c.adt[0] = dt, c.bdt[0] = 0.0, c.cdt[0] = 0.0;
    c.adt[1] = 3.0 * dt / 2.0, c.bdt[1] = -dt / 2.0, c.cdt[1] = 0.0;
    c.adt[2] = 23.0 * dt / 12.0, c.bdt[2] = -16.0 * dt / 12.0, c.cdt[2] = 5.0 * dt / 12.0;
    c.itscheme = 3, c.iles = o.cs > 0.0 ? 1 : 0, c.nscr = o.nscr;
    c.omega = o.omega, c.eps = o.eps, c.kmax = o.kmax, c.idyn = o.idyn, c.multigrid = 0;
    c.sor_order = o.wavefront ? O3D_SOR_LEXI_WAVEFRONT : O3D_SOR_RED_BLACK;
    c.rank = 0, c.nranks = 1;

    try {
        o3d::Session ses(c);
        const size_t N = (size_t)n * n * n;
        std::vector<double> x(n), f(N);
        for (int i = 0; i < n; ++i) x[i] = d * i;
        auto fill = [&](int which) {  // src/initial_conditions.f90:141-153
            for (int k = 0; k < n; ++k)
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < n; ++i) {
                        double v = 0.0;
                        if (which == 0) v = std::sin(x[i]) * std::cos(x[j]) * std::cos(x[k]);
                        else if (which == 1) v = -std::cos(x[i]) * std::sin(x[j]) * std::cos(x[k]);
                        else if (which == 3)
                            v = 0.0625 * (std::cos(2.0 * x[i]) + std::cos(2.0 * x[j])) *
                                (std::cos(2.0 * x[k]) + 2.0);
                        f[(size_t)i + (size_t)n * (j + (size_t)n * k)] = v;
                    }
        };
        const int ids[4] = {O3D_F_UX, O3D_F_UY, O3D_F_UZ, O3D_F_PP};
        for (int q = 0; q < 4; ++q) {
            fill(q);
            ses.upload(ids[q], f.data());
        }
        if (o.nscr) {
            for (size_t q = 0; q < N; ++q) f[q] = 0.0;
            ses.upload(O3D_F_PHI, f.data());
        }
        std::vector<double>().swap(f);

        double st[17], diag[23], res[15];
        int num = 0, numx = 0;
        if (o.stats_at_start) {
            ses.statistics_calc(0.0, st);
            write_stats_row(stats_path, st);
        }
        if (o.nfre > 0) ses.write_all_data(outputs, numx++);
        long long total_iters = 0;
        const auto t_go = std::chrono::steady_clock::now();
        const double time0 = 0.0, t_ref = 1.0, u_ref = 1.0;  // l0 = u0 = 1
        for (int itime = 1; itime <= o.steps; ++itime) {
            const double time = time0 + itime * dt;  // :99
            if (!o.quiet) {
                printf(" ========================\n Iteration: %6d/%6d\n TIME = %10.3f/%6.0f\n"
                       " ========================\n", itime, o.steps, time, time0 + o.steps * dt);
            }
            // old_values (:104) feeds calculate_residuals only, which runs every 25 steps (:167)
            if (itime % 25 == 0) ses.old_values();
            double dmax = 0.0;
            const int iters = ses.step(itime, &dmax);  // :105-115
            total_iters += iters;
            ses.step_diagnostics(diag);                // :116-128
            if (!o.quiet) {
                printf(" * SOR iterations %d, dmax %.6e\n", iters, dmax);
                printf(" * div(u*) min/max/mean %.6e %.6e %.6e\n", diag[0], diag[1], diag[2]);
                printf(" * div(u)  min/max/mean %.6e %.6e %.6e at (%d,%d,%d)\n", diag[6], diag[7],
                       diag[8], (int)diag[9], (int)diag[10], (int)diag[11]);
                printf(" * ux %.6e %.6e uy %.6e %.6e uz %.6e %.6e\n", diag[12], diag[15], diag[13],
                       diag[16], diag[14], diag[17]);
                printf(" * CFL %.6e %.6e %.6e\n", diag[18], diag[19], diag[20]);
            }
            if (o.nfre > 0 && itime % o.nfre == 0) {   // :130-150
                ses.write_all_data(outputs, numx++);
                ++num;
            }
            if (itime % 25 == 0) {                      // :167-181
                ses.calculate_residuals(dt, t_ref, u_ref, res);
                if (!o.quiet)
                    printf(" * residuals %.6e %.6e %.6e\n", res[0], res[1], res[2]);
                if (res[0] > 1.0e6) {                   // src/utils.f90:155-158
                    printf(" Residue too high\n");
                    return 3;
                }
                if (time > o.initstat) {
                    ses.statistics_calc(time, st);
                    write_stats_row(stats_path, st);
                }
            }
            if (o.nsve > 0 && itime % o.nsve == 0) {    // :183
                char name[64];
                snprintf(name, sizeof(name), "/fields_%06d.bin", itime);
                ses.save_fields(o.out + name, time, x.data(), x.data(), x.data());
            }
        }
        ses.io_wait();
        ses.sync();
        const double sec =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_go).count();
        printf("o3d_mainloop: %d steps of %d^3 in %.3f s (%.1f Mpts*steps/s incl. diagnostics and "
               "output), %.2f SOR iterations/step\n", o.steps, n, sec,
               (double)N * o.steps / sec / 1e6, (double)total_iters / o.steps);
    } catch (const o3d::Error& e) {
        if (e.code() == O3D_ERR_DIVERGED) {
            // write_velocity_diverged + stop, src/integration.f90:309-325
            printf(" Velocity diverged: %s\n", e.what());
            return 4;
        }
        fprintf(stderr, "o3d_mainloop: %s\n", e.what());
        return 1;
    }
    return 0;
}
