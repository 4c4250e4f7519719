// host_copier_test.cpp -- the host-side history shift of the pipelined o3d_predict_velocity
// (csrc/host_copier.h, csrc/pipeline.cu) with the device replaced by two plain threads, so that
// the ORDER PROTOCOL between the DMA transfers and the worker threads can be checked on the CPU --
// under -fsanitize=thread any host array touched by a "DMA" thread and a worker without an
// event / flag between them is reported as a data race.
//
//   upload thread  : for j = 0..C-1: reads the host levels 2 and 3 of chunk j into "device"
//                    buffers (the H2D copies), then signals up_event[j] (both lanes)
//   compute thread : for the chunks in issue order (chunk c needs upload need(c), a periodic z
//                    axis sends chunk 0 behind the last upload -- Plan of pipeline.cu): new f =
//                    g(level 2, level 3) on the device copies, written to host level 1 (the D2H
//                    copy), then signals l1_event[c][k]
//   main thread    : pushes the jobs in the order pipe_predict_velocity does
// Expected afterwards (src/integration.f90:176-188):
//   itscheme 3: level 3 = old level 2, level 2 = level 1 = new;  itscheme 2: level 2 = level 1 =
//   new, level 3 untouched;  itscheme 1: levels 2, 3 untouched;  nu_t = 0.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../osinco3d_b200/csrc/host_copier.h"

using o3d::HostCopier;

static int wait_flag(void* ev) {
    auto* f = static_cast<std::atomic<int>*>(ev);
    while (!f->load(std::memory_order_acquire)) std::this_thread::yield();
    return 0;
}

static double g(double l2, double l3, int k) { return 2.0 * l2 - 0.5 * l3 + k + 1.0; }

static int run_case(int C, long long plane, int nz, int itscheme, bool wrapz, int threads) {
    const long long N = plane * nz;
    std::vector<int> z(C + 1), need(C);
    for (int c = 0; c <= C; ++c) z[c] = (int)((long long)nz * c / C);
    for (int c = 0; c < C; ++c) need[c] = (c + 1 < C) ? c + 1 : C - 1;
    if (wrapz) need[0] = C - 1;

    std::vector<double> f[3], old2[3], old3[3], dev2[3], dev3[3], nu_t(N, 7.0);
    for (int k = 0; k < 3; ++k) {
        f[k].resize(3 * N);
        for (long long i = 0; i < 3 * N; ++i) f[k][i] = 0.001 * (double)((i * 7 + k * 13) % 1009);
        old2[k].assign(f[k].begin() + N, f[k].begin() + 2 * N);
        old3[k].assign(f[k].begin() + 2 * N, f[k].end());
        dev2[k].assign(N, -1.0), dev3[k].assign(N, -1.0);
    }
    std::vector<std::atomic<int>> up_ev(2 * C), l1_ev(3 * C), a_done(3 * C);
    for (auto& e : up_ev) e.store(0);
    for (auto& e : l1_ev) e.store(0);
    for (auto& e : a_done) e.store(0);

    HostCopier hc(threads, wait_flag);
    // the "device"
    std::thread uploader([&] {
        for (int j = 0; j < C; ++j) {
            for (int k = 0; k < 3; ++k)
                for (long long i = plane * z[j]; i < plane * z[j + 1]; ++i) {
                    dev2[k][i] = f[k][N + i];
                    dev3[k][i] = f[k][2 * N + i];
                }
            up_ev[2 * j].store(1, std::memory_order_release);
            up_ev[2 * j + 1].store(1, std::memory_order_release);
        }
    });
    std::thread computer([&] {
        std::vector<char> issued(C, 0);
        for (int j = 0; j < C; ++j) {
            wait_flag(&up_ev[2 * j]);
            wait_flag(&up_ev[2 * j + 1]);
            for (int c = 0; c < C; ++c) {
                if (issued[c] || need[c] > j) continue;
                issued[c] = 1;
                for (int k = 0; k < 3; ++k) {
                    for (long long i = plane * z[c]; i < plane * z[c + 1]; ++i)
                        f[k][i] = g(dev2[k][i], dev3[k][i], k);  // D2H of level 1
                    l1_ev[3 * c + k].store(1, std::memory_order_release);
                }
            }
        }
    });
    // the host side of pipe_predict_velocity
    {
        std::vector<char> issued(C, 0);
        for (int c = 0; c < C; ++c)
            hc.push(o3d::zero_job(nu_t.data(), plane * z[c],
                                  (size_t)(plane * (z[c + 1] - z[c])) * sizeof(double)));
        for (int j = 0; j < C; ++j) {
            const long long off = plane * z[j];
            const size_t bytes = (size_t)(plane * (z[j + 1] - z[j])) * sizeof(double);
            if (itscheme == 3)
                for (int k = 0; k < 3; ++k)
                    hc.push(o3d::shift_job_a(f[k].data(), N, off, bytes, &up_ev[2 * j],
                                             &up_ev[2 * j + 1], &a_done[3 * j + k]));
            for (int c = 0; c < C; ++c) {
                if (issued[c] || need[c] > j) continue;
                issued[c] = 1;
                if (itscheme != 2 && itscheme != 3) continue;
                for (int k = 0; k < 3; ++k)
                    hc.push(o3d::shift_job_b(
                        f[k].data(), N, plane * z[c],
                        (size_t)(plane * (z[c + 1] - z[c])) * sizeof(double), &l1_ev[3 * c + k],
                        itscheme == 3 ? &a_done[3 * c + k] : nullptr));
            }
        }
    }
    uploader.join();
    computer.join();
    if (hc.drain() != 0) return 1;

    int bad = 0;
    for (int k = 0; k < 3; ++k)
        for (long long i = 0; i < N; ++i) {
            const double nw = g(old2[k][i], old3[k][i], k);
            const double e1 = nw;
            const double e2 = (itscheme == 2 || itscheme == 3) ? nw : old2[k][i];
            const double e3 = (itscheme == 3) ? old2[k][i] : old3[k][i];
            if (f[k][i] != e1 || f[k][N + i] != e2 || f[k][2 * N + i] != e3) ++bad;
        }
    for (long long i = 0; i < N; ++i)
        if (nu_t[i] != 0.0) ++bad;
    if (bad)
        std::printf("case C=%d itscheme=%d wrapz=%d threads=%d: %d wrong values\n", C, itscheme,
                    (int)wrapz, threads, bad);
    return bad != 0;
}

int main() {
    int fails = 0, cases = 0;
    const int threads[] = {1, 2, 5};
    for (int rep = 0; rep < 3; ++rep)
        for (int C : {2, 3, 8, 16})
            for (int its : {1, 2, 3})
                for (int wrap = 0; wrap < 2; ++wrap)
                    for (int t : threads) {
                        fails += run_case(C, 300 + 17 * rep, 8 * C + rep + 3, its, wrap != 0, t);
                        ++cases;
                    }
    // error propagation: a failing wait is reported by drain(), dependants still released
    {
        HostCopier hc(2, [](void*) -> int { return 42; });
        std::atomic<int> done{0};
        double a[4] = {1, 2, 3, 4}, b[4] = {0, 0, 0, 0};
        int dummy;
        o3d::HostJob j1;
        j1.ev[0] = &dummy, j1.dst = b, j1.src = a, j1.bytes = sizeof(a), j1.done = &done;
        o3d::HostJob j2;
        j2.after = &done, j2.dst = a, j2.bytes = sizeof(a);
        hc.push(j1);
        hc.push(j2);
        if (hc.drain() != 42 || b[3] != 4.0 || a[0] != 0.0) ++fails;
        if (hc.drain() != 0) ++fails;  // the status is reported once
        ++cases;
    }
    if (fails) {
        std::printf("host copier FAILED: %d of %d cases\n", fails, cases);
        return 1;
    }
    std::printf("host copier OK (%d cases)\n", cases);
    return 0;
}
