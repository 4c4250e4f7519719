// hpp_mirror_test.cpp -- the C++ host-side mirror (include/o3d_b200.hpp) in a compiled program.
//   hpp_mirror_test nodevice
//       without a CUDA device: every module procedure throws o3d::Error(O3D_ERR_NO_DEVICE) (there
//       is no CPU fallback); schemes() throws O3D_ERR_BC on the flag combinations the reference
//       stops on (src/initialization.f90:238-242) -- that check needs no device
//   hpp_mirror_test run in.bin out.bin nx ny nz dx dy dz dt
//       on a device, all-free-slip boundaries: in = f | ux uy uz | pp ;
//       out = the 18 routines of src/derivation.f90 applied to f (x, y, z) x (order 1, 2) x
//             (_00, p_11, i_11) | divergence(ux,uy,uz, odd=1) | correct_velocity(ux,uy,uz as u*, pp)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/o3d_b200.hpp"

using o3d::Shape;

static int expect(int code, const char* what, void (*fn)()) {
    try {
        fn();
    } catch (const o3d::Error& e) {
        if (e.code() == code) return 0;
        printf("FAIL %s: code %d instead of %d (%s)\n", what, e.code(), code, e.what());
        return 1;
    }
    printf("FAIL %s: no exception\n", what);
    return 1;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "nodevice")) {
        int bad = 0;
        bad += expect(O3D_ERR_BC, "schemes(mixed x flags)", [] { o3d::initialization::schemes(0, 1, 1, 1, 0, 0); });
        bad += expect(O3D_ERR_BC, "schemes(Dirichlet x)", [] { o3d::initialization::schemes(2, 2, 1, 1, 0, 0); });
        o3d::initialization::schemes(1, 1, 1, 1, 1, 1);
        if (o3d_device_count() == 0) {
            static std::vector<double> f(512, 0.0), df(512);
            bad += expect(O3D_ERR_NO_DEVICE, "derx_00", [] {
                o3d::derivation::derx_00(df.data(), f.data(), 0.1, Shape{8, 8, 8});
            });
            bad += expect(O3D_ERR_NO_DEVICE, "divergence", [] {
                o3d::diffoper::divergence(df.data(), f.data(), f.data(), f.data(), 0.1, 0.1, 0.1, Shape{8, 8, 8}, 1);
            });
            bad += expect(O3D_ERR_NO_DEVICE, "Session", [] {
                o3d_config c;
                memset(&c, 0, sizeof(c));
                c.nx = c.ny = c.nz = 16, c.dx = c.dy = c.dz = 0.1;
                c.nbcx1 = c.nbcxn = c.nbcy1 = c.nbcyn = c.nbcz1 = c.nbczn = 1;
                c.re = 100, c.sc = 1, c.dt = 1e-3, c.itscheme = 3, c.omega = 1.8, c.eps = 1e-6, c.kmax = 10;
                o3d::Session s(c);
            });
        }
        printf(bad ? "%d failure(s)\n" : "hpp mirror OK\n", bad);
        return bad ? 1 : 0;
    }
    if (argc != 11 || strcmp(argv[1], "run")) return 2;
    const Shape s{atoi(argv[4]), atoi(argv[5]), atoi(argv[6])};
    const double d[3] = {atof(argv[7]), atof(argv[8]), atof(argv[9])};
    const double dt = atof(argv[10]);
    const size_t N = (size_t)s.nx * s.ny * s.nz;
    std::vector<double> in(5 * N), out(22 * N);
    FILE* fi = fopen(argv[2], "rb");
    if (!fi || fread(in.data(), 8, 5 * N, fi) != 5 * N) return 3;
    fclose(fi);
    const double *f = &in[0], *ux = &in[N], *uy = &in[2 * N], *uz = &in[3 * N], *pp = &in[4 * N];
    try {
        using namespace o3d::derivation;
        typedef void (*der_t)(double*, const double*, double, Shape);
        const der_t routines[18] = {derx_00,  derxp_11,  derxi_11,  derxx_00, derxxp_11, derxxi_11,
                                    dery_00,  deryp_11,  deryi_11,  deryy_00, deryyp_11, deryyi_11,
                                    derz_00,  derzp_11,  derzi_11,  derzz_00, derzzp_11, derzzi_11};
        for (int q = 0; q < 18; ++q) routines[q](&out[q * N], f, d[q / 6], s);
        o3d::initialization::schemes(1, 1, 1, 1, 1, 1);
        o3d::diffoper::divergence(&out[18 * N], ux, uy, uz, d[0], d[1], d[2], s, 1);
        o3d::integration::correct_velocity(&out[19 * N], &out[20 * N], &out[21 * N], ux, uy, uz, pp, dt,
                                           d[0], d[1], d[2], s);
    } catch (const o3d::Error& e) {
        printf("FAIL %s\n", e.what());
        return 1;
    }
    FILE* fo = fopen(argv[3], "wb");
    if (!fo || fwrite(out.data(), 8, 22 * N, fo) != 22 * N) return 4;
    fclose(fo);
    printf("hpp mirror run OK\n");
    return 0;
}
