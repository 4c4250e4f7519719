// mg_tables_dump.cu -- prints the multigrid transfer tables of the product's host-side builder
// (csrc/multigrid.cu: coarse_extent, axis_tables) as JSON, one line per (n, mode), so that
// tests/test_host_rules_cpu.py can compare them with the NumPy design model oracle/mg_model.py,
// whose tables are checked for partition of unity, transpose structure and V-cycle convergence in
// tests/test_mg_model_cpu.py.  Links against libo3d_b200.so for the symbols multigrid.cu refers
// to; makes no CUDA call.
#include <cstdio>

#include "../../osinco3d_b200/csrc/multigrid.cu"

using namespace o3d;

int main() {
    const int ns[] = {5, 6, 7, 8, 9, 16, 17, 32, 33, 81, 129, 185, 241, 256, 257, 513};
    for (int mode : {BM_WRAP, BM_MIRROR})
        for (int n : ns) {
            const double d = 0.0371;
            const int nc = coarse_extent(n, mode);
            printf("{\"n\": %d, \"mode\": %d, \"nc\": %d", n, mode, nc);
            if (nc) {
                const AxisTab a = axis_tables(n, d, mode, nc);
                printf(", \"D\": %.17g, \"c0\": [", a.D);
                for (int i = 0; i < n; ++i) printf("%s%d", i ? "," : "", a.c0[i]);
                printf("], \"w\": [");
                for (int i = 0; i < n; ++i) printf("%s%.17g", i ? "," : "", a.w[i]);
                printf("], \"ridx\": [");
                for (int i = 0; i < 4 * nc; ++i) printf("%s%d", i ? "," : "", a.ridx[i]);
                printf("], \"rw\": [");
                for (int i = 0; i < 4 * nc; ++i) printf("%s%.17g", i ? "," : "", a.rw[i]);
                printf("]");
            }
            printf("}\n");
        }
    return 0;
}
