/* Host-only part of the C-ABI check: compiled with gcc against include/eigenexa_b200.h and linked to the
 * shared library.  Calls only entry points that need no GPU (index helpers in the by-value C form and in
 * the by-reference Fortran form of C/eigen_exa_interfaces.h:14-31, queries before eigen_init). */
#include <stdio.h>
#include <string.h>
#include "eigenexa_b200.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); fails++; } } while (0)

int main(void)
{
    int v = 0; char date[64] = "", vcode[64] = "";
    eigen_get_version(&v, date, vcode);
    CHECK(v == 21300);
    int v2 = 0; char d2[64] = "", c2[64] = "";
    eigen_libs0_eigen_get_version_(&v2, d2, c2);
    CHECK(v2 == v && strcmp(d2, date) == 0);

    /* not initialised: queries answer, nothing crashes */
    eigenexa_b200_comm_t w, x, y;
    eigen_get_comm(&w, &x, &y);
    CHECK(w.rank == -1 && x.rank == -1 && y.rank == -1);
    int cw = 7, cx = 7, cy = 7;
    eigen_libs0_eigen_get_comm_(&cw, &cx, &cy);
    CHECK(cw == -1 && cx == -1 && cy == -1);
    CHECK(eigen_blacs_eigen_get_blacs_context_() == -1);

    for (int nnod = 1; nnod <= 8; nnod++)
        for (int inod = 1; inod <= nnod; inod++)
            for (int i = 1; i <= 40; i++) {
                int a = i, b = nnod, c = inod;
                /* formulas of src/eigen_libs0.F:1816-2258 */
                CHECK(eigen_loop_start(i, nnod, inod) == (i + nnod - 1 - inod) / nnod + 1);
                CHECK(eigen_loop_end(i, nnod, inod) == (i + nnod - inod) / nnod);
                CHECK(eigen_libs0_eigen_loop_start_(&a, &b, &c) == eigen_loop_start(i, nnod, inod));
                CHECK(eigen_libs0_eigen_loop_end_(&a, &b, &c) == eigen_loop_end(i, nnod, inod));
                CHECK(eigen_libs0_eigen_translate_l2g_(&a, &b, &c) == (i - 1) * nnod + inod);
                CHECK(eigen_libs0_eigen_translate_g2l_(&a, &b, &c) == (i - 1) / nnod + 1);
                CHECK(eigen_libs0_eigen_owner_node_(&a, &b, &c) == (i - 1) % nnod + 1);
                CHECK(eigen_libs0_eigen_owner_index_(&a, &b, &c) == eigen_owner_index(i, nnod, inod));
                int ls = -9, le = -9, one = 1;
                eigen_libs0_eigen_loop_info_(&one, &a, &ls, &le, &b, &c);
                CHECK(ls == eigen_loop_start(1, nnod, inod) && le == eigen_loop_end(i, nnod, inod));
                int ls2, le2;
                eigen_loop_info(1, i, &ls2, &le2, nnod, inod);
                CHECK(ls2 == ls && le2 == le);
            }
    /* 1x1 grid before eigen_init: ids as the reference computes them (eigen_libs0.F:2316-2356) */
    int one = 1, xi = 0, yi = 0;
    eigen_libs0_eigen_convert_id_w2xy_(&one, &xi, &yi);
    CHECK(xi == 1 && yi == 1);
    CHECK(eigen_libs0_eigen_convert_id_xy2w_(&one, &one) == eigen_convert_id_xy2w(1, 1));
    int n = 1000, nx = 0, ny = 0, mf = 48, mb = 128;
    eigen_libs_eigen_get_matdims_(&n, &nx, &ny, &mf, &mb, "O");
    int nx2 = 0, ny2 = 0;
    eigen_get_matdims(n, &nx2, &ny2, 48, 128, "O");
    CHECK(nx == nx2 && ny == ny2 && nx >= n && ny >= n);
    int lda = 1000, ldz = 1000;
    CHECK(eigen_libs0_eigen_memory_internal_(&n, &lda, &ldz, &mf, &mb) > 0);
    CHECK(eigen_memory_internal(n, lda, ldz, mf, mb) > 0);
    /* eigen_s before eigen_init returns silently (src/eigen_s.F:81-84) */
    double a[4] = {-2, 1, 1, -2}, ww[2] = {7, 7}, z[4] = {0, 0, 0, 0};
    eigen_s(2, 2, a, 2, ww, z, 2, 1, 1, "A");
    CHECK(ww[0] == 7 && ww[1] == 7);
    eigen_free();
    printf(fails ? "CABI_HOST_FAIL %d\n" : "CABI_HOST_OK\n", fails);
    return fails != 0;
}
