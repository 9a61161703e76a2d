/* GPU part of the C-ABI check, mirroring the reference's C/c_test.c:19-44 (single process): the 2x2 matrix
 * [-2 1; 1 -2] must give eigenvalues (-3, -1) -- through the C entry point eigen_s and through the
 * Fortran-style by-reference symbol eigen_libs_eigen_s_ (C/EigenExa.fh:14, C/eigen_exa_interfaces.h:8). */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "eigenexa_b200.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); fails++; } } while (0)

static void check_pairs(const double *w, const double *z, const char *what)
{
    const double r = sqrt(0.5);
    printf("%s: %le :: %le %le\n", what, w[0], z[0], z[1]);
    printf("%s: %le :: %le %le\n", what, w[1], z[2], z[3]);
    CHECK(fabs(w[0] + 3.0) < 1e-14 && fabs(w[1] + 1.0) < 1e-14);
    /* eigenvectors (1,-1)/sqrt2 and (1,1)/sqrt2 up to sign */
    CHECK(fabs(fabs(z[0]) - r) < 1e-14 && fabs(fabs(z[1]) - r) < 1e-14 && z[0] * z[1] < 0);
    CHECK(fabs(fabs(z[2]) - r) < 1e-14 && fabs(fabs(z[3]) - r) < 1e-14 && z[2] * z[3] > 0);
}

int main(void)
{
    eigen_init(NULL, "C");
    int nnod = 0, xn = 0, yn = 0, inod = 0, xi = 0, yi = 0;
    eigen_get_procs(&nnod, &xn, &yn);
    eigen_get_id(&inod, &xi, &yi);
    eigenexa_b200_comm_t cw, cx, cy;
    eigen_get_comm(&cw, &cx, &cy);
    if (cw.rank < 0 || nnod != 1) { printf("eigen_init failed: %s\n", eigenexa_b200_last_error()); return 2; }
    CHECK(xn == 1 && yn == 1 && inod == 1 && xi == 1 && yi == 1);
    CHECK(cw.rank == 0 && cw.nranks == 1 && cx.nranks == 1 && cy.nranks == 1);

    int n = 2, nv = 2, lda = 2, ldz = 2, mf = 1, mb = 1;
    double a[4] = {-2, 1, 1, -2}, w[2] = {0, 0}, z[4] = {0, 0, 0, 0};
    eigen_s(n, nv, a, lda, w, z, ldz, mf, mb, "A");
    check_pairs(w, z, "eigen_s");
    /* a(1:3,1) = flop count, seconds, comm seconds (src/eigen_s.F:284-295) */
    CHECK(a[0] != -2.0 && a[1] >= 0.0);

    double a2[4] = {-2, 1, 1, -2}, w2[2] = {0, 0}, z2[4] = {0, 0, 0, 0};
    eigen_libs_eigen_s_(&n, &nv, a2, &lda, w2, z2, &ldz, &mf, &mb, "A");
    check_pairs(w2, z2, "eigen_libs_eigen_s_");

    double a3[4] = {-2, 1, 1, -2}, w3[2] = {0, 0}, z3[4] = {0, 0, 0, 0};
    eigen_libs_eigen_sx_(&n, &nv, a3, &lda, w3, z3, &ldz, &mf, &mb, "A");
    check_pairs(w3, z3, "eigen_libs_eigen_sx_");

    /* mode 'N' through the by-reference symbol: eigenvalues only, z untouched */
    double a4[4] = {-2, 1, 1, -2}, w4[2] = {0, 0}, z4[4] = {9, 9, 9, 9};
    int zero = 0;
    eigen_libs_eigen_s_(&n, &zero, a4, &lda, w4, z4, &ldz, &mf, &mb, "N");
    CHECK(fabs(w4[0] + 3.0) < 1e-13 && fabs(w4[1] + 1.0) < 1e-13 && z4[0] == 9);
    int info = 99;
    eigen_libs0_eigen_get_errinfo_(&info);
    CHECK(info == 0);
    eigen_libs_eigen_free_();
    printf(fails ? "CABI_GPU_FAIL %d\n" : "CABI_GPU_OK\n", fails);
    return fails != 0;
}
