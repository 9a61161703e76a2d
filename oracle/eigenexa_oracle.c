/*
 * eigenexa_oracle.c -- CPU restatement of the EigenExa eigen_s hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * product in eigenexa_b200/csrc.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * library never links, imports or calls anything in oracle/.
 *
 * It restates, on a full (1x1 grid) column-major matrix, the arithmetic of the
 * reference (RIKEN-RCCS/EigenExa v2.13, /root/reference):
 *   index algebra       src/eigen_libs0.F:1816-2258
 *   eigen_get_matdims0  src/eigen_libs0.F:1254-1371, CSTAB_get_optdim src/CSTAB.F:73-131
 *   eigen_scaling       src/eigen_scaling.F:59-154
 *   eigen_trd           src/eigen_trd.F:349-723, eigen_trd_t2.F:352-614 (deferred
 *                       normalisation), eigen_trd_t6_3.F:255-281, eigen_trd_t8.F:188-219,
 *                       eigen_t1.F:250-306 (rank-2k update)
 *   eigen_common_trbakwy src/trbakwy4.F:299-336,345-499,538-602,
 *                       src/trbakwy4_body.F:206-213,302-313,504-741
 *   eigen_bisect        src/bisect.F:67-358 (Sturm-count bisection; restated as plain
 *                       bisection on the Gershgorin interval)
 *
 * Parity pinning: tests/test_oracle_*.py check this file against the reference's own
 * known answers (Frank spectrum benchmark/mat_set.f:638-647, Helmert families
 * :651-712, C/c_test.c 2x2 case, ev_test/w_test thresholds) and against LAPACK
 * dsytrd('U')/dstevd from SciPy's OpenBLAS.  The reference itself cannot be built in
 * this image (no Fortran compiler / MPI / ScaLAPACK), see DESIGN.md.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define A_(j, i) a[((size_t)(i) - 1) * lda + ((j) - 1)] /* 1-based (row j, col i) */

/* ------------------------------------------------------------------ */
/* index algebra (1-based, as in the reference)                        */
/* ------------------------------------------------------------------ */
int ora_loop_start(int istart, int nnod, int inod) { return (istart + nnod - 1 - inod) / nnod + 1; }
int ora_loop_end(int iend, int nnod, int inod) { return (iend + nnod - inod) / nnod; }
int ora_translate_l2g(int ictr, int nnod, int inod) { return (ictr - 1) * nnod + inod; }
int ora_translate_g2l(int ictr, int nnod, int inod) { (void)inod; return (ictr - 1) / nnod + 1; }
int ora_owner_node(int ictr, int nnod, int inod) { (void)inod; return (ictr - 1) % nnod + 1; }
int ora_owner_index(int ictr, int nnod, int inod)
{
    int j2 = ora_loop_start(ictr, nnod, inod), j3 = ora_loop_end(ictr, nnod, inod);
    return j2 == j3 ? j2 : -1;
}

/* grid shape chosen by eigen_init for P processes: src/eigen_libs0.F:526-540 */
void ora_grid_dims(int nnod, int *x_nnod, int *y_nnod)
{
    int x = (int)sqrt((double)nnod);
    int k = 1;
    for (;;) {
        if (x <= k) break;
        if (x % k == 0 && nnod % x == 0) break;
        x--;
    }
    if (x < 1) x = 1;
    *x_nnod = x;
    *y_nnod = nnod / x;
}

/* rank (1-based inod) -> (x_inod, y_inod): src/eigen_libs0.F:553-570 */
void ora_grid_coords(int inod, int x_nnod, int y_nnod, char order, int *x_inod, int *y_inod)
{
    if (order == 'R' || order == 'r') {
        *x_inod = (inod - 1) / y_nnod + 1;
        *y_inod = (inod - 1) % y_nnod + 1;
    } else {
        *x_inod = (inod - 1) % x_nnod + 1;
        *y_inod = (inod - 1) / x_nnod + 1;
    }
}

/* CSTAB_get_optdim, A64FX cache geometry of src/CSTAB.h */
static int cstab_get_optdim(int n_min, int n_unroll, int delta_L1, int delta_L2)
{
    const int L1_SIZE = 64 * 1024, L1_WAY = 4, L1_LINE = 256;
    const int L2_SIZE = 8 * 1024 * 1024, L2_WAY = 16;
    const int L1_LSIZE = (L1_SIZE / L1_WAY) / 8, L1_WINDOW = L1_LINE / 8;
    const int L2_LSIZE = (L2_SIZE / L2_WAY) / 8;
    int n_opt = n_min;
    for (;;) {
        int n_delta = 0, i, k, found = 0;
        n_opt = (n_opt - 1) / L1_WINDOW + 1;
        n_opt = (n_opt / 2) * 2 + 1;
        n_opt = n_opt * L1_WINDOW;
        int lim1 = (int)((n_unroll * 1.2 - 1.0) / L1_WAY + 1);
        for (i = 1; i <= lim1 && !found; i++) {
            k = (i * n_opt + L1_LSIZE / 2) % L1_LSIZE - L1_LSIZE / 2;
            if (abs(k) <= delta_L1 / 2) { n_delta = (delta_L1 / 2 - k - 1) / i + 1; found = 1; }
        }
        int lim2 = (int)((n_unroll * 1.2 - 1.0) / L2_WAY + 1);
        for (i = 1; i <= lim2 && !found; i++) {
            k = (i * n_opt + L2_LSIZE / 2) % L2_LSIZE - L2_LSIZE / 2;
            if (abs(k) <= delta_L2 / 2) { n_delta = (delta_L2 / 2 - k - 1) / i + 1; found = 1; }
        }
        if (n_delta == 0) break;
        n_opt += n_delta;
    }
    return n_opt;
}

/* eigen_get_matdims0 + the FS_get_matdims max of eigen_libs.F:138-143 (fs=1). */
void ora_get_matdims(int n, int x_nnod, int y_nnod, int m_f, int m_b, char mode,
                     int fs, int *nx_out, int *ny_out)
{
    (void)m_f;
    int nx, ny;
    if (n <= 0) { *nx_out = -1; *ny_out = -1; return; }
    if (mode == 'M') {
        nx = (n - 1) / x_nnod + 1; ny = (n - 1) / y_nnod + 1;
    } else if (mode == 'L') {
        nx = (n - 1) / x_nnod + 1; nx = ((nx - 1) / 32 + 1) * 32; ny = (n - 1) / y_nnod + 1;
    } else {
        const int eigen_NB = 64;
        int NPROW = x_nnod, NPCOL = y_nnod;
        int n1 = (n - 1) / NPROW + 1;
        int nm = cstab_get_optdim(n1, 6, 16 * 4, 16 * 4 * 2);
        int NB = m_b > eigen_NB ? m_b : eigen_NB;
        int nmz = (n - 1) / NPROW + 1; nmz = ((nmz - 1) / NB + 1) * NB + 1;
        int nn = nmz; nmz = (n - 1) / NB + 1; nmz = ((nmz - 1) / NPROW + 1) * NB; if (nn > nmz) nmz = nn;
        int nmw = (n - 1) / NPCOL + 1; nmw = ((nmw - 1) / NB + 1) * NB + 1;
        nn = nmw; nmw = (n - 1) / NB + 1; nmw = ((nmw - 1) / NPCOL + 1) * NB; if (nn > nmw) nmw = nn;
        int larray = (nmz > nm ? nmz : nm) * nmw;
        nx = nm; ny = (larray - 1) / nm + 1;
        NB = eigen_NB < n ? eigen_NB : n;
        int mn = NPCOL < NPROW ? NPCOL : NPROW;
        int64_t lddz = (n - 1) / mn + 1; lddz = ((lddz - 1) / NB + 1) * NB;
        int64_t nxx = (n - 1) / mn + 1;
        if (lddz * lddz >= ((int64_t)1 << 31) || nxx * nxx >= ((int64_t)1 << 31)) { nx = -1; ny = -1; }
    }
    if (fs && nx > 0) {
        /* FS grid = largest 2^p processes (src/FS_libs.F90:183-192,356-375) */
        int nnod = x_nnod * y_nnod, p = 1;
        while (p * 2 <= nnod) p *= 2;
        int fx, fy; ora_grid_dims(p, &fx, &fy);
        int n1 = n / p; if (n % p) n1++;
        int nx0 = n1 * (p / fx), ny0 = n1 * (p / fy);
        if (nx0 > nx) nx = nx0;
        if (ny0 > ny) ny = ny0;
    }
    *nx_out = nx; *ny_out = ny;
}

/* ------------------------------------------------------------------ */
/* eigen_scaling: src/eigen_scaling.F:76-134                           */
/* returns sigma (NaN when the upper triangle holds a non-finite value) */
/* ------------------------------------------------------------------ */
double ora_scaling(int n, double *a, int lda)
{
    const double SAFMIN = DBL_MIN, EPS = DBL_EPSILON * 0.5; /* DLAMCH('S'), DLAMCH('P')=eps*base/2.. */
    /* LAPACK: DLAMCH('Precision') = eps*base = 2^-52 */
    const double PREC = EPS * 2.0;
    const double SMLNUM = SAFMIN / PREC, BIGNUM = 1.0 / SMLNUM;
    const double RMIN = sqrt(SMLNUM);
    double RMAX = sqrt(BIGNUM), t2 = 1.0 / sqrt(sqrt(SAFMIN));
    if (t2 < RMAX) RMAX = t2;
    double anrm = 0.0; int bad = 0;
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= i; j++) {
            double t = A_(j, i);
            if (isfinite(t)) { if (fabs(t) > anrm) anrm = fabs(t); } else bad = 1;
        }
    if (bad) return NAN;
    double sigma = 1.0;
    if (anrm != 0.0 && anrm < RMIN) sigma = RMIN / anrm;
    else if (anrm > RMAX) sigma = RMAX / anrm;
    if (sigma == 1.0) return sigma;
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= i; j++) A_(j, i) *= sigma;
    return sigma;
}

/* ------------------------------------------------------------------ */
/* eigen_trd: blocked Householder tridiagonalisation, upper triangle   */
/* On exit: d(1:n), e(1:n) (e(i) couples i-1 and i, e(1)=0); column i  */
/* (i>=3) of a holds the reflector u_i in rows 1..i-1, a(1,2)=2*a(1,2). */
/* ------------------------------------------------------------------ */
static double sign_(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

/* symmetric access to the panel-start matrix: upper triangle of a, with
 * the diagonal taken from dt (panel_load pulls it out, src/eigen_trd_t7.F:72-169) */
void ora_trd(int n, double *a, int lda, double *d, double *e, int m_in)
{
    if (n <= 0) return;
    for (int i = 1; i <= n; i++) { d[i - 1] = 0.0; e[i - 1] = 0.0; }
    if (n == 1) { d[0] = A_(1, 1); return; }
    int m = m_in < n ? m_in : n; if (m < 1) m = 1;
    int mm = (n - 1) / m + 1;
    size_t nn = (size_t)n;
    double *W = (double *)calloc(nn * m, sizeof(double));  /* panel copy (u_x buffer)       */
    double *U = (double *)calloc(nn * m, sizeof(double));  /* finished reflectors of panel  */
    double *V = (double *)calloc(nn * m, sizeof(double));
    double *p = (double *)calloc(nn, sizeof(double));
    double *st = (double *)calloc(2 * (size_t)m, sizeof(double));
    double *pt = NULL; size_t pt_cap = 0; int pt_n = 1;
    int blk_stop = 3 * (2 - m) > 1 ? 3 * (2 - m) : 1;
    for (int i_block = mm; i_block >= blk_stop; i_block--) {
        int i_base = (i_block - 1) * m;
        int m0 = m < n - i_base ? m : n - i_base;
        int rows = i_base + m0;
        /* panel load: W(:,k) = A(1:i_base+k, i_base+k) incl. diagonal */
        for (int k = 1; k <= m0; k++)
            for (int j = 1; j <= i_base + k; j++) W[(size_t)(k - 1) * nn + (j - 1)] = A_(j, i_base + k);
        int k3 = 3 * (2 - i_block) > 1 ? 3 * (2 - i_block) : 1;
        for (int k1 = m0; k1 >= k3; k1--) {
            int i = i_base + k1, L = i - 1;
            double *w = W + (size_t)(k1 - 1) * nn; /* current column, all updates applied */
            d[i - 1] = w[i - 1];
            /* --- SYMV on the raw column (deferred normalisation, trd_t2:352-368) --- */
            double anorm2 = 0.0, prod_uv = 0.0, prod_ua = 0.0;
            double a_n = w[L - 1];
            double d_n = A_(L, L);
            /* one sweep over the stored upper triangle: every element is read once and
             * used twice (column dot + row axpy), as eigen_trd_au_body1 does
             * (trd_t2.F:970-1400); per-thread partial vectors summed in thread order. */
#pragma omp parallel
            {
                int nt = 1, tid = 0;
#ifdef _OPENMP
                nt = omp_get_num_threads(); tid = omp_get_thread_num();
#endif
#pragma omp single
                {
                    if (pt_cap < (size_t)nt * nn) { free(pt); pt = (double *)malloc((size_t)nt * nn * sizeof(double)); pt_cap = (size_t)nt * nn; }
                    pt_n = nt;
                }
                double *q = pt + (size_t)tid * nn;
                for (int j = 0; j < L; j++) q[j] = 0.0;
                for (int j = 1 + tid; j <= L; j += nt) {
                    const double *col = &A_(1, j);
                    double wj = w[j - 1], s = 0.0;
                    for (int c = 1; c < j; c++) { s += col[c - 1] * w[c - 1]; q[c - 1] += col[c - 1] * wj; }
                    q[j - 1] += s + col[j - 1] * wj;
                }
#pragma omp barrier
#pragma omp for schedule(static)
                for (int j = 0; j < L; j++) {
                    double s = 0.0;
                    for (int t = 0; t < pt_n; t++) s += pt[(size_t)t * nn + j];
                    p[j] = s;
                }
            }
            for (int j = 1; j <= L; j++) {
                anorm2 += w[j - 1] * w[j - 1];
                prod_uv += w[j - 1] * p[j - 1];
            }
            for (int j = 1; j < L; j++) prod_ua += w[j - 1] * A_(j, L);
            prod_ua += d_n * w[L - 1];
            /* --- Householder scalars (trd_t2:574-614) --- */
            double g_n, u_n, beta;
            if (anorm2 != 0.0) {
                double nrm = sqrt(anorm2);
                g_n = -sign_(nrm, a_n); u_n = a_n - g_n; beta = -u_n * g_n;
            } else { g_n = 0.0; u_n = 0.0; beta = 1.0; }
            e[i - 1] = g_n;
            /* p = A u = p0 - g * A(:,L) */
            for (int j = 1; j < L; j++) p[j - 1] -= g_n * A_(j, L);
            p[L - 1] -= g_n * d_n;
            prod_uv = prod_uv + g_n * (g_n * d_n - 2.0 * prod_ua);
            w[L - 1] = u_n; /* w(1:L) is now u */
            /* --- panel corrections (trd_t2:695-748, t6_3:160-250) --- */
            int ndone = m0 - k1;
            for (int l = 0; l < ndone; l++) {
                const double *ul = U + (size_t)(k1 + l) * nn, *vl = V + (size_t)(k1 + l) * nn;
                double s = 0.0, t = 0.0;
                for (int j = 1; j <= L; j++) { s += vl[j - 1] * w[j - 1]; t += ul[j - 1] * w[j - 1]; }
                st[2 * l] = s; st[2 * l + 1] = t;
            }
            double corr = 0.0;
            for (int l = 0; l < ndone; l++) {
                const double *ul = U + (size_t)(k1 + l) * nn, *vl = V + (size_t)(k1 + l) * nn;
                double s = st[2 * l], t = st[2 * l + 1];
                for (int j = 1; j <= L; j++) p[j - 1] -= ul[j - 1] * s + vl[j - 1] * t;
                corr += s * t;
            }
            prod_uv -= 2.0 * corr;
            /* --- v = (p - alpha u)/beta (t6_3:255-281) --- */
            double alpha = prod_uv / (2.0 * beta);
            double *uk = U + (size_t)(k1 - 1) * nn, *vk = V + (size_t)(k1 - 1) * nn;
            for (int j = 1; j <= L; j++) { uk[j - 1] = w[j - 1]; vk[j - 1] = (p[j - 1] - alpha * w[j - 1]) / beta; }
            for (int j = L + 1; j <= n; j++) { uk[j - 1] = 0.0; vk[j - 1] = 0.0; }
            /* --- lazy update of the remaining panel columns (trd_t5.F, t5x.F) --- */
            for (int c = 1; c < k1; c++) {
                int gc = i_base + c; /* global column; rows 1..gc are live (incl. diagonal) */
                double *wc = W + (size_t)(c - 1) * nn;
                double uc = uk[gc - 1], vc = vk[gc - 1];
                for (int j = 1; j <= gc; j++) wc[j - 1] -= uk[j - 1] * vc + vk[j - 1] * uc;
            }
        }
        /* panel restore: processed columns <- reflectors; untouched columns <- updated W */
        for (int k = 1; k <= m0; k++) {
            int gc = i_base + k;
            const double *src = (k >= k3) ? U + (size_t)(k - 1) * nn : W + (size_t)(k - 1) * nn;
            int top = (k >= k3) ? gc - 1 : gc;
            for (int j = 1; j <= top; j++) A_(j, gc) = src[j - 1];
        }
        (void)rows;
        /* rank-2k update of the trailing matrix (eigen_t1.F:250-306) */
        if (i_block > 1) {
#pragma omp parallel for schedule(dynamic, 16)
            for (int c = 1; c <= i_base; c++)
                for (int k = 1; k <= m0; k++) {
                    const double *uk = U + (size_t)(k - 1) * nn, *vk = V + (size_t)(k - 1) * nn;
                    double uc = uk[c - 1], vc = vk[c - 1];
                    double *col = &A_(1, c);
                    for (int j = 1; j <= c; j++) col[j - 1] -= uk[j - 1] * vc + vk[j - 1] * uc;
                }
        }
    }
    /* eigen_trd_final (trd_t8.F:188-219) */
    {
        double t = A_(1, 2);
        e[0] = 0.0; e[1] = -t; A_(1, 2) = 2.0 * t;
        d[0] = A_(1, 1); d[1] = A_(2, 2);
    }
    free(W); free(U); free(V); free(p); free(st); free(pt);
}

/* ------------------------------------------------------------------ */
/* eigen_prd: blocked reduction to penta-diagonal form, two columns per */
/* step (src/eigen_prd.F:341-580).  Outputs d(1:n), e1(i) = T(i-1,i),   */
/* e2(i) = T(i-2,i); reflector of column i (length i-2) left in a.      */
/*   compute_u  src/eigen_prd_t4x.F:114-353 (two passes of Cholesky-QR  */
/*              on the column pair, then two Householder reflectors)    */
/*   au         src/eigen_prd_t2.F:153-206 (two right-hand sides)       */
/*   compute_v  src/eigen_prd_t6_3.F:160-458 (panel corrections, 2x2    */
/*              coupling c, V = (AU)C - U M with M + M^T = C^T U^T A U C)*/
/*   local_2update src/eigen_prd_t5.F:69 ; panel load/store _t7.F:74,195 */
/*   init/final src/eigen_prd_t8.F:75,207                               */
/* ------------------------------------------------------------------ */
void ora_prd(int n, double *a, int lda, double *d, double *e1, double *e2, int m_in)
{
    if (n <= 0) return;
    for (int i = 0; i < n; i++) { d[i] = 0.0; e1[i] = 0.0; e2[i] = 0.0; }
    const int nrem = 2 + n % 2;                 /* MBAND + mod(n, MBAND) leading columns stay */
    size_t nn = (size_t)n;
    if (n > nrem) {
        int m = m_in < n ? m_in : n; m -= m % 2; if (m < 2) m = 2;
        int mm = ((n - nrem) - 1) / m + 1 + 1;
        double *W = (double *)calloc(nn * m, sizeof(double));
        double *U = (double *)calloc(nn * m, sizeof(double));
        double *V = (double *)calloc(nn * m, sizeof(double));
        double *P = (double *)calloc(nn * 2, sizeof(double));
        for (int i_block = mm; i_block >= 2; i_block--) {
            int i_base = (i_block - 2) * m + nrem;
            int m0 = m < n - i_base ? m : n - i_base;
            if (m0 < 1) continue;
            for (int k = 1; k <= m0; k++) {
                double *wk = W + (size_t)(k - 1) * nn;
                for (int j = 1; j <= n; j++) wk[j - 1] = (j <= i_base + k) ? A_(j, i_base + k) : 0.0;
            }
            memset(U, 0, nn * m * sizeof(double)); memset(V, 0, nn * m * sizeof(double));
            for (int k1 = m0; k1 >= 2; k1 -= 2) {
                const int i = i_base + k1, L = i - 2;
                double *x2 = W + (size_t)(k1 - 1) * nn;   /* column i   -> u_x(:,2) */
                double *x1 = W + (size_t)(k1 - 2) * nn;   /* column i-1 -> u_x(:,1) */
                /* ---- compute_u ------------------------------------------------------ */
                const double e1_i = x2[L];                /* A(i-1,i), u_t(8) */
                double ut4 = 0, ut5 = 0, ut6 = 0, ut7 = 0, r12 = 0, s11 = 0, s12 = 0, s22 = 0;
                int mask1 = 1, mask2 = 1;
                for (int itr = 1; itr <= 2; itr++) {
                    double u1 = 0.0, u2 = 0.0;
                    for (int j = 0; j < L; j++) { u2 = fmax(u2, fabs(x2[j])); u1 = fmax(u1, fabs(x1[j])); }
                    if (u1 == 0.0) u1 = 1.0;
                    if (u2 == 0.0) u2 = 1.0;
                    double t11 = 0, t12 = 0, t22 = 0;
                    for (int j = 0; j < L; j++) {
                        double t = x2[j] / u2, sj = x1[j] / u1;
                        t11 += t * t; t12 += sj * t; t22 += sj * sj;
                    }
                    t12 *= u2 * u1; t11 *= u2 * u2; t22 *= u1 * u1;
                    ut4 = (L >= 2) ? x1[L - 2] : 0.0; ut5 = x1[L - 1];
                    if (itr == 1) { ut6 = (L >= 2) ? x2[L - 2] : 0.0; ut7 = x2[L - 1]; }
                    if (t11 == 0.0) { mask2 = 0; s11 = 0.0; s12 = 0.0; s22 = t22; }
                    else { mask2 = 1; s11 = t11; s12 = t12 / t11; s22 = t22 - s12 * t12; }
                    mask1 = (s22 != 0.0);
                    if (mask2) {
                        if (mask1) {
                            for (int j = 0; j < L; j++) x1[j] -= s12 * x2[j];
                            ut4 -= s12 * ut6; ut5 -= s12 * ut7;
                        }
                    } else { for (int j = 0; j < L; j++) x2[j] = 0.0; ut6 = 0.0; ut7 = 0.0; }
                    if (!mask1) { for (int j = 0; j < L; j++) x1[j] = 0.0; ut4 = 0.0; ut5 = 0.0; }
                    r12 += s12;
                }
                const double rr1 = sqrt(s22 > 0.0 ? s22 : 0.0), rr2 = sqrt(s11);
                double bet1 = 1.0, bet2 = 1.0, sgm1 = 0.0, sgm2 = 0.0;
                if (mask2) {
                    sgm2 = -sign_(rr2, ut7);
                    x2[L - 1] -= sgm2; ut7 -= sgm2; bet2 = -ut7 * sgm2;
                    if (mask1) {
                        double sc = sgm2 * ut5 / bet2;
                        for (int j = 0; j < L - 1; j++) x1[j] += sc * x2[j];
                        ut4 += sc * ut6;
                    }
                }
                if (mask1) {
                    sgm1 = -sign_(rr1, ut4);
                    if (L >= 2) x1[L - 2] -= sgm1;
                    ut4 -= sgm1; bet1 = -ut4 * sgm1;
                }
                if (mask2) { x1[L - 1] = 0.0; x2[L] = 0.0; }
                e1[i - 2] = sgm2 * r12; e1[i - 1] = e1_i; e2[i - 2] = sgm1; e2[i - 1] = sgm2;
                const double c11 = 1.0 / bet1, c22 = 1.0 / bet2;
                /* ---- au: P = A(1:L,1:L) [x1 x2] on the panel-start matrix ------------ */
                double *p1 = P, *p2 = P + nn;
#pragma omp parallel for schedule(static)
                for (int r = 1; r <= L; r++) {
                    double s1 = 0.0, s2 = 0.0;
                    for (int c = 1; c <= L; c++) {
                        double arc = (r <= c) ? A_(r, c) : A_(c, r);
                        s1 += arc * x1[c - 1]; s2 += arc * x2[c - 1];
                    }
                    p1[r - 1] = s1; p2[r - 1] = s2;
                }
                /* ---- compute_v: corrections with the finished pairs of this panel ---- */
                for (int l = k1 + 1; l <= m0; l++) {
                    const double *ul = U + (size_t)(l - 1) * nn, *vl = V + (size_t)(l - 1) * nn;
                    double vu1 = 0, uu1 = 0, vu2 = 0, uu2 = 0;
                    for (int j = 0; j < L; j++) {
                        vu1 += vl[j] * x1[j]; uu1 += ul[j] * x1[j];
                        vu2 += vl[j] * x2[j]; uu2 += ul[j] * x2[j];
                    }
                    for (int j = 0; j < L; j++) {
                        p1[j] -= ul[j] * vu1 + vl[j] * uu1;
                        p2[j] -= ul[j] * vu2 + vl[j] * uu2;
                    }
                }
                double g11 = 0, g12a = 0, g12b = 0, g22 = 0, g21 = 0;
                for (int j = 0; j < L; j++) {
                    g11 += x1[j] * p1[j]; g12a += x1[j] * p2[j]; g12b += x2[j] * p1[j]; g22 += x2[j] * p2[j];
                    g21 += x2[j] * x1[j];
                }
                const double g12 = 0.5 * (g12a + g12b);
                const double c12 = -c22 * c11 * g21;
                double t11 = g11 * c11 + g12 * c12, t21 = g12 * c11 + g22 * c12, t12 = g12 * c22, t22 = g22 * c22;
                double q11 = c11 * t11 + c12 * t21, q21 = c22 * t21, q12 = c11 * t12 + c12 * t22, q22 = c22 * t22;
                const double m11 = 0.5 * q11, m12 = 0.5 * (q21 + q12), m22 = 0.5 * q22;
                double *uk1 = U + (size_t)(k1 - 2) * nn, *uk2 = U + (size_t)(k1 - 1) * nn;
                double *vk1 = V + (size_t)(k1 - 2) * nn, *vk2 = V + (size_t)(k1 - 1) * nn;
                for (int j = 0; j < L; j++) {
                    double y1 = p1[j] * c11 + p2[j] * c12, y2 = p2[j] * c22;
                    uk1[j] = x1[j]; uk2[j] = x2[j];
                    vk1[j] = y1 - x1[j] * m11 - x2[j] * m12;
                    vk2[j] = y2 - x2[j] * m22;
                }
                /* ---- local_2update: the other panel columns at once ------------------ */
                for (int c = 1; c <= k1 - 2; c++) {
                    int gc = i_base + c;
                    double *wc = W + (size_t)(c - 1) * nn;
                    for (int q = 0; q < 2; q++) {
                        const double *uq = q ? uk2 : uk1, *vq = q ? vk2 : vk1;
                        double uc = uq[gc - 1], vc = vq[gc - 1];
                        for (int j = 0; j < gc; j++) wc[j] -= uq[j] * vc + vq[j] * uc;
                    }
                }
            }
            /* panel store */
            for (int k = 1; k <= m0; k++) {
                int gc = i_base + k;
                const double *wk = W + (size_t)(k - 1) * nn;
                for (int j = 1; j <= gc; j++) A_(j, gc) = wk[j - 1];
            }
            /* rank-2k update of the trailing matrix (eigen_t1.F:250-306) */
#pragma omp parallel for schedule(dynamic, 16)
            for (int c = 1; c <= i_base; c++)
                for (int k = 1; k <= m0; k++) {
                    const double *uk = U + (size_t)(k - 1) * nn, *vk = V + (size_t)(k - 1) * nn;
                    double uc = uk[c - 1], vc = vk[c - 1];
                    double *col = &A_(1, c);
                    for (int j = 1; j <= c; j++) col[j - 1] -= uk[j - 1] * vc + vk[j - 1] * uc;
                }
        }
        free(W); free(U); free(V); free(P);
    }
    /* eigen_prd_final (prd_t8.F:207-315) */
    for (int i = (nrem < n ? nrem : n); i >= 2; i--) {
        e1[i - 1] = A_(i - 1, i); A_(i - 1, i) = 0.0;
        if (i - 2 >= 1) { e2[i - 1] = A_(i - 2, i); A_(i - 2, i) = 0.0; }
    }
    for (int j = 1; j <= n; j++) d[j - 1] = A_(j, j);
    e1[0] = 0.0; e2[0] = 0.0; if (n >= 2) e2[1] = 0.0;
}

/* ------------------------------------------------------------------ */
/* eigen_common_trbakwy: Z <- H_n ... H_2 Z                            */
/* a: output of ora_trd; e: off-diagonal from ora_trd (clobbered like  */
/* the reference's beta); z: n x nvec, ldz; m_b block; iblk = 1        */
/* ------------------------------------------------------------------ */
void ora_trbakwy(int n, int nvec, const double *a, int lda, double *z, int ldz, double *beta,
                 int m_in, int iblk)
{
    if (n <= iblk || nvec <= 0) return;
    int m = m_in < 256 ? m_in : 256; if (m < 1) m = 1; if (m > n) m = n;
    int nx = (n - (1 + iblk) + 1) % m + (1 + iblk) - 1; if (nx > n) nx = n;
    /* beta'(i) = u_L * g  (= -beta_i), inverted; zero -> 1  (trbakwy4.F:309-336) */
    for (int i = 1; i <= nx; i++) {
        int L = i - iblk;
        double b = (i >= 1 + iblk) ? A_(L, i) * beta[i - 1] : 0.0;
        beta[i - 1] = b;
    }
    for (int i = 1 + iblk; i <= nx; i++) beta[i - 1] = (beta[i - 1] == 0.0) ? 1.0 : 1.0 / beta[i - 1];
    /* head: one reflector at a time (trbakwy4.F:345-499) */
    for (int i = 1 + iblk; i <= nx; i++) {
        int L = i - iblk;
        const double *u = &A_(1, i);
#pragma omp parallel for schedule(static)
        for (int c = 0; c < nvec; c++) {
            double *zc = z + (size_t)c * ldz;
            double s = 0.0;
            for (int j = 0; j < L; j++) s += u[j] * zc[j];
            s *= beta[i - 1];
            for (int j = 0; j < L; j++) zc[j] += s * u[j];
        }
    }
    /* blocks of m reflectors in compact WY form (trbakwy4.F:538-602, trbakwy4_body.F) */
    double *Vb = (double *)malloc((size_t)n * m * sizeof(double));
    double *SM = (double *)malloc((size_t)m * m * sizeof(double));
    for (int i = nx + 1; i <= n; i += m) {
        int rows = i + m - 1 - iblk;
        for (int k = 0; k < m; k++) {
            int len = i + k - iblk;
            for (int j = 0; j < rows; j++) Vb[(size_t)k * rows + j] = (j < len) ? A_(j + 1, i + k) : 0.0;
        }
        /* SM = -V^T V (lower), diagonal halved, zero -> 1 (body:206-213,302-313) */
        for (int j = 0; j < m; j++)
            for (int k = 0; k <= j; k++) {
                double s = 0.0;
                for (int r = 0; r < rows; r++) s += Vb[(size_t)j * rows + r] * Vb[(size_t)k * rows + r];
                SM[(size_t)k * m + j] = -s; /* column-major SM(j,k), j>=k */
            }
        for (int j = 0; j < m; j++) {
            double t = SM[(size_t)j * m + j];
            SM[(size_t)j * m + j] = (t == 0.0) ? 1.0 : t * 0.5;
        }
        /* V <- V * SM^{-1}: solve X*SM = V, SM lower (dtrsm R,L,N,N body:687-688) */
        for (int k = m - 1; k >= 0; k--) {
            double dkk = SM[(size_t)k * m + k];
            for (int r = 0; r < rows; r++) {
                double s = Vb[(size_t)k * rows + r];
                for (int j = k + 1; j < m; j++) s -= Vb[(size_t)j * rows + r] * SM[(size_t)k * m + j];
                Vb[(size_t)k * rows + r] = s / dkk;
            }
        }
        /* per column of Z: ss = V0^T z ; z += X ss  (body:604-608,721-725) */
#pragma omp parallel
        {
            double *ss = (double *)malloc(m * sizeof(double));
#pragma omp for schedule(static)
            for (int c = 0; c < nvec; c++) {
                double *zc = z + (size_t)c * ldz;
                for (int k = 0; k < m; k++) {
                    int len = i + k - iblk;
                    const double *u = &A_(1, i + k);
                    double s = 0.0;
                    for (int j = 0; j < len; j++) s += u[j] * zc[j];
                    ss[k] = s;
                }
                for (int k = 0; k < m; k++) {
                    const double *x = Vb + (size_t)k * rows;
                    double s = ss[k];
                    for (int j = 0; j < rows; j++) zc[j] += x[j] * s;
                }
            }
            free(ss);
        }
    }
    free(Vb); free(SM);
}

/* ------------------------------------------------------------------ */
/* eigen_bisect (mode 'N'): eigenvalues of the tridiagonal (d,e),      */
/* e(i) couples i-1,i.  Sturm count + bisection (src/bisect.F:322-358) */
/* ------------------------------------------------------------------ */
static int sturm_count(int n, const double *d, const double *e, double x, double pivmin)
{
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) cnt++;
    for (int i = 1; i < n; i++) {
        q = d[i] - x - e[i] * e[i] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        if (q < 0.0) cnt++;
    }
    return cnt;
}

void ora_bisect(int n, const double *d, const double *e, double *w)
{
    double lo = d[0], hi = d[0], emax = 0.0;
    for (int i = 0; i < n; i++) {
        double r = (i > 0 ? fabs(e[i]) : 0.0) + (i + 1 < n ? fabs(e[i + 1]) : 0.0);
        if (d[i] - r < lo) lo = d[i] - r;
        if (d[i] + r > hi) hi = d[i] + r;
        if (i > 0 && e[i] * e[i] > emax) emax = e[i] * e[i];
    }
    double pivmin = DBL_MIN * (emax > 1.0 ? emax : 1.0);
    double span = hi - lo; lo -= 2.0 * DBL_EPSILON * n * (fabs(lo) + span) + 2 * pivmin;
    hi += 2.0 * DBL_EPSILON * n * (fabs(hi) + span) + 2 * pivmin;
#pragma omp parallel for schedule(dynamic, 16)
    for (int k = 0; k < n; k++) {
        double a0 = lo, b0 = hi;
        for (int it = 0; it < 200; it++) {
            double mid = 0.5 * (a0 + b0);
            if (mid <= a0 || mid >= b0) break;
            if (sturm_count(n, d, e, mid, pivmin) > k) b0 = mid; else a0 = mid;
        }
        w[k] = 0.5 * (a0 + b0);
    }
}

/* ------------------------------------------------------------------ */
/* test matrices: benchmark/mat_set.f                                  */
/* random family: counter-based, so every grid builds the same matrix  */
/* ------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
double ora_rand_ij(uint64_t seed, int i, int j, int n) /* U[0,1), i,j 1-based */
{
    uint64_t h = splitmix64(seed * 0x2545F4914F6CDD1Dull + (uint64_t)(i - 1) * (uint64_t)n + (uint64_t)(j - 1));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
/* element (j,i) of test matrix mtype in {0 Frank,1 Toeplitz,2 random R+R^T,3 Frank2} */
double ora_mat_elem(int mtype, int n, int j, int i, uint64_t seed)
{
    switch (mtype) {
    case 0: return (double)(i < j ? i : j);
    case 1: return i == j ? -7.2 : -3.0 / ((double)(i - j) * (double)(i - j));
    case 2: return ora_rand_ij(seed, i, j, n) + ora_rand_ij(seed, j, i, n);
    case 3: return (double)(n + 1 - (i > j ? i : j));
    default: return 0.0;
    }
}
/* local part of the 2D-cyclic matrix on rank (x_inod,y_inod) of an x_nnod x y_nnod grid */
void ora_mat_set_local(int mtype, int n, double *a, int lda, int x_nnod, int y_nnod, int x_inod,
                       int y_inod, uint64_t seed)
{
    int ie = ora_loop_end(n, y_nnod, y_inod), je = ora_loop_end(n, x_nnod, x_inod);
#pragma omp parallel for schedule(static)
    for (int i1 = 1; i1 <= ie; i1++) {
        int i = ora_translate_l2g(i1, y_nnod, y_inod);
        for (int j1 = 1; j1 <= je; j1++) {
            int j = ora_translate_l2g(j1, x_nnod, x_inod);
            a[(size_t)(i1 - 1) * lda + (j1 - 1)] = ora_mat_elem(mtype, n, j, i, seed);
        }
    }
}

/* w_set for the Frank family (mat_set.f:638-647), ascending */
void ora_w_frank(int n, double *w)
{
    const double PAI = 3.14159265358979323846;
    for (int i = 1; i <= n; i++) {
        int j = n - i;
        double theta = PAI * (2 * j + 1) / (2 * n + 1);
        w[i - 1] = 0.5 / (1.0 - cos(theta));
    }
}
