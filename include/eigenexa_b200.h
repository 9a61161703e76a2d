/*
 * eigenexa_b200.h -- C ABI of the B200-native EigenExa hot path.
 *
 * Drop-in boundary: the entry points below are what the reference's C binding
 * (C/EigenExa.h:12-46, C/EigenExa.c) and its Fortran shims (C/EigenExa.fh:10-19,
 * C/eigen_exa_interfaces.F90) expose for eigen_init / eigen_free / eigen_get_matdims /
 * eigen_s / eigen_sx and the query helpers.  Argument meaning, defaults and the
 * "silent return" error convention follow src/eigen_libs.F:70-216 and
 * src/eigen_s.F:81-133.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Host arrays a, z use the reference's 2D cyclic layout (block size 1,
 * src/eigen_libs0.F:1986-2166): local a(j_loc,i_loc) = A(global row
 * (j_loc-1)*x_nnod+x_inod, global col (i_loc-1)*y_nnod+y_inod), column-major with
 * leading dimension lda >= local row count.  Only the upper triangle of A is read;
 * a is destroyed (a(1:3,1) = flop count, seconds, -1; src/eigen_s.F:284-295).
 *
 * The reference takes an MPI communicator.  This build has no MPI: ranks are one
 * process per GPU and the collectives are NCCL, so eigen_init takes a small POD
 * describing the rank and the NCCL bootstrap id instead (INTEGRATION.md shows the
 * MPI shim a maintainer adds: MPI_Comm_rank/size + MPI_Bcast of the id).
 */
#ifndef EIGENEXA_B200_H
#define EIGENEXA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EIGENEXA_B200_UNIQUE_ID_BYTES 128

/* replaces the MPI_Comm argument of eigen_init (C/EigenExa.h:12) */
typedef struct eigenexa_b200_comm {
    int rank;             /* 0-based rank in the job                                */
    int nranks;           /* number of ranks (= GPUs)                               */
    int device;           /* CUDA device ordinal for this rank, -1: keep current    */
    int reserved;
    unsigned char unique_id[EIGENEXA_B200_UNIQUE_ID_BYTES]; /* ncclUniqueId made by
                           eigenexa_b200_get_unique_id on rank 0 and broadcast by the
                           host (ignored when nranks == 1)                           */
} eigenexa_b200_comm_t;

/* rank 0 fills id[128]; returns 0 on success */
int eigenexa_b200_get_unique_id(unsigned char *id);

/* ---- reference API (C/EigenExa.h) ------------------------------------------------ */
/* eigen_init(comm, order): src/eigen_libs.F:70-104; order 'C' (default) or 'R'.
 * comm == NULL means a single-rank job on the current device.                       */
void eigen_init(const eigenexa_b200_comm_t *comm, const char *order);
void eigen_free(void);                                       /* src/eigen_libs.F:204-216 */

/* eigen_s / eigen_sx: src/eigen_libs.F:150-202, src/eigen_sx.F:30.  mode "A" (values +
 * vectors), "N" (values only), "X" (values refined by bisection + vectors).          */
void eigen_s(int n, int nvec, double *a, int lda, double *w, double *z, int ldz,
             int m_forward, int m_backward, const char *mode);
void eigen_sx(int n, int nvec, double *a, int lda, double *w, double *z, int ldz,
              int m_forward, int m_backward, const char *mode);

void eigen_get_version(int *version, char *date, char *vcode); /* eigen_libs0.F:175-188 */
void eigen_get_procs(int *nnod, int *x_nnod, int *y_nnod);     /* eigen_libs0.F:1551-1600 */
void eigen_get_id(int *inod, int *x_inod, int *y_inod);        /* 1-based ids            */
void eigen_get_matdims(int n, int *nx, int *ny, int m_forward, int m_backward,
                       const char *mode);                      /* eigen_libs.F:106-148   */
void eigen_get_errinfo(int *info);                             /* eigen_libs0.F:1689-1698 */
int64_t eigen_memory_internal(int n, int lda, int ldz, int m1, int m0); /* device bytes */

/* index helpers, explicit (nnod, inod) forms: eigen_libs0.F:1816-2258 */
int eigen_loop_start(int istart, int nnod, int inod);
int eigen_loop_end(int iend, int nnod, int inod);
int eigen_translate_l2g(int ictr, int nnod, int inod);
int eigen_translate_g2l(int ictr, int nnod, int inod);
int eigen_owner_node(int ictr, int nnod, int inod);
int eigen_owner_index(int ictr, int nnod, int inod);

/* ---- Fortran-77 style symbols the reference's C layer binds (C/EigenExa.fh:10-19):
 * all arguments by reference, hidden string lengths ignored.                         */
void eigen_libs_eigen_init_(const eigenexa_b200_comm_t *comm, const char *order);
void eigen_libs_eigen_free_(void);
void eigen_libs_eigen_s_(int *n, int *nvec, double *a, int *lda, double *w, double *z,
                         int *ldz, int *m_forward, int *m_backward, const char *mode);
void eigen_libs_eigen_sx_(int *n, int *nvec, double *a, int *lda, double *w, double *z,
                          int *ldz, int *m_forward, int *m_backward, const char *mode);
void eigen_libs_eigen_get_matdims_(int *n, int *nx, int *ny, int *m_forward,
                                   int *m_backward, const char *mode);
void eigen_libs0_eigen_get_version_(int *version, char *date, char *vcode);
void eigen_libs0_eigen_get_procs_(int *nnod, int *x_nnod, int *y_nnod);
void eigen_libs0_eigen_get_id_(int *inod, int *x_inod, int *y_inod);
void eigen_libs0_eigen_get_errinfo_(int *info);

/* ---- the rest of the reference's C-visible surface (C/eigen_exa_interfaces.h:3-33, C/EigenExa.h:40) -- */
void eigen_show_version(void);                                     /* eigen_libs0.F:207-236  */
void eigen_loop_info(int istart, int iend, int *lstart, int *lend,
                     int nnod, int inod);                          /* eigen_libs0.F:1744-1760 */
int eigen_convert_id_xy2w(int xinod, int yinod);                   /* eigen_libs0.F:2316-2334 (values as the reference computes them) */
void eigen_convert_id_w2xy(int inod, int *xinod, int *yinod);      /* eigen_libs0.F:2336-2356 */
/* eigen_get_comm(comm, x_comm, y_comm) (C/EigenExa.h:40): no MPI here, so the three communicators are
 * described by the POD eigen_init takes: rank/nranks of this process in the world, in its x group (ranks
 * sharing y_inod) and in its y group; rank = -1 before eigen_init.  unique_id is left zero.            */
void eigen_get_comm(eigenexa_b200_comm_t *comm, eigenexa_b200_comm_t *x_comm,
                    eigenexa_b200_comm_t *y_comm);
/* Fortran face of the same (integer handles in place of MPI_Fint: 0 world, 1 x, 2 y; -1 = not initialised) */
void eigen_libs0_eigen_get_comm_(int *comm, int *x_comm, int *y_comm);
void eigen_libs0_eigen_show_version_(void);
int eigen_libs0_eigen_memory_internal_(int *n, int *lda, int *ldz, int *m1, int *m0); /* bytes, clamped to INT_MAX */
int eigen_libs0_eigen_loop_start_(int *istart, int *nnod, int *inod);
int eigen_libs0_eigen_loop_end_(int *iend, int *nnod, int *inod);
void eigen_libs0_eigen_loop_info_(int *istart, int *iend, int *lstart, int *lend, int *nnod, int *inod);
int eigen_libs0_eigen_translate_l2g_(int *ictr, int *nnod, int *inod);
int eigen_libs0_eigen_translate_g2l_(int *ictr, int *nnod, int *inod);
int eigen_libs0_eigen_owner_node_(int *ictr, int *nnod, int *inod);
int eigen_libs0_eigen_owner_index_(int *ictr, int *nnod, int *inod);
int eigen_libs0_eigen_convert_id_xy2w_(int *xinod, int *yinod);
void eigen_libs0_eigen_convert_id_w2xy_(int *inod, int *xinod, int *yinod);
int eigen_blacs_eigen_get_blacs_context_(void);   /* BLACS glue is out of scope: always -1 */
/* not provided: eigen_h / eigen_libs_eigen_h_ (Hermitian solver, out of scope) */

/* ---- stage-level entry points (the reference's public module procedures) ----------
 * eigen_trd(n,a,lda,d,e,m)                 src/eigen_trd.F:82
 * eigen_common_trbakwy(n,nvec,a,lda,z,ldz,e,m,iblk)  src/trbakwy4.F:77
 * Host arrays in the 2D cyclic layout; d, e, of length n are replicated outputs.
 * Return 0 on success, nonzero on error (not initialised, bad arguments).            */
int eigenexa_b200_trd(int n, double *a, int lda, double *d, double *e, int m_forward);
int eigenexa_b200_trbakwy(int n, int nvec, const double *a, int lda, double *z, int ldz,
                          const double *e, int m_backward);
/* tridiagonal eigen-decomposition stage (eigen_dc2, src/dc2.F:78): d,e replicated in,
 * w ascending out, z = local part of the eigenvector matrix in the 2D cyclic layout.  */
int eigenexa_b200_dc(int n, int nvec, const double *d, const double *e, double *w,
                     double *z, int ldz);
/* eigenvalues only by Sturm bisection (eigen_bisect, src/bisect.F:67)                 */
int eigenexa_b200_bisect(int n, const double *d, const double *e, double *w);

/* ---- penta-diagonal path of eigen_sx ------------------------------------------------
 * eigen_prd(n,a,lda,d,e,ne,m)   src/eigen_prd.F:80    e is (ne x 2): e(:,1) = T(i-1,i),
 *                                                      e(:,2) = T(i-2,i); reflectors of length i-2 in a
 * eigen_dcx(n,nvec,d,e,ne,z,ldz,info,ret) src/dcx.F:75 band divide & conquer
 * eigen_bisect2(d,e,f,w,n,mode)  src/bisect2.F:71
 * eigen_common_trbakwy(...,nb)   src/trbakwy4.F:77     nb = 2 for the reflectors of eigen_prd */
int eigenexa_b200_prd(int n, double *a, int lda, double *d, double *e, int ne, int m_forward);
int eigenexa_b200_dcx(int n, int nvec, const double *d, const double *e, int ne, double *w,
                      double *z, int ldz);
int eigenexa_b200_bisect2(int n, const double *d, const double *e, int ne, double *w);
int eigenexa_b200_trbakwy_nb(int n, int nvec, const double *a, int lda, double *z, int ldz,
                             const double *e, int m_backward, int nb);

/* ---- device-resident variant: a_dev / z_dev / w_dev are DEVICE pointers on this
 * rank's GPU (same layout).  Used to time the path with inputs already in HBM.       */
int eigenexa_b200_eigen_s_dev(int n, int nvec, double *a_dev, int lda, double *w_dev,
                              double *z_dev, int ldz, int m_forward, int m_backward,
                              const char *mode);
int eigenexa_b200_eigen_sx_dev(int n, int nvec, double *a_dev, int lda, double *w_dev,
                               double *z_dev, int ldz, int m_forward, int m_backward,
                               const char *mode);

/* ---- benchmark/mat_set.f generators, written straight into device or host memory:
 * mtype 0 Frank, 1 Toeplitz, 2 random R+R^T (counter-based, seed), 3 Frank-2.        */
int eigenexa_b200_mat_set_dev(int n, double *a_dev, int lda, int mtype, uint64_t seed);
int eigenexa_b200_mat_set_host(int n, double *a, int lda, int mtype, uint64_t seed);
/* benchmark/ev_test.f on device: out[0]=|AZ-ZW|_F/(N eps |A|_F), out[1]=|Z^TZ-I|_F/(N eps),
 * out[2]=|A|_F; out must hold 4 doubles (single rank; a_dev holds the full symmetric matrix, upper triangle read)           */
int eigenexa_b200_ev_test_dev(int n, int nvec, const double *a_dev, int lda,
                              const double *w_dev, const double *z_dev, int ldz, double *out);

/* the FP64 tensor-core (DMMA) GEMM the path is built on, BLAS dgemm argument order
 * (replaces the dgemm calls at src/eigen_t1.F:285-295, src/trbakwy4_body.F:573,604,721);
 * device pointers, column-major, asynchronous on the library stream.                  */
int eigenexa_b200_dgemm_dev(char transa, char transb, int m, int n, int k, double alpha,
                            const double *a_dev, int lda, const double *b_dev, int ldb,
                            double beta, double *c_dev, int ldc);
/* same, restricted to the tiles that reach the upper "staircase" of a 2D-cyclic local matrix
 * (the form eigen_common_2update uses, src/eigen_t1.F:250-306); tiles strictly below it are
 * left untouched, tiles on it are updated in full.                                       */
int eigenexa_b200_dgemm_tri_dev(char transa, char transb, int m, int n, int k, double alpha,
                                const double *a_dev, int lda, const double *b_dev, int ldb,
                                double beta, double *c_dev, int ldc, int px, int py, int x, int y);
void eigenexa_b200_sync(void);      /* wait for the library stream                     */
void *eigenexa_b200_stream(void);   /* cudaStream_t of the library (for event timing)  */

/* ---- instrumentation ------------------------------------------------------------- */
/* number of kernels this library launched since the last reset (bench gpu_launches)  */
int64_t eigenexa_b200_launch_count(int reset);
/* stage timings of the last eigen_s call on this rank, seconds (CUDA events), up to 48 slots:
 * t[0]=h2d t[1]=scaling+trd t[2]=tridiagonal solver t[3]=back-transform t[4]=d2h (w only: a and z stream out earlier)
 * with profiling on:  t[5]=panel kernels of eigen_trd (or symv launches on the multi-launch paths)  t[6]=rank-2k GEMMs
 * t[13]=merge GEMM flops of the D&C  t[17..21]=D&C breakdown
 * persistent panel kernel, in-kernel %globaltimer of CTA 0:  t[15]=SYMV phase  t[16]=p phase  t[31]=v phase;
 * on a grid, inside the p phase:  t[32]=partial sums  t[33]=grid barrier  t[34]=corrections + flag round  t[35]=peer sum */
void eigenexa_b200_last_timings(double *t, int nt);
void eigenexa_b200_set_profiling(int level); /* 0 off, 1 async events (symv, syr2k), 2 debug */
/* per-launch symv_kernel milliseconds of the last eigen_trd (first entry = column n);
 * needs profiling >= 1; returns the number of launches recorded                        */
int eigenexa_b200_symv_trace(float *out, int cap);
/* profiling aid: eigen_trd stops after ncols columns (results are then meaningless); 0 = off */
void eigenexa_b200_set_debug_maxcols(int ncols);
const char *eigenexa_b200_last_error(void);
/* Error convention on device / allocation / NCCL failures: the reference aborts the job (eigen_abort -> MPI_Abort,
 * src/eigen_devel.F:148-164).  This library unwinds to the entry point instead: eigen_s / eigen_sx fill w with NaN,
 * the int-returning entry points return 99, eigen_get_errinfo reports -1 and eigenexa_b200_last_error the message.
 * EIGENEXA_B200_ABORT_ON_ERROR=1 in the environment restores abort().  eigenexa_b200_debug_raise exercises that
 * path (test hook): it returns 99.                                                                                 */
int eigenexa_b200_debug_raise(void);

#ifdef __cplusplus
}
#endif
#endif /* EIGENEXA_B200_H */
