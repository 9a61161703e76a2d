// ee_common.cuh -- shared declarations for the B200 EigenExa hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <string>
#include <vector>
#include <algorithm>

namespace ee {

// ---------------------------------------------------------------------------------------
// error handling: the product fails loudly (no CPU fallback anywhere)
// ---------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
struct FatalError {};   // thrown by fatal(), caught at every C-ABI entry point (ee_capi.cu)
[[noreturn]] void fatal(const char *what, const char *file, int line);

#define EE_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t err__ = (call);                                                        \
        if (err__ != cudaSuccess) {                                                        \
            char buf__[512];                                                               \
            snprintf(buf__, sizeof buf__, "CUDA error %s: %s", #call, cudaGetErrorString(err__)); \
            ::ee::fatal(buf__, __FILE__, __LINE__);                                        \
        }                                                                                  \
    } while (0)

extern std::atomic<long long> g_launches;
#define EE_LAUNCHED() (::ee::g_launches.fetch_add(1, std::memory_order_relaxed))
#define EE_CHECK_LAUNCH()                                                                  \
    do {                                                                                   \
        EE_LAUNCHED();                                                                     \
        EE_CUDA(cudaGetLastError());                                                       \
    } while (0)

// ---------------------------------------------------------------------------------------
// process grid + cyclic index algebra (0-based inside the library).
// Reference (1-based): src/eigen_libs0.F:1816-2258.
// ---------------------------------------------------------------------------------------
struct Grid {
    int nnod = 1, inod = 0;  // world size / 0-based world rank
    int px = 1, py = 1;      // x_nnod (rows dealt over x), y_nnod (cols dealt over y)
    int x = 0, y = 0;        // 0-based coordinates
    char order = 'C';
};

// number of local indices l with global g = l*P + r < G
__host__ __device__ inline int cyc_count(int G, int P, int r) { return G > r ? (G - r + P - 1) / P : 0; }

// ---------------------------------------------------------------------------------------
// library state (singleton, like the reference's module variables)
// ---------------------------------------------------------------------------------------
struct Comm;  // NCCL plumbing, ee_comm.cu
struct Context {
    bool initialized = false;
    Grid g;
    int device = 0;
    cudaStream_t stream = nullptr;   // main stream: every kernel of the path
    cudaStream_t stream2 = nullptr;  // copies / side work
    cudaStream_t stream3 = nullptr;  // reflectors back to the caller's a (its own stream: stream2 carries the V prefetch)
    cudaEvent_t ev_side = nullptr;   // main stream -> side stream hand-over
    Comm *comm = nullptr;
    int errinfo = 0;
    int profiling = 0;  // 0 off; 1 async CUDA events around symv/syr2k launches; 2 sync per kernel class
    double timings[48] = {0};
    int sm_count = 148;
    int debug_maxcols = 0;              // > 0: eigen_trd stops after this many columns (profiling aid)
    std::vector<cudaEvent_t> ev_pool;   // reusable timing events (profiling level 1)
    std::vector<float> symv_trace;      // per-column symv milliseconds of the last trd (profiling >= 1)
};
Context &ctx();

// device scratch with simple grow-only caching across calls
void *dev_alloc(size_t bytes);
void dev_free(void *p);

// ---------------------------------------------------------------------------------------
// stage drivers (device pointers, local 2D-cyclic parts)
// ---------------------------------------------------------------------------------------
// eigen_scaling: returns sigma (NaN if non-finite input).  src/eigen_scaling.F:59-154
double scaling_dev(int n, double *a, int lda);
// eigen_trd: a (lda x ncl) in/out, d_out/e_out device arrays of length n (replicated)
void trd_dev(int n, double *a, int lda, double *d_out, double *e_out, int m_forward);
// eigen_common_trbakwy with z distributed 2D cyclic (ldz x nvl)
void trbak_dev(int n, int nvec, const double *a, int lda, double *z, int ldz, const double *e, int m_backward, int iblk = 1);
// the next trbak_dev call also delivers Z to this host array (ld ldz_host), chunk by chunk on the side stream;
// the caller synchronises stream2 before it reads z_host
void trbak_set_host_output(double *z_host, int ldz_host);
// eigen_prd (penta-diagonal reduction): e1(i) = T(i-1,i), e2(i) = T(i-2,i)
void prd_dev(int n, double *a, int lda, double *d_out, double *e1_out, double *e2_out, int m_forward);
// tridiagonal divide and conquer; z (ldz x nvl) receives the local cyclic part
int dc_dev(int n, int nvec, const double *d, const double *e, double *w, double *z, int ldz);
// same for the penta-diagonal matrix (d, e, e2) of eigen_prd (eigen_dcx, src/dcx.F:75)
int dc_band_dev(int n, int nvec, const double *d, const double *e, const double *e2, double *w, double *z, int ldz);
// bisection
void bisect_dev(int n, const double *d, const double *e, double *w);
void bisect2_dev(int n, const double *d, const double *e1, const double *e2, double *w);   // penta-diagonal
// mat_set / ev_test
void mat_set_dev(int n, double *a, int lda, int mtype, uint64_t seed);
void ev_test_dev(int n, int nvec, const double *a, int lda, const double *w, const double *z, int ldz, double *out);

// ---------------------------------------------------------------------------------------
// FP64 tensor-core GEMM (DMMA m8n8k4), ee_gemm.cu.  Column-major everywhere.
//   C(MxN) = alpha * opA(A) * opB(B) + beta * C
//   transA = 'N': A is M x K (lda);  'T': A is K x M (lda), opA(A) = A^T
//   transB = 'N': B is K x N (ldb);  'T': B is N x K (ldb), opB(B) = B^T
// tri_mode: 0 full; 1 only tiles that touch the upper "staircase" of a cyclic local
// matrix (row g = jl*px+x, col g = il*py+y, keep g_row <= g_col) are updated.
// ---------------------------------------------------------------------------------------
struct TriSpec { int mode = 0, px = 1, py = 1, x = 0, y = 0; };
void dgemm(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha, const double *A,
           int lda, const double *B, int ldb, double beta, double *C, int ldc, TriSpec tri = TriSpec());

// split-K / 64-bit leading dimension form: blockIdx.z = 0..ksplit-1 handles a K slice and
// writes its partial product to C + z*c_stride
void dgemm_ex(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha, const double *A,
              long long lda, const double *B, long long ldb, double beta, double *C, long long ldc, int ksplit,
              long long c_stride, TriSpec tri = TriSpec());
int trd_lda_pad(int nrl);
int trd_ncl_pad(int ncl);
void scale_vec_dev(double *v, int n, double s, cudaStream_t st);
void mat_set_host(int n, double *a, int lda, int mtype, uint64_t seed, const Grid &g);

}  // namespace ee
