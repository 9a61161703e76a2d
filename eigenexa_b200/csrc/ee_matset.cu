// ee_matset.cu -- benchmark/mat_set.f generators and benchmark/ev_test.f checks on device.
//
// mat_set families (benchmark/mat_set.f:117-203): 0 Frank min(i,j), 1 Toeplitz, 2 random
// R + R^T with R ~ U[0,1), 3 Frank-2 n+1-max(i,j).  The reference seeds Fortran's
// random_number with the rank id (mat_set.f:167-179), which is neither portable across
// compilers nor across grids; here R(i,j) is a counter-based hash of (seed, i, j) so every
// grid (and the CPU oracle) builds the same global matrix.
// ev_test (benchmark/ev_test.f:113-205): |AZ-ZW|_F/(N eps |A|_F) and |Z^T Z-I|_F/(N eps).
#include "ee_common.cuh"
#include "ee_comm.h"
#include <thread>
#include <vector>

namespace ee {

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// U[0,1); i, j 1-based global indices
__host__ __device__ inline double rand_ij(uint64_t seed, int i, int j, int n)
{
    uint64_t h = splitmix64(seed * 0x2545F4914F6CDD1Dull + (uint64_t)(i - 1) * (uint64_t)n + (uint64_t)(j - 1));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
__host__ __device__ inline double mat_elem(int mtype, int n, int j, int i, uint64_t seed)
{
    switch (mtype) {
    case 0: return (double)(i < j ? i : j);
    case 1: return i == j ? -7.2 : -3.0 / ((double)(i - j) * (double)(i - j));
    case 2: return rand_ij(seed, i, j, n) + rand_ij(seed, j, i, n);
    case 3: return (double)(n + 1 - (i > j ? i : j));
    default: return 0.0;
    }
}

__global__ void mat_set_kernel(int n, double *a, int lda, int mtype, uint64_t seed, int px, int py, int x, int y, int nrl,
                               int ncl)
{
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= nrl) return;
    for (int il = blockIdx.y; il < ncl; il += gridDim.y) {     // gridDim.y is capped at 32768 by the callers
        const int gj = jl * px + x + 1, gi = il * py + y + 1;
        a[(size_t)il * lda + jl] = mat_elem(mtype, n, gj, gi, seed);
    }
}

__global__ void symmetrize_kernel(int n, const double *a, int lda, double *full, int ldf)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = blockIdx.y; i < n; i += gridDim.y)
        full[(size_t)i * ldf + j] = (j <= i) ? a[(size_t)i * lda + j] : a[(size_t)j * lda + i];
}

// r(:, c) -= w[c] * z(:, c)
__global__ void sub_zw_kernel(int n, int nv, double *r, int ldr, const double *z, int ldz, const double *w)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int c = blockIdx.y; c < nv; c += gridDim.y) r[(size_t)c * ldr + j] -= w[c] * z[(size_t)c * ldz + j];
}
__global__ void sub_eye_kernel(int nv, double *g, int ldg)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nv) g[(size_t)c * ldg + c] -= 1.0;
}
__global__ void sumsq_kernel(int rows, int cols, const double *a, int lda, double *out)
{
    double s = 0.0;
    for (int c = blockIdx.y; c < cols; c += gridDim.y)
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < rows; j += gridDim.x * blockDim.x) {
            double t = a[(size_t)c * lda + j];
            s = fma(t, t, s);
        }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

double fro2(cudaStream_t st, int rows, int cols, const double *a, int lda, double *scratch)
{
    EE_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
    dim3 grid(std::min(64, (rows + 255) / 256), std::min(cols, 2048));
    sumsq_kernel<<<grid, 256, 0, st>>>(rows, cols, a, lda, scratch);
    EE_CHECK_LAUNCH();
    double h;
    EE_CUDA(cudaMemcpyAsync(&h, scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
    EE_CUDA(cudaStreamSynchronize(st));
    return h;
}

}  // namespace

void mat_set_dev(int n, double *a, int lda, int mtype, uint64_t seed)
{
    Context &c = ctx();
    const Grid &g = c.g;
    const int nrl = cyc_count(n, g.px, g.x), ncl = cyc_count(n, g.py, g.y);
    if (nrl <= 0 || ncl <= 0) return;
    dim3 grid((nrl + 255) / 256, std::min(ncl, 32768));
    mat_set_kernel<<<grid, 256, 0, c.stream>>>(n, a, lda, mtype, seed, g.px, g.py, g.x, g.y, nrl, ncl);
    EE_CHECK_LAUNCH();
}

void mat_set_host(int n, double *a, int lda, int mtype, uint64_t seed, const Grid &g)
{
    const int nrl = cyc_count(n, g.px, g.x), ncl = cyc_count(n, g.py, g.y);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 64) nt = 64;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=]() {
            for (int il = (int)t; il < ncl; il += (int)nt) {
                const int gi = il * g.py + g.y + 1;
                double *col = a + (size_t)il * lda;
                for (int jl = 0; jl < nrl; jl++) col[jl] = mat_elem(mtype, n, jl * g.px + g.x + 1, gi, seed);
            }
        });
    for (auto &t : th) t.join();
}

// single-rank check (the reference redistributes to block-cyclic and calls PDGEMM; one GPU
// holds the N = 50000 problem, so the check runs where the data already is)
void ev_test_dev(int n, int nvec, const double *a, int lda, const double *w, const double *z, int ldz, double *out)
{
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const double eps = 2.220446049250313e-16;
    const int ldf = (n + 15) & ~15;
    double *full = (double *)dev_alloc((size_t)ldf * n * sizeof(double));
    double *r = (double *)dev_alloc((size_t)ldf * nvec * sizeof(double));
    double *scr = (double *)dev_alloc(64);
    dim3 grid((n + 255) / 256, std::min(n, 32768));
    symmetrize_kernel<<<grid, 256, 0, st>>>(n, a, lda, full, ldf);
    EE_CHECK_LAUNCH();
    const double anorm = sqrt(fro2(st, n, n, full, ldf, scr));
    dgemm(st, 'N', 'N', n, nvec, n, 1.0, full, ldf, z, ldz, 0.0, r, ldf);
    dim3 grid2((n + 255) / 256, std::min(nvec, 32768));
    sub_zw_kernel<<<grid2, 256, 0, st>>>(n, nvec, r, ldf, z, ldz, w);
    EE_CHECK_LAUNCH();
    const double err1 = sqrt(fro2(st, n, nvec, r, ldf, scr));
    // Z^T Z - I  (reuse "full" as nvec x nvec)
    dgemm(st, 'T', 'N', nvec, nvec, n, 1.0, z, ldz, z, ldz, 0.0, full, ldf);
    sub_eye_kernel<<<(nvec + 255) / 256, 256, 0, st>>>(nvec, full, ldf);
    EE_CHECK_LAUNCH();
    const double err2 = sqrt(fro2(st, nvec, nvec, full, ldf, scr));
    out[0] = err1 / ((double)n * eps * anorm);
    out[1] = err2 / ((double)n * eps);
    out[2] = anorm;
    dev_free(full); dev_free(r); dev_free(scr);
}

}  // namespace ee
