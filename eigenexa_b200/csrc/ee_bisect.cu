// ee_bisect.cu -- eigenvalues of the symmetric tridiagonal (d, e) by Sturm-count bisection.
//
// Replaces eigen_bisect (src/bisect.F:67-318, sturm :322-358) used by mode 'N' and by the
// mode 'X' refinement.  e(i) couples rows i-1 and i (e(0) unused), as produced by eigen_trd.
// One thread per eigenvalue; every thread walks the same (d, e^2) stream so the loads are
// warp broadcasts served from L1/L2.  Interval = Gershgorin bounds widened like
// bisect.F:159-176; stop when the midpoint no longer moves (bisect.F:272-283).
#include "ee_common.cuh"

namespace ee {

namespace {

__global__ void bisect_prep_kernel(int n, const double *d, const double *e, double *e2, double *bounds)
{
    // single CTA: Gershgorin interval, max|e|, pivmin
    __shared__ double s_lo[256], s_hi[256], s_em[256];
    double lo = d[0], hi = d[0], em = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double el = (i > 0) ? fabs(e[i]) : 0.0;
        double er = (i + 1 < n) ? fabs(e[i + 1]) : 0.0;
        lo = fmin(lo, d[i] - el - er);
        hi = fmax(hi, d[i] + el + er);
        em = fmax(em, el);
        e2[i] = (i > 0) ? e[i] * e[i] : 0.0;
    }
    s_lo[threadIdx.x] = lo; s_hi[threadIdx.x] = hi; s_em[threadIdx.x] = em;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < blockDim.x; i++) { lo = fmin(lo, s_lo[i]); hi = fmax(hi, s_hi[i]); em = fmax(em, s_em[i]); }
        const double eps = 2.220446049250313e-16;
        double epsilon = eps * em;
        double x = (fabs(lo) + fabs(hi)) * eps;
        bounds[0] = (lo - x) - epsilon;
        bounds[1] = (hi + x) + epsilon;
        double pivmin = 2.2250738585072014e-308 * fmax(1.0, em * em);
        bounds[2] = pivmin;
    }
}

// number of eigenvalues < x
__device__ __forceinline__ int sturm_count(int n, const double *__restrict__ d, const double *__restrict__ e2, double x,
                                           double pivmin)
{
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += (q < 0.0);
    for (int i = 1; i < n; i++) {
        q = d[i] - x - e2[i] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += (q < 0.0);
    }
    return cnt;
}

__global__ void bisect_kernel(int n, const double *__restrict__ d, const double *__restrict__ e2,
                              const double *__restrict__ bounds, double *w)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double lb = bounds[0], ub = bounds[1];
    const double pivmin = bounds[2];
    double x = lb;
    for (int it = 0; it < 2048; it++) {
        double t = x;
        x = 0.5 * (lb + ub);
        if (x == t || x <= lb || x >= ub) break;
        int s = sturm_count(n, d, e2, x, pivmin);
        if (s <= k) lb = x; else ub = x;   // eigenvalue k (0-based) has exactly k eigenvalues below it
    }
    w[k] = x;
}

// w is produced in ascending order by construction; enforce monotonicity exactly like the
// reference's final sort (bisect.F:306-310) with a cheap odd-even clean-up pass
__global__ void monotone_fix_kernel(int n, double *w)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int i = 1; i < n; i++)
        if (w[i] < w[i - 1]) w[i] = w[i - 1];
}

// ---- penta-diagonal: inertia of T - x I from a banded block L D L^T -------------------------------
// eigen_bisect2 / sturm2_LDLT (src/bisect2.F:393-678) counts the negative pivots of an L D L^T factorisation with a
// "diagonal-neighbour" pivoting strategy, because the unpivoted band recurrence is not backward stable (a tiny
// pivot makes the multipliers blow up and the next pivot absorbs the sub-diagonal entry).  Same safeguard here,
// written as Bunch's rule for band matrices without interchanges: the reduced matrix keeps a 2x2 window
//     [ p s ]     p = R(i,i), s = R(i,i+1), r = R(i+1,i+1)      (everything right / below is still original)
//     [ s r ]
// 1x1 pivot p when sigma |p| >= alpha s^2 (alpha = (sqrt5-1)/2, sigma = size of T - xI), otherwise the 2x2 block is
// eliminated at once (its inertia: det < 0 -> one negative pivot, det > 0 -> two or none by the sign of p + r).
__global__ void bisect2_prep_kernel(int n, const double *d, const double *e1, const double *e2, double *bounds)
{
    __shared__ double s_lo[256], s_hi[256], s_em[256];
    double lo = d[0], hi = d[0], em = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double r = 0.0;
        if (i > 0) r += fabs(e1[i]);
        if (i > 1) r += fabs(e2[i]);
        if (i + 1 < n) r += fabs(e1[i + 1]);
        if (i + 2 < n) r += fabs(e2[i + 2]);
        lo = fmin(lo, d[i] - r); hi = fmax(hi, d[i] + r);
        em = fmax(em, r);
    }
    s_lo[threadIdx.x] = lo; s_hi[threadIdx.x] = hi; s_em[threadIdx.x] = em;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < blockDim.x; i++) { lo = fmin(lo, s_lo[i]); hi = fmax(hi, s_hi[i]); em = fmax(em, s_em[i]); }
        const double eps = 2.220446049250313e-16;
        double x = (fabs(lo) + fabs(hi)) * eps + eps * em;
        bounds[0] = lo - x; bounds[1] = hi + x;
        bounds[2] = 2.2250738585072014e-308 * fmax(1.0, em * em);
        bounds[3] = fmax((hi - lo) + em, 2.2250738585072014e-308);   // sigma: bound of |T - xI| entries for x in [lo, hi]
    }
}

__device__ __forceinline__ int sturm2_count(int n, const double *__restrict__ d, const double *__restrict__ e1,
                                            const double *__restrict__ e2, double x, double pivmin, double sigma)
{
    const double alpha = 0.6180339887498949;
    int cnt = 0, i = 0;
    double p = d[0] - x;
    double s = (n > 1) ? e1[1] : 0.0;
    double r = (n > 1) ? d[1] - x : 0.0;
    while (i < n) {
        const double f = (i + 2 < n) ? e2[i + 2] : 0.0;      // R(i+2, i)
        const double g = (i + 2 < n) ? e1[i + 2] : 0.0;      // R(i+2, i+1)
        if (i == n - 1 || sigma * fabs(p) >= alpha * s * s) {
            if (!(fabs(p) >= pivmin)) p = -pivmin;            // also catches a non-finite window
            cnt += (p < 0.0);
            const double sp = s / p, fp = f / p;
            const double pn = r - s * sp;
            const double sn = g - s * fp;
            const double rn = (i + 2 < n) ? (d[i + 2] - x) - f * fp : 0.0;
            p = pn; s = sn; r = rn;
            i += 1;
        } else {
            double det = p * r - s * s;
            if (!(fabs(det) >= pivmin)) det = -pivmin;
            cnt += (det < 0.0) ? 1 : ((p + r < 0.0) ? 2 : 0);
            const double h = (i + 3 < n) ? e2[i + 3] : 0.0;  // R(i+3, i+1)
            const double upu = (r * f * f - 2.0 * s * f * g + p * g * g) / det;
            const double upv = h * (p * g - s * f) / det;
            const double vpv = p * h * h / det;
            const double pn = (i + 2 < n) ? (d[i + 2] - x) - upu : 0.0;
            const double sn = (i + 3 < n) ? e1[i + 3] - upv : 0.0;
            const double rn = (i + 3 < n) ? (d[i + 3] - x) - vpv : 0.0;
            p = pn; s = sn; r = rn;
            i += 2;
        }
    }
    return cnt;
}

__global__ void bisect2_kernel(int n, const double *__restrict__ d, const double *__restrict__ e1,
                               const double *__restrict__ e2, const double *__restrict__ bounds, double *w)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double lb = bounds[0], ub = bounds[1];
    const double pivmin = bounds[2], sigma = bounds[3];
    double x = lb;
    for (int it = 0; it < 2048; it++) {
        double t = x;
        x = 0.5 * (lb + ub);
        if (x == t || x <= lb || x >= ub) break;
        int s = sturm2_count(n, d, e1, e2, x, pivmin, sigma);
        if (s <= k) lb = x; else ub = x;
    }
    w[k] = x;
}

}  // namespace

void bisect2_dev(int n, const double *d, const double *e1, const double *e2, double *w)
{
    Context &c = ctx();
    cudaStream_t st = c.stream;
    if (n == 1) { EE_CUDA(cudaMemcpyAsync(w, d, sizeof(double), cudaMemcpyDeviceToDevice, st)); return; }
    double *bounds = (double *)dev_alloc(sizeof(double) * 4);
    bisect2_prep_kernel<<<1, 256, 0, st>>>(n, d, e1, e2, bounds);
    EE_CHECK_LAUNCH();
    bisect2_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, d, e1, e2, bounds, w);
    EE_CHECK_LAUNCH();
    monotone_fix_kernel<<<1, 32, 0, st>>>(n, w);
    EE_CHECK_LAUNCH();
    EE_CUDA(cudaStreamSynchronize(st));
    dev_free(bounds);
}

void bisect_dev(int n, const double *d, const double *e, double *w)
{
    Context &c = ctx();
    cudaStream_t st = c.stream;
    if (n == 1) { EE_CUDA(cudaMemcpyAsync(w, d, sizeof(double), cudaMemcpyDeviceToDevice, st)); return; }
    double *e2 = (double *)dev_alloc(sizeof(double) * (n + 4));
    double *bounds = e2 + n;
    bisect_prep_kernel<<<1, 256, 0, st>>>(n, d, e, e2, bounds);
    EE_CHECK_LAUNCH();
    bisect_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, d, e2, bounds, w);
    EE_CHECK_LAUNCH();
    monotone_fix_kernel<<<1, 32, 0, st>>>(n, w);
    EE_CHECK_LAUNCH();
    EE_CUDA(cudaStreamSynchronize(st));
    dev_free(e2);
}

}  // namespace ee
