// ee_dc.cu -- symmetric tridiagonal eigensolver (Cuppen divide & conquer) on B200.
//
// Stands in for eigen_dc2 / dc2_FS (src/dc2.F:78, src/dc2_FS.F), i.e. the ScaLAPACK
// PDSTEDC-derived code (mx_pdstedc.F, mx_pdlaed0-3.F) that sits between eigen_trd and
// eigen_common_trbakwy on the critical path of mode 'A'.  SURVEY 8(f) row 1.
// The algorithm is the published one (Cuppen 1981; Gu & Eisenstat 1995; LAPACK
// dstedc/dlaed0-4 working notes): rank-one tearing, deflation, secular equation,
// Loewner-formula eigenvectors, merge by GEMM.  Written from that description, not from the
// reference's sources, and laid out for the GPU:
//   * leaves (<= 32) : one warp each, parallel-order Jacobi in shared memory (batched launch)
//   * deflation      : O(n) scalar logic on the host (the only host arithmetic; it decides
//                      sizes of the device launches), rotations applied by a device kernel
//   * secular equation: one warp per root, bracketed "middle way" rational iteration in the
//                      shifted variable tau = lambda - d_origin (high relative accuracy of
//                      every d_j - lambda_i, which is what makes the vectors orthogonal)
//   * z~ (Gu/Eisenstat) and the k x k secular eigenvectors: one CTA per row / column
//   * merge          : FP64 tensor-core GEMMs (ee_gemm.cu) exploiting the block structure of
//                      diag(Q1,Q2) (column types 1/2/3 as in dlaed2), i.e. half the flops.
// The eigenvector matrix is never sorted physically between levels; an index permutation is
// carried on the host and applied once when the caller's cyclic z is written.
//
// Memory / multi-GPU layout.  The k x k secular eigenvector matrix is never stored: it is generated in
// column blocks (<= 2 GB) that go straight into the merge GEMMs, so the work space is two n x n matrices.
// Two distributions of the eigenvector matrix Q over P ranks:
//   * replicated (EIGENEXA_B200_DC_ROWS=0, only while 4 n^2 doubles fit): merges >= 1024 are split by
//     column slices over the ranks and all-gathered;
//   * ROW distributed (the default; P = 1 trivially): row owner q = x + y px owns the rows g = q (mod P).  Every step of the algorithm acts on
//     rows independently (rotations, column gathers, Q_new(rows,:) = Q_old(rows,:) V) except the four rows
//     that form z (one small all-reduce per rank-one update); the O(n^2) secular work is replicated, no
//     eigenvector data moves until the end, where the ranks that share a grid row exchange column sets to
//     reach the caller's 2D cyclic layout.  Memory per rank: 2 n^2 / P doubles (N = 100000 on 8 GPUs: 20 GB).
//
// The same code solves the penta-diagonal matrix of eigen_prd (eigen_dcx, src/dcx.F:75,
// my_pdsxedc.F, my_pdlaed0.F:226-391): the coupling between the halves of a split at row m is the
// 2x2 block B = [T(m-2,m) 0 ; T(m-1,m) T(m-1,m+1)], i.e. the sum of two rank-one terms
//     sigma_1 x_1 e_m^T       x_1 = (T(m-2,m), T(m-1,m)) / sigma_1   on rows (m-2, m-1)
//     sigma_2 (+-e_{m-1}) e_{m+1}^T                                   sigma_2 = |T(m-1,m+1)|
// and sigma (x y^T + y x^T) = sigma [(x+y)(x+y)^T - x x^T - y y^T], so every merge is two
// successive rank-one updates: the first on the block-diagonal diag(Q1,Q2) (structured GEMM, as
// in the tridiagonal case, which is sigma_2 = 0), the second on the dense merged matrix.
#include "ee_common.cuh"
#include "ee_comm.h"
#include <numeric>
#include <chrono>

namespace ee {

namespace {

constexpr int LEAF = 32;
constexpr int DIST_MIN = 1024;   // merges at least this large are split over the ranks
constexpr double EPS = 2.220446049250313e-16;  // 2^-52
constexpr double HALF_EPS = 1.1102230246251565e-16;

// ---------------------------------------------------------------------------------------
// leaves: cyclic Jacobi with the round-robin parallel ordering, one warp per leaf
// ---------------------------------------------------------------------------------------
struct LeafDesc { int lo, sz; };

__global__ void __launch_bounds__(32) leaf_jacobi_kernel(const LeafDesc *leaves, const double *d, const double *e,
                                                         const double *e2, double *Q, long long ldq, double *dout,
                                                         int RP, int rw)
{
    __shared__ double A[LEAF][LEAF + 1];
    __shared__ double V[LEAF][LEAF + 1];
    __shared__ double cs[LEAF / 2][2];
    __shared__ int pr[LEAF / 2][2];
    __shared__ int order[LEAF];
    const LeafDesc L = leaves[blockIdx.x];
    const int s = L.sz, lane = threadIdx.x;
    for (int c = 0; c < LEAF; c++) { A[lane][c] = 0.0; V[lane][c] = (lane == c) ? 1.0 : 0.0; }
    __syncwarp();
    if (lane < s) {
        A[lane][lane] = d[L.lo + lane];
        if (lane > 0) { double t = e[L.lo + lane]; A[lane][lane - 1] = t; A[lane - 1][lane] = t; }
        if (e2 && lane > 1) { double t = e2[L.lo + lane]; A[lane][lane - 2] = t; A[lane - 2][lane] = t; }
    }
    __syncwarp();
    double nf = 0.0;
    for (int c = 0; c < LEAF; c++) nf += A[lane][c] * A[lane][c];
    for (int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
    const double floor_abs = sqrt(nf) * EPS * 1e-3 / LEAF;
    const int m = LEAF;  // tournament over 32 slots (rows >= s are decoupled identity rows)
    for (int sweep = 0; sweep < 60; sweep++) {
        int nrot = 0;
        for (int round = 0; round < m - 1; round++) {
            // round-robin pairing: slot 0 fixed, the others rotate
            if (lane < m / 2) {
                int a = (lane == 0) ? 0 : (round + lane - 1) % (m - 1) + 1;
                int b = (round + (m - 1) - lane - 1 + (m - 1)) % (m - 1) + 1;
                if (lane == 0) b = (round + m - 2) % (m - 1) + 1;
                int p = min(a, b), q = max(a, b);
                double c = 1.0, sn = 0.0;
                if (q < s && p != q) {
                    double apq = A[p][q], app = A[p][p], aqq = A[q][q];
                    if (fabs(apq) > floor_abs && fabs(apq) > HALF_EPS * 0.5 * sqrt(fabs(app) * fabs(aqq))) {
                        double theta = (aqq - app) / (2.0 * apq);
                        double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        sn = t * c;
                    }
                }
                pr[lane][0] = p; pr[lane][1] = q; cs[lane][0] = c; cs[lane][1] = sn;
            }
            __syncwarp();
            // columns: A <- A J, V <- V J   (lane = row)
            for (int k = 0; k < m / 2; k++) {
                const double c = cs[k][0], sn = cs[k][1];
                if (sn != 0.0) {
                    const int p = pr[k][0], q = pr[k][1];
                    double ap = A[lane][p], aq = A[lane][q];
                    A[lane][p] = c * ap - sn * aq; A[lane][q] = sn * ap + c * aq;
                    double vp = V[lane][p], vq = V[lane][q];
                    V[lane][p] = c * vp - sn * vq; V[lane][q] = sn * vp + c * vq;
                    nrot++;
                }
            }
            __syncwarp();
            // rows: A <- J^T A   (lane = column)
            for (int k = 0; k < m / 2; k++) {
                const double c = cs[k][0], sn = cs[k][1];
                if (sn != 0.0) {
                    const int p = pr[k][0], q = pr[k][1];
                    double ap = A[p][lane], aq = A[q][lane];
                    A[p][lane] = c * ap - sn * aq; A[q][lane] = sn * ap + c * aq;
                }
            }
            __syncwarp();
        }
        if (nrot == 0) break;
    }
    // ascending order of the diagonal (rank by counting), columns written in that order
    if (lane < s) {
        double v = A[lane][lane];
        int r = 0;
        for (int j = 0; j < s; j++) {
            double u = A[j][j];
            r += (u < v) || (u == v && j < lane);
        }
        order[r] = lane;
    }
    __syncwarp();
    if (lane < s) {
        // row L.lo + lane lives on rank (row mod RP) at local index row / RP
        if ((L.lo + lane) % RP == rw) {
            const int jl = (L.lo + lane) / RP;
            for (int cdx = 0; cdx < s; cdx++) {
                int src = order[cdx];
                Q[(long long)(L.lo + cdx) * ldq + jl] = V[lane][src];
            }
        }
        dout[L.lo + lane] = A[order[lane]][order[lane]];
    }
}

// z = Q_node^T w for a vector w with (at most) four non-zeros: rows m-2 .. m+1 around the split
struct ZSpec { int row[4]; double coef[4]; };
__global__ void gather_z_kernel(const double *Q, long long ldq, int lo, int ns, ZSpec zs, double *z, int RP, int rw)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ns) return;
    const double *col = Q + (long long)(lo + j) * ldq;
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++)
        if (zs.coef[t] != 0.0 && zs.row[t] % RP == rw) s = fma(zs.coef[t], col[zs.row[t] / RP], s);
    z[j] = s;   // row-distributed: the caller sums the contributions of the owners
}

struct Rot { int p, q; double c, s; };
// apply the deflation rotations in order: (q_p, q_q) <- (c q_p + s q_q, -s q_p + c q_q)
__global__ void apply_rot_kernel(double *Qb, long long ldq, int ns, const Rot *rots, int nrot)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ns) return;
    for (int t = 0; t < nrot; t++) {
        const Rot R = rots[t];
        double a = Qb[(long long)R.p * ldq + r], b = Qb[(long long)R.q * ldq + r];
        Qb[(long long)R.p * ldq + r] = R.c * a + R.s * b;
        Qb[(long long)R.q * ldq + r] = -R.s * a + R.c * b;
    }
}

// dst(:, g) = src(:, map[g])
__global__ void gather_cols_kernel(const double *src, long long lds, double *dst, long long ldd, int rows,
                                   const int *map, int ncols)
{
    for (int c = blockIdx.y; c < ncols; c += gridDim.y) {     // gridDim.y is capped at 32768 by the callers
        const double *s = src + (long long)map[c] * lds;
        double *d = dst + (long long)c * ldd;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) d[r] = s[r];
    }
}

// ---------------------------------------------------------------------------------------
// secular equation  1/rho + sum_j w_j^2 / (d_j - lambda) = 0 , one warp per root
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128) secular_kernel(int k, const double *__restrict__ dl, const double *__restrict__ w,
                                                      double rho, double *lam, double *tau_out, int *org_out)
{
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= k) return;
    const double rhoinv = 1.0 / rho;
    if (k == 1) {
        if (lane == 0) { double t = rho * w[0] * w[0]; tau_out[0] = t; org_out[0] = 0; lam[0] = dl[0] + t; }
        return;
    }
    const bool last = (i == k - 1);
    int K, p1, p2;
    double lo, hi, tau;
    if (!last) {
        const double di = dl[i], del = dl[i + 1] - di;
        const double wi = w[i], wi1 = w[i + 1];
        const double tm = 0.5 * del;
        double cr = 0.0;
        for (int j = lane; j < k; j += 32)
            if (j != i && j != i + 1) { double wj = w[j]; cr += wj * wj / ((dl[j] - di) - tm); }
        cr = wsum(cr) + rhoinv;
        const double fmid = cr + wi * wi / (-tm) + wi1 * wi1 / (del - tm);
        p1 = i; p2 = i + 1;
        if (fmid > 0.0) {
            K = i; lo = 0.0; hi = tm;
            double a = cr * del + wi * wi + wi1 * wi1, b = wi * wi * del;
            double sq = sqrt(fabs(a * a - 4.0 * b * cr));
            tau = (a > 0.0) ? 2.0 * b / (a + sq) : (a - sq) / (2.0 * cr);
        } else {
            K = i + 1; lo = -tm; hi = 0.0;
            double a = -cr * del + wi * wi + wi1 * wi1, b = wi1 * wi1 * del;
            double sq = sqrt(fabs(a * a + 4.0 * b * cr));
            tau = (a < 0.0) ? 2.0 * b / (a - sq) : -(a + sq) / (2.0 * cr);
        }
        if (!(tau > lo && tau < hi)) tau = 0.5 * (lo + hi);
    } else {
        K = k - 1; p1 = k - 2; p2 = k - 1;
        double s = 0.0;
        for (int j = lane; j < k; j += 32) s += w[j] * w[j];
        s = wsum(s);
        lo = 0.0; hi = rho * s * (1.0 + 4.0 * EPS);
        tau = 0.5 * hi;
    }
    const double dK = dl[K];
    for (int it = 0; it < 80; it++) {
        double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0, ab = 0.0;
        for (int j = lane; j < k; j += 32) {
            double dj = (dl[j] - dK) - tau;
            double wj = w[j];
            double t = wj / dj;
            double term = wj * t;
            if (j <= p1) { psi += term; dpsi += t * t; } else { phi += term; dphi += t * t; }
            ab += fabs(term);
        }
        psi = wsum(psi); dpsi = wsum(dpsi); phi = wsum(phi); dphi = wsum(dphi); ab = wsum(ab);
        const double g = rhoinv + psi + phi;
        const double dw = dpsi + dphi;
        const double tol = 8.0 * EPS * (rhoinv + ab) + EPS * fabs(tau) * dw;
        if (fabs(g) <= tol) break;
        if (g < 0.0) lo = fmax(lo, tau); else hi = fmin(hi, tau);
        if (hi - lo <= 2.0 * EPS * fmax(fabs(lo), fabs(hi))) { tau = 0.5 * (lo + hi); break; }
        const double D1 = (dl[p1] - dK) - tau, D2 = (dl[p2] - dK) - tau;
        const double C = g - D1 * dpsi - D2 * dphi;
        const double A = (D1 + D2) * g - D1 * D2 * dw;
        const double B = D1 * D2 * g;
        // roots of C eta^2 - A eta + B = 0
        double e1, e2;
        if (C == 0.0) { e1 = (A != 0.0) ? B / A : 0.0; e2 = e1; }
        else {
            double disc = A * A - 4.0 * B * C;
            if (disc < 0.0) disc = 0.0;
            double qq = 0.5 * (A + copysign(sqrt(disc), A));
            e1 = (qq != 0.0) ? B / qq : 0.0;
            e2 = qq / C;
        }
        auto ok = [&](double eta) { double t = tau + eta; return (g * eta < 0.0) && (t > lo) && (t < hi); };
        double eta;
        const bool o1 = ok(e1), o2 = ok(e2);
        if (o1 && o2) eta = (fabs(e1) < fabs(e2)) ? e1 : e2;
        else if (o1) eta = e1;
        else if (o2) eta = e2;
        else {
            eta = -g / dw;  // Newton (f is increasing between the poles)
            if (!ok(eta)) eta = 0.5 * (lo + hi) - tau;
        }
        const double tn = tau + eta;
        if (tn == tau) break;
        tau = tn;
    }
    if (lane == 0) { tau_out[i] = tau; org_out[i] = K; lam[i] = dK + tau; }
}

// Gu/Eisenstat: z~_j = sign(w_j) sqrt( | prod_i (d_j - lam_i) / prod_{i != j} (d_j - d_i) | )
__global__ void __launch_bounds__(128) loewner_kernel(int k, const double *__restrict__ dl, const double *__restrict__ w,
                                                      const double *__restrict__ tau, const int *__restrict__ org, double *zt)
{
    __shared__ double sp[4];
    const int j = blockIdx.x;
    const double dj = dl[j];
    double p = 1.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        double num = (dj - dl[org[i]]) - tau[i];
        if (i != j) p *= num / (dj - dl[i]); else p *= num;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = sp[0] * sp[1] * sp[2] * sp[3];
        zt[j] = copysign(sqrt(fabs(t)), w[j]);
    }
}

// column i of the secular eigenvector matrix: V(rowperm[j], i) = z~_j / (d_j - lam_i), normalised
__global__ void __launch_bounds__(256) secvec_kernel(int k, const double *__restrict__ dl, const double *__restrict__ zt,
                                                     const double *__restrict__ tau, const int *__restrict__ org,
                                                     const int *__restrict__ rowperm, double *V, long long ldv, int c0)
{
    __shared__ double sred[8];
    __shared__ double s_inv;
    const int i = c0 + blockIdx.x;     // columns c0 .. c0 + gridDim.x - 1 of the secular eigenvector matrix
    const double dK = dl[org[i]], ti = tau[i];
    double s = 0.0;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        double v = zt[j] / ((dl[j] - dK) - ti);
        s = fma(v, v, s);
    }
    s = wsum(s);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; q++) t += sred[q];
        s_inv = 1.0 / sqrt(t);
    }
    __syncthreads();
    const double inv = s_inv;
    double *col = V + (long long)blockIdx.x * ldv;
    for (int j = threadIdx.x; j < k; j += blockDim.x) col[rowperm[j]] = zt[j] / ((dl[j] - dK) - ti) * inv;
}

// final extraction: z_loc(jl, il) = Q(jl*px + x, ord[il*py + y])
__global__ void extract_kernel(const double *Q, long long ldq, const int *ord, int n, int nvec, int px, int py, int x,
                               int y, double *z, int ldz, int nrl, int nvl)
{
    for (int il = blockIdx.y; il < nvl; il += gridDim.y) {
        const double *src = Q + (long long)ord[il * py + y] * ldq;
        double *dst = z + (size_t)il * ldz;
        for (int jl = blockIdx.x * blockDim.x + threadIdx.x; jl < nrl; jl += gridDim.x * blockDim.x)
            dst[jl] = src[(long long)jl * px + x];
    }
}

// row-distributed output, step 1: S_y(jl, il) = Qloc(jl, ord[il*py + y]) for every destination y of the grid row
__global__ void pack_rows_kernel(const double *Q, long long ldq, const int *ord, int nvec, int py, int nrow, double *S,
                                 const long long *off)
{
    for (int c = blockIdx.y; c < nvec; c += gridDim.y) {   // column position in ascending eigenvalue order
        const int y = c % py, il = c / py;
        const double *src = Q + (long long)ord[c] * ldq;
        double *dst = S + off[y] + (long long)il * nrow;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += gridDim.x * blockDim.x) dst[r] = src[r];
    }
}
// step 2: z_loc(jl, il) = R_{jl mod py}(jl / py, il); R_y' has nrow(y') rows
__global__ void unpack_rows_kernel(const double *R, const long long *off, const int *nrows, int py, int nrl, int nvl,
                                   double *z, int ldz)
{
    for (int il = blockIdx.y; il < nvl; il += gridDim.y)
        for (int jl = blockIdx.x * blockDim.x + threadIdx.x; jl < nrl; jl += gridDim.x * blockDim.x) {
            const int ys = jl % py, t = jl / py;
            z[(size_t)il * ldz + jl] = R[off[ys] + (long long)il * nrows[ys] + t];
        }
}

struct Node { int lo, mid, hi; };

void build_tree(int lo, int hi, std::vector<Node> &merges, std::vector<LeafDesc> &leaves)
{
    if (hi - lo <= LEAF) { leaves.push_back({lo, hi - lo}); return; }
    int mid = lo + (hi - lo) / 2;
    build_tree(lo, mid, merges, leaves);
    build_tree(mid, hi, merges, leaves);
    merges.push_back({lo, mid, hi});  // post-order: children first
}

}  // namespace

int dc_band_dev(int n, int nvec, const double *d_in, const double *e_in, const double *e2_in, double *w_out, double *z,
                int ldz);
int dc_dev(int n, int nvec, const double *d_in, const double *e_in, double *w_out, double *z, int ldz)
{
    return dc_band_dev(n, nvec, d_in, e_in, nullptr, w_out, z, ldz);
}

// e_in(i) = T(i-1,i); e2_in(i) = T(i-2,i) or nullptr for a tridiagonal matrix
int dc_band_dev(int n, int nvec, const double *d_in, const double *e_in, const double *e2_in, double *w_out, double *z,
                int ldz)
{
    Context &c = ctx();
    const Grid &g = c.g;
    cudaStream_t st = c.stream;
    const int nrl = cyc_count(n, g.px, g.x);
    const int nvl = cyc_count(nvec, g.py, g.y);

    const bool penta = e2_in != nullptr;
    std::vector<double> hd(n), he(n), he2(n, 0.0);
    EE_CUDA(cudaMemcpyAsync(hd.data(), d_in, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    EE_CUDA(cudaMemcpyAsync(he.data(), e_in, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (penta) EE_CUDA(cudaMemcpyAsync(he2.data(), e2_in, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    EE_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++)
        if (!std::isfinite(hd[i]) || !std::isfinite(he[i]) || !std::isfinite(he2[i])) { set_error("dc: non-finite band matrix"); return 1; }

    std::vector<Node> merges;
    std::vector<LeafDesc> leaves;
    build_tree(0, n, merges, leaves);
    // tearing at every split point m (blocks are >= 16 rows, so the touched entries never overlap):
    //   first term : sigma_1 = |(T(m-2,m), T(m-1,m))|, x = that vector / sigma_1 on rows (m-2,m-1), y = e_m
    //   second term: sigma_2 = |T(m-1,m+1)|, x = +-e_{m-1}, y = e_{m+1}
    // diag(T1,T2) loses sigma x x^T and sigma y y^T  (tridiagonal: d(m-1) -= |e|, d(m) -= |e|)
    struct Tear { double sig1, x0, x1, sig2, s2; };
    std::vector<Tear> tears(merges.size());
    for (size_t t = 0; t < merges.size(); t++) {
        const int m = merges[t].mid;
        Tear tr = {0, 0, 0, 0, 0};
        const double b0 = he2[m], b1 = he[m];
        tr.sig1 = hypot(b0, b1);
        if (tr.sig1 > 0.0) {
            tr.x0 = b0 / tr.sig1; tr.x1 = b1 / tr.sig1;
            if (b0 != 0.0) { hd[m - 2] -= tr.sig1 * tr.x0 * tr.x0; he[m - 1] -= tr.sig1 * tr.x0 * tr.x1; }
            hd[m - 1] -= tr.sig1 * tr.x1 * tr.x1;
            hd[m] -= tr.sig1;
        }
        if (penta && m + 1 < merges[t].hi && he2[m + 1] != 0.0) {
            tr.sig2 = fabs(he2[m + 1]); tr.s2 = he2[m + 1] > 0.0 ? 1.0 : -1.0;
            hd[m - 1] -= tr.sig2; hd[m + 1] -= tr.sig2;
        }
        tears[t] = tr;
    }

    // ---- distribution of the eigenvector matrix -------------------------------------------------------
    const int P = g.nnod;
    // Row-distributed is the default on a grid: the merge GEMMs (all but a few % of the D&C time at N = 50000) are
    // divided by P at every level and no eigenvector data moves until the final exchange; the replicated form
    // (EIGENEXA_B200_DC_ROWS=0) all-gathers every merged block and needs 4 n^2 doubles per rank.
    bool rows_mode = true;
    if (P > 1) {
        const char *env = getenv("EIGENEXA_B200_DC_ROWS");
        const double repl_bytes = 4.2 * (double)n * (double)n * sizeof(double);   // Q, Q2, slice + gather buffers
        if (env && env[0] == '0' && repl_bytes <= 100e9) rows_mode = false;
    }
    // row partition: row gr belongs to the row owner q = gr % RP, and rank (x, y) IS owner q = x + y px whatever
    // the rank order of eigen_init ('C' or 'R'): the exchange below relies on jl % py = y'
    const int RP = rows_mode ? P : 1, rw = rows_mode ? g.x + g.y * g.px : 0;
    auto world_rank_of = [&](int x, int y) { return g.order == 'R' ? x * g.py + y : x + y * g.px; };
    auto lrows = [&](int gcount) { return cyc_count(gcount, RP, rw); };  // owned rows with global index < gcount
    const int nrow_loc = lrows(n);
    // (+4 rows of slack in the distributed form: the final exchange reuses Q / Q2 as receive / send buffers of
    //  nrl x nvl <= (n/px + 1)(n/py + 1) doubles)
    const long long ldq = (((long long)(nrow_loc > 0 ? nrow_loc : 1)) + (rows_mode && P > 1 ? 4 : 0) + 15) & ~15LL;
    const size_t qbytes = (size_t)ldq * n * sizeof(double);
    double *Q = (double *)dev_alloc(qbytes);
    double *Q2 = (merges.empty() && !(rows_mode && P > 1)) ? nullptr : (double *)dev_alloc(qbytes);
    // secular eigenvectors: generated in column blocks of at most VS_CAP doubles
    const long long VS_CAP = 1LL << 28;
    auto block_cols = [&](int k) -> int {
        if ((long long)(k + 1) * k <= VS_CAP) return k;
        long long kb = VS_CAP / (k + 1);
        kb = kb / 64 * 64;
        return (int)(kb < 64 ? 64 : kb);
    };
    const long long vs_doubles = std::min((long long)(n + 1) * n, VS_CAP + (long long)n * 66);
    double *Vs = merges.empty() ? nullptr : (double *)dev_alloc((size_t)vs_doubles * sizeof(double));
    // replicated multi-rank merges: this rank's column slice and the gathered block
    const bool slices = (P > 1) && !rows_mode;
    double *Tb = nullptr, *Gb = nullptr;
    if (slices && n >= DIST_MIN) {
        const size_t sl = (size_t)(((n + P - 1) / P + 2) & ~1);
        Tb = (double *)dev_alloc((size_t)(n + 2) * sl * sizeof(double));
        Gb = (double *)dev_alloc((size_t)(n + 2) * sl * P * sizeof(double));
    }
    EE_CUDA(cudaMemsetAsync(Q, 0, qbytes, st));
    // small device arrays
    double *dd = (double *)dev_alloc(sizeof(double) * n * 9);
    double *d_d = dd, *d_e = dd + n, *d_z = dd + 2 * n, *d_dl = dd + 3 * n, *d_w = dd + 4 * n, *d_tau = dd + 5 * n,
           *d_lam = dd + 6 * n, *d_zt = dd + 7 * n, *d_e2 = dd + 8 * n;
    int *di = (int *)dev_alloc(sizeof(int) * n * 4);
    int *d_org = di, *d_rowperm = di + n, *d_map = di + 2 * n, *d_ord = di + 3 * n;
    Rot *d_rot = (Rot *)dev_alloc(sizeof(Rot) * (n + 1));
    LeafDesc *d_leaves = (LeafDesc *)dev_alloc(sizeof(LeafDesc) * leaves.size());
    EE_CUDA(cudaMemcpyAsync(d_d, hd.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    EE_CUDA(cudaMemcpyAsync(d_e, he.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (penta) EE_CUDA(cudaMemcpyAsync(d_e2, he2.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    EE_CUDA(cudaMemcpyAsync(d_leaves, leaves.data(), sizeof(LeafDesc) * leaves.size(), cudaMemcpyHostToDevice, st));

    // ---- leaves (every rank solves all of them, keeps its rows) -----------------------------------------
    leaf_jacobi_kernel<<<(unsigned)leaves.size(), 32, 0, st>>>(d_leaves, d_d, d_e, penta ? d_e2 : nullptr, Q, ldq, d_lam, RP, rw);
    EE_CHECK_LAUNCH();
    std::vector<double> D(n);  // eigenvalues in physical column order
    EE_CUDA(cudaMemcpyAsync(D.data(), d_lam, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    EE_CUDA(cudaStreamSynchronize(st));
    std::vector<int> ord(n);  // ord[lo + t] = physical column (relative to lo) of the t-th smallest of the node
    for (const LeafDesc &L : leaves)
        for (int t = 0; t < L.sz; t++) ord[L.lo + t] = t;

    std::vector<double> hz(n), dl(n), ww(n), lam(n);
    std::vector<int> idx(n), coltype(n), nondefl, defl, grouped(n), rowperm(n), map(n), neword(n);
    std::vector<Rot> rots;

    double dc_flops = 0.0; long long n_defl_total = 0;
    // breakdown (seconds): [17] z gather + host deflation [18] rotations + column gather
    // [19] secular/Loewner [20] secular vectors + merge GEMMs + deflated copy [21] host sort
    cudaEvent_t evs[5];
    for (auto &e : evs) EE_CUDA(cudaEventCreate(&e));
    double t_defl = 0, t_perm = 0, t_sec = 0, t_gemm = 0, t_sort = 0;
    auto wall = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    // ---- merges (post-order) -----------------------------------------------------------------
    // one rank-one update  diag(D) + rho z z^T  of the node [m.lo, m.hi), z = Q_node^T w / sqrt(2) (w given by
    // zs, |w|^2 = 2).  blockdiag: the node's eigenvector matrix is still diag(Q1, Q2) and ord holds the two
    // children's orders; otherwise it is dense and ord holds the node's own order.
    auto rank_one_update = [&](const Node &m, double rho, const ZSpec &zs, bool blockdiag) {
        const int lo = m.lo, n1 = m.mid - m.lo, ns = m.hi - m.lo, n2 = ns - n1;
        // local rows of the node: r1 of the upper child, r2 of the lower one
        const int jl0 = lrows(lo), r1 = lrows(m.mid) - jl0, r2 = lrows(m.hi) - lrows(m.mid), rs = r1 + r2;
        double *Qb = Q + (long long)lo * ldq + jl0, *Q2b = Q2 + (long long)lo * ldq + jl0;
        double *Dn = D.data() + lo;
        double tw0 = wall();
        gather_z_kernel<<<(ns + 255) / 256, 256, 0, st>>>(Q, ldq, lo, ns, zs, d_z, RP, rw);
        EE_CHECK_LAUNCH();
        if (RP > 1) comm_allreduce_sum(d_z, (size_t)ns, COMM_WORLD, st);
        EE_CUDA(cudaMemcpyAsync(hz.data(), d_z, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
        EE_CUDA(cudaStreamSynchronize(st));
        const double isq2 = 1.0 / sqrt(2.0);
        for (int j = 0; j < ns; j++) hz[j] *= isq2;
        if (blockdiag) {
            // ascending order over both children (merge of the two sorted index lists)
            int a = 0, b = 0, t = 0;
            const int *oa = ord.data() + lo, *ob = ord.data() + lo + n1;
            while (a < n1 && b < n2) {
                if (Dn[oa[a]] <= Dn[n1 + ob[b]]) idx[t++] = oa[a++]; else idx[t++] = n1 + ob[b++];
            }
            while (a < n1) idx[t++] = oa[a++];
            while (b < n2) idx[t++] = n1 + ob[b++];
        } else {
            for (int t = 0; t < ns; t++) idx[t] = ord[lo + t];
        }
        double dmax = 0.0, zmax = 0.0;
        for (int j = 0; j < ns; j++) { dmax = fmax(dmax, fabs(Dn[j])); zmax = fmax(zmax, fabs(hz[j])); }
        const double tol = 8.0 * HALF_EPS * fmax(dmax, zmax);
        if (rho * zmax <= tol) {
            for (int t = 0; t < ns; t++) ord[lo + t] = idx[t];
            return;
        }
        // ---- deflation (dlaed2 logic) ---------------------------------------------------
        nondefl.clear(); defl.clear(); rots.clear();
        for (int j = 0; j < ns; j++) coltype[j] = !blockdiag ? 2 : (j < n1) ? 1 : 3;
        int pj = -1;
        for (int t = 0; t < ns; t++) {
            const int nj = idx[t];
            if (rho * fabs(hz[nj]) <= tol) { coltype[nj] = 4; defl.push_back(nj); continue; }
            if (pj < 0) { pj = nj; continue; }
            double s = hz[pj], cc = hz[nj];
            const double tau = hypot(cc, s);
            const double tt = Dn[nj] - Dn[pj];
            cc /= tau; s = -s / tau;
            if (fabs(tt * cc * s) <= tol) {
                hz[nj] = tau; hz[pj] = 0.0;
                if (coltype[nj] != coltype[pj]) coltype[nj] = 2;
                coltype[pj] = 4;
                rots.push_back({pj, nj, cc, s});
                const double t1 = Dn[pj] * cc * cc + Dn[nj] * s * s;
                Dn[nj] = Dn[pj] * s * s + Dn[nj] * cc * cc;
                Dn[pj] = t1;
                defl.push_back(pj);
                pj = nj;
            } else {
                nondefl.push_back(pj);
                pj = nj;
            }
        }
        if (pj >= 0) nondefl.push_back(pj);
        const int k = (int)nondefl.size();
        // group the surviving columns by type (1: top only, 2: dense, 3: bottom only)
        int k1 = 0, k2 = 0, k3 = 0;
        for (int i = 0; i < k; i++) { int ty = coltype[nondefl[i]]; k1 += ty == 1; k2 += ty == 2; k3 += ty == 3; }
        {
            int p1 = 0, p2 = k1, p3 = k1 + k2;
            for (int i = 0; i < k; i++) {
                int col = nondefl[i], ty = coltype[col];
                int pos = (ty == 1) ? p1++ : (ty == 2) ? p2++ : p3++;
                grouped[pos] = col; rowperm[i] = pos;
                dl[i] = Dn[col]; ww[i] = hz[col];
            }
        }
        for (int gidx = 0; gidx < k; gidx++) map[gidx] = grouped[gidx];
        for (size_t t = 0; t < defl.size(); t++) map[k + t] = defl[t];
        t_defl += wall() - tw0;
        // ---- device: rotations, permuted copy, secular system, merge GEMMs ------------
        EE_CUDA(cudaEventRecord(evs[0], st));
        if (!rots.empty() && rs > 0) {
            EE_CUDA(cudaMemcpyAsync(d_rot, rots.data(), sizeof(Rot) * rots.size(), cudaMemcpyHostToDevice, st));
            apply_rot_kernel<<<(rs + 127) / 128, 128, 0, st>>>(Qb, ldq, rs, d_rot, (int)rots.size());
            EE_CHECK_LAUNCH();
        }
        EE_CUDA(cudaMemcpyAsync(d_map, map.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
        if (rs > 0) {
            dim3 grid(std::min(16, (rs + 255) / 256), std::min(ns, 32768));
            gather_cols_kernel<<<grid, 256, 0, st>>>(Qb, ldq, Q2b, ldq, rs, d_map, ns);
            EE_CHECK_LAUNCH();
        }
        const long long ldv = ((long long)k + 1) & ~1LL;
        EE_CUDA(cudaEventRecord(evs[1], st));
        EE_CUDA(cudaMemcpyAsync(d_dl, dl.data(), sizeof(double) * k, cudaMemcpyHostToDevice, st));
        EE_CUDA(cudaMemcpyAsync(d_w, ww.data(), sizeof(double) * k, cudaMemcpyHostToDevice, st));
        EE_CUDA(cudaMemcpyAsync(d_rowperm, rowperm.data(), sizeof(int) * k, cudaMemcpyHostToDevice, st));
        secular_kernel<<<(k + 3) / 4, 128, 0, st>>>(k, d_dl, d_w, rho, d_lam, d_tau, d_org);
        EE_CHECK_LAUNCH();
        EE_CUDA(cudaMemcpyAsync(lam.data(), d_lam, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
        loewner_kernel<<<k, 128, 0, st>>>(k, d_dl, d_w, d_tau, d_org, d_zt);
        EE_CHECK_LAUNCH();
        const int k12 = k1 + k2, k23 = k2 + k3;
        EE_CUDA(cudaEventRecord(evs[2], st));
        n_defl_total += ns - k;
        dc_flops += 2.0 * (double)k * ((double)n1 * k12 + (double)n2 * k23);   // as mx_pdlaed1.F:291,304 counts them
        // new columns [c0, c1) produced by this rank into dst (ld ldd): secular vectors of a column block,
        // then  top rows = Q2(0:r1, 0:k12) V(0:k12, blk)  and  bottom rows = Q2(r1:, k1:k) V(k1:k, blk)
        auto produce = [&](int c0, int c1, double *dst, long long ldd) {
            const int kbmax = block_cols(k);
            for (int cb = c0; cb < c1; cb += kbmax) {
                const int kb = std::min(kbmax, c1 - cb);
                secvec_kernel<<<kb, 256, 0, st>>>(k, d_dl, d_zt, d_tau, d_org, d_rowperm, Vs, ldv, cb);
                EE_CHECK_LAUNCH();
                double *dcol = dst + (long long)(cb - c0) * ldd;
                if (r1 > 0) {
                    if (k12 > 0) dgemm_ex(st, 'N', 'N', r1, kb, k12, 1.0, Q2b, ldq, Vs, ldv, 0.0, dcol, ldd, 1, 0);
                    else EE_CUDA(cudaMemset2DAsync(dcol, ldd * sizeof(double), 0, (size_t)r1 * sizeof(double), kb, st));
                }
                if (r2 > 0) {
                    if (k23 > 0) dgemm_ex(st, 'N', 'N', r2, kb, k23, 1.0, Q2b + r1 + (long long)k1 * ldq, ldq, Vs + k1, ldv, 0.0, dcol + r1, ldd, 1, 0);
                    else EE_CUDA(cudaMemset2DAsync(dcol + r1, ldd * sizeof(double), 0, (size_t)r2 * sizeof(double), kb, st));
                }
            }
        };
        if (slices && ns >= DIST_MIN && k >= 2 * P) {
            // replicated Q: every rank forms its slice of the k merged columns, then one all-gather puts the
            // whole block on every rank (everything else is replicated, so all ranks take identical
            // deflation decisions in the next level)
            int slice = ((k + P - 1) / P + 1) & ~1;
            const long long ldt = ((long long)ns + 1) & ~1LL;
            const int c0 = std::min(k, g.inod * slice), c1 = std::min(k, c0 + slice);
            if (c1 > c0) produce(c0, c1, Tb, ldt);
            comm_allgather(Tb, Gb, (size_t)ldt * slice, COMM_WORLD, st);
            EE_CUDA(cudaMemcpy2DAsync(Qb, ldq * sizeof(double), Gb, ldt * sizeof(double), (size_t)ns * sizeof(double), k,
                                      cudaMemcpyDeviceToDevice, st));
        } else {
            produce(0, k, Qb, ldq);
        }
        if (ns > k && rs > 0)
            EE_CUDA(cudaMemcpy2DAsync(Qb + (long long)k * ldq, ldq * sizeof(double), Q2b + (long long)k * ldq, ldq * sizeof(double),
                                      (size_t)rs * sizeof(double), ns - k, cudaMemcpyDeviceToDevice, st));
        EE_CUDA(cudaEventRecord(evs[3], st));
        EE_CUDA(cudaStreamSynchronize(st));
        {
            float ms;
            cudaEventElapsedTime(&ms, evs[0], evs[1]); t_perm += ms * 1e-3;
            cudaEventElapsedTime(&ms, evs[1], evs[2]); t_sec += ms * 1e-3;
            cudaEventElapsedTime(&ms, evs[2], evs[3]); t_gemm += ms * 1e-3;
        }
        double tw1 = wall();
        // ---- new physical eigenvalues and their ascending order -----------------------
        {
            std::vector<double> tmp(ns);
            for (int i = 0; i < k; i++) tmp[i] = lam[i];
            for (size_t t = 0; t < defl.size(); t++) tmp[k + t] = Dn[defl[t]];
            for (int j = 0; j < ns; j++) Dn[j] = tmp[j];
            for (int j = 0; j < ns; j++) neword[j] = j;
            std::stable_sort(neword.begin(), neword.begin() + ns, [&](int a, int b) { return Dn[a] < Dn[b]; });
            for (int j = 0; j < ns; j++) ord[lo + j] = neword[j];
        }
        t_sort += wall() - tw1;
    };
    for (size_t t = 0; t < merges.size(); t++) {
        const Node &m = merges[t];
        const Tear &tr = tears[t];
        bool blockdiag = true;
        if (tr.sig1 > 0.0) {
            ZSpec zs = {{m.mid - 2, m.mid - 1, m.mid, m.mid}, {tr.x0, tr.x1, 1.0, 0.0}};
            if (tr.x0 == 0.0) zs.row[0] = m.mid - 1;   // tridiagonal split next to the block edge: never read
            rank_one_update(m, 2.0 * tr.sig1, zs, true);
            blockdiag = false;
        }
        if (tr.sig2 > 0.0) {
            ZSpec zs = {{m.mid - 1, m.mid + 1, m.mid, m.mid}, {tr.s2, 1.0, 0.0, 0.0}};
            rank_one_update(m, 2.0 * tr.sig2, zs, blockdiag);
            blockdiag = false;
        }
        if (blockdiag) {
            // no coupling at all: the node's order is the merge of the children's orders
            const int lo = m.lo, n1 = m.mid - m.lo, ns = m.hi - m.lo, n2 = ns - n1;
            const double *Dn = D.data() + lo;
            int a = 0, b = 0, q = 0;
            const int *oa = ord.data() + lo, *ob = ord.data() + lo + n1;
            while (a < n1 && b < n2) {
                if (Dn[oa[a]] <= Dn[n1 + ob[b]]) idx[q++] = oa[a++]; else idx[q++] = n1 + ob[b++];
            }
            while (a < n1) idx[q++] = oa[a++];
            while (b < n2) idx[q++] = n1 + ob[b++];
            for (int j = 0; j < ns; j++) ord[lo + j] = idx[j];
        }
    }
    for (auto &e : evs) cudaEventDestroy(e);
    c.timings[17] = t_defl; c.timings[18] = t_perm; c.timings[19] = t_sec; c.timings[20] = t_gemm; c.timings[21] = t_sort;
    // ---- outputs: w ascending, z = local cyclic part of Q(:, ord) ---------------------------------
    {
        std::vector<double> wv(n);
        for (int i = 0; i < n; i++) wv[i] = D[ord[i]];
        EE_CUDA(cudaMemcpyAsync(w_out, wv.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
        EE_CUDA(cudaMemcpyAsync(d_ord, ord.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        if (!(rows_mode && P > 1)) {
            // the rank holds every row it needs
            if (nrl > 0 && nvl > 0) {
                dim3 grid(std::min(16, (nrl + 255) / 256), std::min(nvl, 32768));
                extract_kernel<<<grid, 256, 0, st>>>(Q, ldq, d_ord, n, nvec, g.px, g.py, g.x, g.y, z, ldz, nrl, nvl);
                EE_CHECK_LAUNCH();
            }
        } else {
            // Row-distributed: rank (x, y') owns the rows g = x + y' px (mod P), i.e. the local rows jl = y' (mod py)
            // of grid row x.  Pack, for every rank (x, y) of the grid row, the columns it owns (position c = y mod py
            // in ascending order), exchange inside the grid row, interleave the received row sets.
            const int px = g.px, py = g.py;
            std::vector<long long> soff(py + 1, 0), roff(py + 1, 0);
            std::vector<int> nrows(py, 0);
            for (int y = 0; y < py; y++) {
                soff[y + 1] = soff[y] + (long long)nrow_loc * cyc_count(nvec, py, y);
                nrows[y] = cyc_count(n, P, g.x + y * px);
                roff[y + 1] = roff[y] + (long long)nrows[y] * nvl;
            }
            if (soff[py] > (long long)ldq * n || roff[py] > (long long)ldq * n)
                fatal("dc: exchange buffers exceed the eigenvector work space", __FILE__, __LINE__);
            long long *d_off = (long long *)dev_alloc(sizeof(long long) * 2 * (py + 1));
            int *d_nrows = (int *)dev_alloc(sizeof(int) * py);
            EE_CUDA(cudaMemcpyAsync(d_off, soff.data(), sizeof(long long) * (py + 1), cudaMemcpyHostToDevice, st));
            EE_CUDA(cudaMemcpyAsync(d_off + py + 1, roff.data(), sizeof(long long) * (py + 1), cudaMemcpyHostToDevice, st));
            EE_CUDA(cudaMemcpyAsync(d_nrows, nrows.data(), sizeof(int) * py, cudaMemcpyHostToDevice, st));
            if (nrow_loc > 0 && nvec > 0) {
                dim3 grid(std::min(16, (nrow_loc + 255) / 256), std::min(nvec, 32768));
                pack_rows_kernel<<<grid, 256, 0, st>>>(Q, ldq, d_ord, nvec, py, nrow_loc, Q2, d_off);
                EE_CHECK_LAUNCH();
            }
            // Q is free now: it receives the row sets of the grid row
            comm_group_start();
            for (int y = 0; y < py; y++) {
                if (y == g.y) continue;
                const int peer = world_rank_of(g.x, y);
                const size_t scount = (size_t)(soff[y + 1] - soff[y]), rcount = (size_t)(roff[y + 1] - roff[y]);
                if (scount) comm_send(Q2 + soff[y], scount, peer, st);
                if (rcount) comm_recv(Q + roff[y], rcount, peer, st);
            }
            comm_group_end();
            {
                const size_t self = (size_t)(soff[g.y + 1] - soff[g.y]);
                if (self) EE_CUDA(cudaMemcpyAsync(Q + roff[g.y], Q2 + soff[g.y], self * sizeof(double), cudaMemcpyDeviceToDevice, st));
            }
            if (nrl > 0 && nvl > 0) {
                dim3 grid(std::min(16, (nrl + 255) / 256), std::min(nvl, 32768));
                unpack_rows_kernel<<<grid, 256, 0, st>>>(Q, d_off + py + 1, d_nrows, py, nrl, nvl, z, ldz);
                EE_CHECK_LAUNCH();
            }
            EE_CUDA(cudaStreamSynchronize(st));
            dev_free(d_off); dev_free(d_nrows);
        }
        EE_CUDA(cudaStreamSynchronize(st));
    }
    c.timings[13] = dc_flops; c.timings[14] = (double)n_defl_total;
    dev_free(Q); if (Q2) dev_free(Q2); if (Vs) dev_free(Vs);
    if (Tb) dev_free(Tb);
    if (Gb) dev_free(Gb);
    dev_free(dd); dev_free(di); dev_free(d_rot); dev_free(d_leaves);
    return 0;
}

}  // namespace ee
