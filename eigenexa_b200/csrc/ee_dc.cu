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
                                                         const double *e2, double *Q, long long ldq, double *dout)
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
        for (int cdx = 0; cdx < s; cdx++) {
            int src = order[cdx];
            Q[(long long)(L.lo + cdx) * ldq + L.lo + lane] = V[lane][src];
        }
        dout[L.lo + lane] = A[order[lane]][order[lane]];
    }
}

// z = Q_node^T w for a vector w with (at most) four non-zeros: rows m-2 .. m+1 around the split
struct ZSpec { int row[4]; double coef[4]; };
__global__ void gather_z_kernel(const double *Q, long long ldq, int lo, int ns, ZSpec zs, double *z)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ns) return;
    const double *col = Q + (long long)(lo + j) * ldq;
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++)
        if (zs.coef[t] != 0.0) s = fma(zs.coef[t], col[zs.row[t]], s);
    z[j] = s;
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
    int c = blockIdx.y;
    if (c >= ncols) return;
    const double *s = src + (long long)map[c] * lds;
    double *d = dst + (long long)c * ldd;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) d[r] = s[r];
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
                                                     const int *__restrict__ rowperm, double *V, long long ldv)
{
    __shared__ double sred[8];
    __shared__ double s_inv;
    const int i = blockIdx.x;
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
    double *col = V + (long long)i * ldv;
    for (int j = threadIdx.x; j < k; j += blockDim.x) col[rowperm[j]] = zt[j] / ((dl[j] - dK) - ti) * inv;
}

// final extraction: z_loc(jl, il) = Q(jl*px + x, ord[il*py + y])
__global__ void extract_kernel(const double *Q, long long ldq, const int *ord, int n, int nvec, int px, int py, int x,
                               int y, double *z, int ldz, int nrl, int nvl)
{
    int il = blockIdx.y;
    if (il >= nvl) return;
    const double *src = Q + (long long)ord[il * py + y] * ldq;
    double *dst = z + (size_t)il * ldz;
    for (int jl = blockIdx.x * blockDim.x + threadIdx.x; jl < nrl; jl += gridDim.x * blockDim.x)
        dst[jl] = src[(long long)jl * px + x];
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

    const long long ldq = ((long long)n + 15) & ~15LL;
    const size_t qbytes = (size_t)ldq * n * sizeof(double);
    double *Q = (double *)dev_alloc(qbytes);
    double *Q2 = merges.empty() ? nullptr : (double *)dev_alloc(qbytes);
    double *Vs = merges.empty() ? nullptr : (double *)dev_alloc(qbytes);
    // multi-rank merge buffers: this rank's column slice and the gathered block
    double *Tb = nullptr, *Gb = nullptr;
    if (g.nnod > 1 && n >= DIST_MIN) {
        const size_t sl = (size_t)(((n + g.nnod - 1) / g.nnod + 2) & ~1);
        Tb = (double *)dev_alloc((size_t)(n + 2) * sl * sizeof(double));
        Gb = (double *)dev_alloc((size_t)(n + 2) * sl * g.nnod * sizeof(double));
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

    // ---- leaves ------------------------------------------------------------------------------
    leaf_jacobi_kernel<<<(unsigned)leaves.size(), 32, 0, st>>>(d_leaves, d_d, d_e, penta ? d_e2 : nullptr, Q, ldq, d_lam);
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
    // breakdown (seconds): [16] leaves [17] z gather + host deflation [18] rotations + column gather
    // [19] secular/Loewner/vectors [20] merge GEMMs + deflated copy [21] host sort
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
        double *Qb = Q + (long long)lo * ldq + lo, *Q2b = Q2 + (long long)lo * ldq + lo;
        double *Dn = D.data() + lo;
        double tw0 = wall();
        gather_z_kernel<<<(ns + 255) / 256, 256, 0, st>>>(Q, ldq, lo, ns, zs, d_z);
        EE_CHECK_LAUNCH();
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
        if (!rots.empty()) {
            EE_CUDA(cudaMemcpyAsync(d_rot, rots.data(), sizeof(Rot) * rots.size(), cudaMemcpyHostToDevice, st));
            apply_rot_kernel<<<(ns + 127) / 128, 128, 0, st>>>(Qb, ldq, ns, d_rot, (int)rots.size());
            EE_CHECK_LAUNCH();
        }
        EE_CUDA(cudaMemcpyAsync(d_map, map.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
        {
            dim3 grid(std::min(16, (ns + 255) / 256), ns);
            gather_cols_kernel<<<grid, 256, 0, st>>>(Qb, ldq, Q2b, ldq, ns, d_map, ns);
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
        secvec_kernel<<<k, 256, 0, st>>>(k, d_dl, d_zt, d_tau, d_org, d_rowperm, Vs, ldv);
        EE_CHECK_LAUNCH();
        const int k12 = k1 + k2, k23 = k2 + k3;
        EE_CUDA(cudaEventRecord(evs[2], st));
        n_defl_total += ns - k;
        dc_flops += 2.0 * (double)k * ((double)n1 * k12 + (double)n2 * k23);   // as mx_pdlaed1.F:291,304 counts them
        if (g.nnod > 1 && ns >= DIST_MIN && k >= 2 * g.nnod) {
            // multi-rank: every rank forms its slice of the k merged columns, then one
            // all-gather puts the whole block on every rank (everything else is replicated, so
            // all ranks take identical deflation decisions in the next level)
            const int P = g.nnod;
            int slice = ((k + P - 1) / P + 1) & ~1;
            const long long ldt = ((long long)ns + 1) & ~1LL;
            const int c0 = std::min(k, g.inod * slice), c1 = std::min(k, c0 + slice);
            const int wdt = c1 - c0;
            if (wdt > 0) {
                if (k12 > 0) dgemm_ex(st, 'N', 'N', n1, wdt, k12, 1.0, Q2b, ldq, Vs + (long long)c0 * ldv, ldv, 0.0, Tb, ldt, 1, 0);
                else EE_CUDA(cudaMemset2DAsync(Tb, ldt * sizeof(double), 0, (size_t)n1 * sizeof(double), wdt, st));
                if (k23 > 0) dgemm_ex(st, 'N', 'N', n2, wdt, k23, 1.0, Q2b + n1 + (long long)k1 * ldq, ldq, Vs + k1 + (long long)c0 * ldv, ldv, 0.0, Tb + n1, ldt, 1, 0);
                else EE_CUDA(cudaMemset2DAsync(Tb + n1, ldt * sizeof(double), 0, (size_t)n2 * sizeof(double), wdt, st));
            }
            comm_allgather(Tb, Gb, (size_t)ldt * slice, COMM_WORLD, st);
            EE_CUDA(cudaMemcpy2DAsync(Qb, ldq * sizeof(double), Gb, ldt * sizeof(double), (size_t)ns * sizeof(double), k,
                                      cudaMemcpyDeviceToDevice, st));
        } else {
        if (k12 > 0) dgemm_ex(st, 'N', 'N', n1, k, k12, 1.0, Q2b, ldq, Vs, ldv, 0.0, Qb, ldq, 1, 0);
        else EE_CUDA(cudaMemset2DAsync(Qb, ldq * sizeof(double), 0, (size_t)n1 * sizeof(double), k, st));
        if (k23 > 0) dgemm_ex(st, 'N', 'N', n2, k, k23, 1.0, Q2b + n1 + (long long)k1 * ldq, ldq, Vs + k1, ldv, 0.0, Qb + n1, ldq, 1, 0);
        else EE_CUDA(cudaMemset2DAsync(Qb + n1, ldq * sizeof(double), 0, (size_t)n2 * sizeof(double), k, st));
        }
        if (ns > k)
            EE_CUDA(cudaMemcpy2DAsync(Qb + (long long)k * ldq, ldq * sizeof(double), Q2b + (long long)k * ldq, ldq * sizeof(double),
                                      (size_t)ns * sizeof(double), ns - k, cudaMemcpyDeviceToDevice, st));
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
    const double SQ2 = sqrt(2.0);
    (void)SQ2;
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
        if (nrl > 0 && nvl > 0) {
            dim3 grid(std::min(16, (nrl + 255) / 256), nvl);
            extract_kernel<<<grid, 256, 0, st>>>(Q, ldq, d_ord, n, nvec, g.px, g.py, g.x, g.y, z, ldz, nrl, nvl);
            EE_CHECK_LAUNCH();
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
