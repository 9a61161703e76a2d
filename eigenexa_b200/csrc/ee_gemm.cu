// ee_gemm.cu -- FP64 tensor-core GEMM for sm_100a.
//
// Blackwell's tcgen05/TMEM path has no f64 kind; FP64 tensor math on sm_100a is the warp
// level mma.sync m8n8k4.f64 (SASS: DMMA.8x8x4 -- the wider PTX shapes m16n8k8/k16 are
// split into this one by ptxas).  The kernel below is a cp.async multi-stage pipeline
// feeding DMMA tiles; it serves
//   * the trailing rank-2k update of eigen_common_2update (src/eigen_t1.F:250-306),
//     as one 'N','T' GEMM with K = 2m restricted to the upper staircase,
//   * the compact-WY back-transformation GEMMs of eigen_trbakwy_block_body
//     (src/trbakwy4_body.F:573-577,604-608,721-725): 'T','N' (V^T Z, V^T V) and 'N','N',
//   * the merge GEMMs of the tridiagonal divide & conquer.
//
// Fragment mapping: the MMA is issued on the TRANSPOSED tile (mma-M <- C columns,
// mma-N <- C rows) so that each thread owns two vertically adjacent elements of the
// column-major C and the epilogue moves 16 B per access.  Both operands are then read from
// shared memory as element (idx = lane>>2, k = lane&3); the two shared layouts
// ([k][idx] with pitch TILE+4, [idx][k] with pitch KC+4) are bank-conflict free for 64-bit
// accesses (pitch == 4 mod 16 doubles).
#include "ee_common.cuh"

namespace ee {

namespace {

constexpr int KC = 16;      // K chunk per stage
constexpr int STAGES = 3;

struct GemmP {
    int M, N, K;
    double alpha, beta;
    const double *A; long long lda;
    const double *B; long long ldb;
    double *C; long long ldc;
    int ksplit; long long kper; long long c_stride;  // split-K: blockIdx.z handles K range, C += z*c_stride
    int tri, px, py, x, y;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *g, int bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *g, int bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// load one operand tile (TILE indices x KC k) into shared memory
//  KCONT = false: global element (idx,k) at base[idx + k*ld]   -> smem [k][TILE+4]
//  KCONT = true : global element (idx,k) at base[k + idx*ld]   -> smem [idx][KC+4]
template <int TILE, bool KCONT, bool AL16, int NT>
__device__ __forceinline__ void load_tile(double *sm, const double *base, long long ld, int i0, int ilim, long long k0,
                                          long long klim)
{
    if (!KCONT) {
        constexpr int PI = TILE + 4;
        constexpr int CH = TILE / 2;
        for (int q = threadIdx.x; q < KC * CH; q += NT) {
            int k = q / CH, idx = (q % CH) * 2;
            long long gk = k0 + k; int gi = i0 + idx;
            int nval = (gk < klim) ? min(2, max(0, ilim - gi)) : 0;
            const double *src = nval > 0 ? base + gk * ld + gi : base;
            double *dst = sm + k * PI + idx;
            if (AL16) cp_async16(dst, src, nval * 8);
            else {
                cp_async8(dst, src, nval > 0 ? 8 : 0);
                cp_async8(dst + 1, nval > 1 ? src + 1 : base, nval > 1 ? 8 : 0);
            }
        }
    } else {
        constexpr int PK = KC + 4;
        constexpr int CH = KC / 2;
        for (int q = threadIdx.x; q < TILE * CH; q += NT) {
            int idx = q / CH, k = (q % CH) * 2;
            long long gk = k0 + k; int gi = i0 + idx;
            int nval = (gi < ilim) ? (int)min(2LL, max(0LL, klim - gk)) : 0;
            const double *src = nval > 0 ? base + (long long)gi * ld + gk : base;
            double *dst = sm + idx * PK + k;
            if (AL16) cp_async16(dst, src, nval * 8);
            else {
                cp_async8(dst, src, nval > 0 ? 8 : 0);
                cp_async8(dst + 1, nval > 1 ? src + 1 : base, nval > 1 ? 8 : 0);
            }
        }
    }
}

template <int TILE, bool KCONT>
__device__ __forceinline__ double frag(const double *sm, int idx, int k)
{
    if (!KCONT) return sm[k * (TILE + 4) + idx];
    return sm[idx * (KC + 4) + k];
}

template <int TILE, bool KCONT>
constexpr int tile_doubles() { return KCONT ? TILE * (KC + 4) : KC * (TILE + 4); }

// BM x BN CTA tile, WM x WN warps (WM along rows m, WN along cols n)
template <int BM, int BN, int WM, int WN, bool A_KCONT, bool B_KCONT, bool AL16, int MINB>
__global__ void __launch_bounds__(WM * WN * 32, MINB) dgemm_kernel(GemmP p)
{
    constexpr int NT = WM * WN * 32;
    constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
    constexpr int MF = WTM / 8, NF = WTN / 8;    // 8x8 fragments
    constexpr int SA = tile_doubles<BM, A_KCONT>(), SB = tile_doubles<BN, B_KCONT>();
    extern __shared__ __align__(16) double smem[];

    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (p.tri) {
        // skip tiles strictly below the staircase (global row > global col for every element)
        long long gr = (long long)m0 * p.px + p.x;
        long long gc = (long long)(min(n0 + BN, p.N) - 1) * p.py + p.y;
        if (gr > gc) return;
    }
    const long long kbeg = (long long)blockIdx.z * p.kper;
    const long long kend = min((long long)p.K, kbeg + p.kper);
    double *C = p.C + (long long)blockIdx.z * p.c_stride;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int fi = lane >> 2, fk = lane & 3;

    double acc[NF][MF][2];
#pragma unroll
    for (int a = 0; a < NF; a++)
#pragma unroll
        for (int b = 0; b < MF; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    const bool vec = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    // alpha = +-1 (every read-modify-write use on this path): start the accumulators from C so
    // that the C reads overlap the operand pipeline instead of sitting in the epilogue
    const bool preload = (p.beta != 0.0) && (p.alpha == 1.0 || p.alpha == -1.0);
    if (preload) {
        const double sc = p.beta * p.alpha;
#pragma unroll
        for (int a = 0; a < NF; a++) {
            const int n = n0 + wn * WTN + a * 8 + fi;
#pragma unroll
            for (int b = 0; b < MF; b++) {
                const int m = m0 + wm * WTM + b * 8 + 2 * fk;
                if (n < p.N && m < p.M) {
                    const double *cp = C + (long long)n * p.ldc + m;
                    if (vec && m + 1 < p.M) {
                        double2 o = __ldcs(reinterpret_cast<const double2 *>(cp));
                        acc[a][b][0] = sc * o.x; acc[a][b][1] = sc * o.y;
                    } else {
                        acc[a][b][0] = sc * cp[0];
                        if (m + 1 < p.M) acc[a][b][1] = sc * cp[1];
                    }
                }
            }
        }
    }

    const int nk = (int)((kend - kbeg + KC - 1) / KC);
    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) {
            double *sa = smem + s * (SA + SB), *sb = sa + SA;
            load_tile<BM, A_KCONT, AL16, NT>(sa, p.A, p.lda, m0, p.M, kbeg + (long long)s * KC, kend);
            load_tile<BN, B_KCONT, AL16, NT>(sb, p.B, p.ldb, n0, p.N, kbeg + (long long)s * KC, kend);
        }
        cp_commit();
    }
    for (int it = 0; it < nk; it++) {
        cp_wait<STAGES - 2>();
        __syncthreads();
        // prefetch stage it+STAGES-1 (its buffer was consumed in iteration it-1)
        {
            int nx = it + STAGES - 1;
            if (nx < nk) {
                double *sa = smem + (nx % STAGES) * (SA + SB), *sb = sa + SA;
                load_tile<BM, A_KCONT, AL16, NT>(sa, p.A, p.lda, m0, p.M, kbeg + (long long)nx * KC, kend);
                load_tile<BN, B_KCONT, AL16, NT>(sb, p.B, p.ldb, n0, p.N, kbeg + (long long)nx * KC, kend);
            }
            cp_commit();
        }
        const double *sa = smem + (it % STAGES) * (SA + SB), *sb = sa + SA;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double af[MF], bf[NF];
#pragma unroll
            for (int b = 0; b < MF; b++) af[b] = frag<BM, A_KCONT>(sa, wm * WTM + b * 8 + fi, kk + fk);
#pragma unroll
            for (int a = 0; a < NF; a++) bf[a] = frag<BN, B_KCONT>(sb, wn * WTN + a * 8 + fi, kk + fk);
#pragma unroll
            for (int a = 0; a < NF; a++)
#pragma unroll
                for (int b = 0; b < MF; b++) dmma(acc[a][b][0], acc[a][b][1], bf[a], af[b]);
        }
    }
    cp_wait<0>();

    // epilogue: thread owns C(m..m+1, n)
    const double beta = preload ? 0.0 : p.beta;
#pragma unroll
    for (int a = 0; a < NF; a++) {
        const int n = n0 + wn * WTN + a * 8 + fi;
        if (n >= p.N) continue;
#pragma unroll
        for (int b = 0; b < MF; b++) {
            const int m = m0 + wm * WTM + b * 8 + 2 * fk;
            if (m >= p.M) continue;
            double *cp = C + (long long)n * p.ldc + m;
            double r0 = p.alpha * acc[a][b][0], r1 = p.alpha * acc[a][b][1];
            if (m + 1 < p.M) {
                if (vec) {
                    if (beta != 0.0) {
                        double2 o = *reinterpret_cast<const double2 *>(cp);
                        r0 = fma(beta, o.x, r0); r1 = fma(beta, o.y, r1);
                    }
                    *reinterpret_cast<double2 *>(cp) = make_double2(r0, r1);
                } else {
                    if (beta != 0.0) { r0 = fma(beta, cp[0], r0); r1 = fma(beta, cp[1], r1); }
                    cp[0] = r0; cp[1] = r1;
                }
            } else {
                if (beta != 0.0) r0 = fma(beta, cp[0], r0);
                cp[0] = r0;
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, bool AK, bool BK, int MINB>
void launch_cfg(cudaStream_t st, const GemmP &p, bool al16)
{
    constexpr int SA = tile_doubles<BM, AK>(), SB = tile_doubles<BN, BK>();
    constexpr size_t smem = (size_t)STAGES * (SA + SB) * sizeof(double);
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, p.ksplit);
    dim3 block(WM * WN * 32);
    if (al16) {
        auto kern = dgemm_kernel<BM, BN, WM, WN, AK, BK, true, MINB>;
        static bool set = false;
        if (!set) { EE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
        kern<<<grid, block, smem, st>>>(p);
    } else {
        auto kern = dgemm_kernel<BM, BN, WM, WN, AK, BK, false, MINB>;
        static bool set = false;
        if (!set) { EE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
        kern<<<grid, block, smem, st>>>(p);
    }
    EE_CHECK_LAUNCH();
}

static int gemm_cfg()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("EIGENEXA_B200_GEMM_CFG"); v = e ? atoi(e) : 0; }
    return v;
}

template <bool AK, bool BK>
void launch_layout(cudaStream_t st, const GemmP &p, bool al16)
{
    // measured on B200 (tools/gemm_sweep.py, profiles/r01_gemm_sweep.md): 128x64 tiles with 16
    // warps per CTA and 2 CTAs/SM (32 resident warps hide the LDS->DMMA and barrier latency)
    // beat the 8-warp 128x128 / 128x64 variants on every shape of the path.
    const int cfg = gemm_cfg();
    if (cfg == 1) { launch_cfg<128, 128, 4, 4, AK, BK, 1>(st, p, al16); return; }
    if (cfg == 4) {
        if (p.K / p.ksplit >= 512 && p.beta == 0.0) launch_cfg<128, 128, 2, 4, AK, BK, 1>(st, p, al16);
        else launch_cfg<128, 64, 2, 4, AK, BK, 2>(st, p, al16);
        return;
    }
    launch_cfg<128, 64, 4, 4, AK, BK, 2>(st, p, al16);
}

}  // namespace

void dgemm_ex(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha, const double *A, long long lda,
              const double *B, long long ldb, double beta, double *C, long long ldc, int ksplit, long long c_stride,
              TriSpec tri)
{
    if (M <= 0 || N <= 0) return;
    GemmP p;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.ksplit = ksplit < 1 ? 1 : ksplit;
    long long kper = (K + p.ksplit - 1) / p.ksplit;
    kper = (kper + KC - 1) / KC * KC;
    p.kper = kper; p.c_stride = c_stride;
    p.tri = tri.mode; p.px = tri.px; p.py = tri.py; p.x = tri.x; p.y = tri.y;
    const bool ak = (transA == 'T' || transA == 't');   // A^T: stored K x M -> k contiguous
    const bool bk = (transB == 'N' || transB == 'n');   // B stored K x N -> k contiguous
    const bool al16 = ((lda & 1) == 0) && ((ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    if (ak && bk) launch_layout<true, true>(st, p, al16);
    else if (ak && !bk) launch_layout<true, false>(st, p, al16);
    else if (!ak && bk) launch_layout<false, true>(st, p, al16);
    else launch_layout<false, false>(st, p, al16);
}

void dgemm(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha, const double *A, int lda,
           const double *B, int ldb, double beta, double *C, int ldc, TriSpec tri)
{
    dgemm_ex(st, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 1, 0, tri);
}

}  // namespace ee
