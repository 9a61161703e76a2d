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
//
// Two kernels share those layouts:
//  * dgemm_tma_kernel (the product path): persistent, warp specialised.  One producer warp
//    moves operand tiles global -> shared with TMA (cp.async.bulk.tensor.2d, SASS UTMALDG)
//    through a ring of mbarrier-guarded stages; the consumer warps never meet a CTA-wide
//    barrier, so the DMMA pipe only drains when every warp of an SM sub-partition waits at
//    once.  The padding of the conflict-free layouts comes for free from a TMA box that is
//    4 elements wider than the tile (the extra elements are neighbouring data / zero fill
//    and are never read).  Tiles are handed out round-robin over the valid (staircase) tiles,
//    so the next tile's operands stream in while the current one runs its epilogue.
//  * dgemm_kernel (cp.async + __syncthreads): operands that TMA cannot describe (odd leading
//    dimension or base pointer not 16-byte aligned).
#include "ee_common.cuh"
#include <cuda.h>

namespace ee {

namespace {

constexpr int KC = 16;      // K chunk per stage
constexpr int STAGES = 3;   // cp.async kernel

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

// ---------------------------------------------------------------------------------------
// TMA + mbarrier, warp-specialised, persistent
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred P1;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    // a protocol error must abort the launch (the host then fails loudly), never hang the GPU
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

struct TileP {
    int tiles_m, tiles_n;      // tile grid
    int group_w;               // rasterisation: tile columns per group (full GEMMs)
    int total_items;           // valid tiles * ksplit
    unsigned stagger_ns;       // > 0: groups start up to this many ns apart (see dgemm_tma_kernel)
    int dbg;                   // EIGENEXA_B200_GEMM_DBG bits: 1 no fast preload, 2 no fast epilogue, 4 no stagger, 8 no C prefetch
};

// maps the linear index of a VALID tile to its coordinates; indices must be queried in
// increasing order (the staircase walk is incremental)
template <int BM, int BN>
struct TileWalk {
    int j = 0, base = 0;
    __device__ __forceinline__ int rows_of(const GemmP &p, const TileP &tp, int jj) const
    {
        const int gc = (min((jj + 1) * BN, p.N) - 1) * p.py + p.y;   // N * py < 2^31 (checked on the host)
        if (gc < p.x) return 0;
        return min((gc - p.x) / (BM * p.px) + 1, tp.tiles_m);
    }
    __device__ __forceinline__ void locate(const GemmP &p, const TileP &tp, int tile, int &tm, int &tn)
    {
        if (p.tri) {
            int nr = rows_of(p, tp, j);
            while (tile >= base + nr) { base += nr; j++; nr = rows_of(p, tp, j); }
            tn = j; tm = tile - base;
        } else {
            const int per_group = tp.group_w * tp.tiles_m;
            const int g = tile / per_group;
            const int r = tile - g * per_group;
            const int gw = min(tp.group_w, tp.tiles_n - g * tp.group_w);
            tm = r / gw; tn = g * tp.group_w + r % gw;
        }
    }
};

// One CTA per SM holds NG independent "groups" (virtual CTAs), each NCW consumer warps with
// their own stage ring and barriers, plus ONE producer warpgroup (warp p of it feeds group p).
// Two groups per SM overlap one group's epilogue / C read-modify-write with the other's math.
// The register file is per SM sub-partition (16 K registers, one warp of every warpgroup):
// the producer warpgroup shrinks to 24 registers (setmaxnreg.dec) and the consumer warpgroups
// grow to what that frees (setmaxnreg.inc), as in the sm_90/sm_100 warp-specialised GEMMs.
constexpr int tma_threads(int ng, int ncw) { return (ng * ncw + 4) * 32; }
constexpr int PRODUCER_REGS = 32;
// registers per thread at launch (what __launch_bounds__(NT, 1) lets ptxas assume) ...
constexpr int launch_regs(int nt) { return (65536 / nt > 255 ? 255 : 65536 / nt) / 8 * 8; }
// ... and what each consumer warpgroup may grow to once the producer warpgroup has shrunk:
// setmaxnreg.inc blocks until the CTA's pool holds the registers, so the sum over warpgroups
// must not exceed the launch allocation
constexpr int consumer_regs(int ng, int ncw)
{
    const int wgs = ng * ncw / 4 + 1;
    const int r = (wgs * launch_regs(tma_threads(ng, ncw)) - PRODUCER_REGS) / (wgs - 1) / 8 * 8;
    return r > 248 ? 248 : r;
}
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

__device__ __forceinline__ void l2_prefetch(const void *g, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

template <int BM, int BN, int NST, int WM, int WN, int NG, bool A_KCONT, bool B_KCONT>
__global__ void __launch_bounds__(tma_threads(NG, WM * WN), 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmP p, TileP tp)
{
    constexpr int NCW = WM * WN;                 // consumer warps per group
    static_assert((NG * NCW) % 4 == 0, "consumer warps must fill whole warpgroups");
    constexpr int WTM = BM / WM, WTN = BN / WN;
    constexpr int MF = WTM / 8, NF = WTN / 8;
    constexpr int SA = tile_doubles<BM, A_KCONT>(), SB = tile_doubles<BN, B_KCONT>();
    constexpr unsigned STAGE_BYTES = (unsigned)((SA + SB) * sizeof(double));
    constexpr unsigned RING_BYTES = NST * STAGE_BYTES;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte aligned stage rings (TMA destinations), then the barriers of every group
    const unsigned raw = smem_u32(smem_raw);
    const unsigned pad = ((raw + 127u) & ~127u) - raw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp >= NG * NCW;
    const int grp = producer ? warp - NG * NCW : warp / NCW;
    const int wl = producer ? NCW : warp - grp * NCW;
    // group g: full[s] at barg + 8 s ; empty[s] at barg + 8 (NST + s)
    const unsigned bar_all = raw + pad + NG * RING_BYTES;
    if (threadIdx.x == 0) {
        for (int g = 0; g < NG; g++)
            for (int s = 0; s < NST; s++) {
                mbar_init(bar_all + 16 * NST * g + 8 * s, 1);
                mbar_init(bar_all + 16 * NST * g + 8 * (NST + s), NCW);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const unsigned ring = raw + pad + grp * RING_BYTES;
    const double *stages = reinterpret_cast<const double *>(smem_raw + pad + grp * RING_BYTES);
    const unsigned bar0 = bar_all + 16 * NST * grp;
    const int vbid = blockIdx.x * NG + grp, vgrid = gridDim.x * NG;

    TileWalk<BM, BN> walk;
    int s = 0; unsigned ph = 0;
    if (producer) {
        reg_dec<PRODUCER_REGS>();
        if (grp >= NG) return;   // spare warps of the producer warpgroup
        // ================= producer warp: lane 0 drives the TMA ring; all lanes prefetch the next
        // tile's C into L2 when the GEMM accumulates into C =====================================
        TileWalk<BM, BN> ahead;
        const bool pf = !(tp.dbg & 8) && (p.beta != 0.0) && ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        for (int item = vbid; item < tp.total_items; item += vgrid) {
            if (pf && item + vgrid < tp.total_items) {
                const int nitem = item + vgrid;
                const int ntile = nitem / p.ksplit;
                int am, an;
                ahead.locate(p, tp, ntile, am, an);
                const int rows = min(BM, p.M - am * BM) & ~1;
                if (rows > 0) {
                    const double *cb = p.C + (long long)(nitem - ntile * p.ksplit) * p.c_stride + am * BM;
                    for (int c = lane; c < BN; c += 32) {
                        const int n = an * BN + c;
                        if (n < p.N) l2_prefetch(cb + (long long)n * p.ldc, (unsigned)rows * 8u);
                    }
                }
            }
            if (lane == 0) {
                const int tile = item / p.ksplit;
                const int z = item - tile * p.ksplit;
                int tm, tn;
                walk.locate(p, tp, tile, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;
                const int kbeg = z * (int)p.kper;
                const int kend = min(p.K, kbeg + (int)p.kper);
                const int nk = (kend - kbeg + KC - 1) / KC;
                for (int kc = 0; kc < nk; kc++) {
                    mbar_wait(bar0 + 8 * (NST + s), ph ^ 1u);
                    const unsigned full = bar0 + 8 * s;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const unsigned da = ring + s * STAGE_BYTES, db = da + SA * (unsigned)sizeof(double);
                    const int k0 = kbeg + kc * KC;
                    if (A_KCONT) tma_load_2d(da, &mapA, full, k0, m0); else tma_load_2d(da, &mapA, full, m0, k0);
                    if (B_KCONT) tma_load_2d(db, &mapB, full, k0, n0); else tma_load_2d(db, &mapB, full, n0, k0);
                    if (++s == NST) { s = 0; ph ^= 1u; }
                }
            }
            __syncwarp();
        }
        return;
    }
    // ===================== consumers: DMMA on the landed stages ============================
    reg_inc<consumer_regs(NG, NCW)>();
    // Every tile costs the same, so groups that start together stay in lockstep and all of them would
    // read their C tiles (DRAM) at the same instant while every DMMA pipe idles.  Spreading the start
    // times over one tile period (golden-ratio sequence of the group id) keeps the C traffic of some
    // groups under the math of the others for the rest of the launch.
    if (tp.stagger_ns) {
        const unsigned frac = ((unsigned)vbid * 0x9E3779B1u) >> 16;               // 0 .. 65535
        unsigned wait = (unsigned)(((unsigned long long)tp.stagger_ns * frac) >> 16);
        while (wait > 0) { const unsigned step = wait > 100000u ? 100000u : wait; __nanosleep(step); wait -= step; }
    }
    const int wm = wl % WM, wn = wl / WM;
    const int fi = lane >> 2, fk = lane & 3;
    int pend = -1;   // stage consumed last, not yet handed back to the producer (the last one never needs to be)
    for (int item = vbid; item < tp.total_items; item += vgrid) {
        const int tile = item / p.ksplit;
        const int z = item - tile * p.ksplit;
        int tm, tn;
        walk.locate(p, tp, tile, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        const int kbeg = z * (int)p.kper;
        const int kend = min(p.K, kbeg + (int)p.kper);
        const int nk = (kend - kbeg + KC - 1) / KC;
        double *C = p.C + (long long)z * p.c_stride;
        const bool vec = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);

        double acc[NF][MF][2];
        const bool preload = (p.beta != 0.0) && (p.alpha == 1.0 || p.alpha == -1.0);
        // interior tile with 16-byte aligned C: no bounds checks, so that all C loads of the tile are
        // in flight together (a branch per element would expose every load's latency in turn)
        const bool interior = vec && (m0 + BM <= p.M) && (n0 + BN <= p.N);
        if (preload && interior && !(tp.dbg & 1)) {
            const double sc = p.beta * p.alpha;
            const double *cb = C + (long long)(n0 + wn * WTN + fi) * p.ldc + m0 + wm * WTM + 2 * fk;
            double2 cv[NF][MF];
#pragma unroll
            for (int a = 0; a < NF; a++)
#pragma unroll
                for (int b = 0; b < MF; b++)
                    cv[a][b] = __ldcs(reinterpret_cast<const double2 *>(cb + (long long)(a * 8) * p.ldc + b * 8));
#pragma unroll
            for (int a = 0; a < NF; a++)
#pragma unroll
                for (int b = 0; b < MF; b++) { acc[a][b][0] = sc * cv[a][b].x; acc[a][b][1] = sc * cv[a][b].y; }
        } else if (preload) {
            const double sc = p.beta * p.alpha;
#pragma unroll
            for (int a = 0; a < NF; a++) {
                const int n = n0 + wn * WTN + a * 8 + fi;
#pragma unroll
                for (int b = 0; b < MF; b++) {
                    const int m = m0 + wm * WTM + b * 8 + 2 * fk;
                    acc[a][b][0] = acc[a][b][1] = 0.0;
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
        } else {
#pragma unroll
            for (int a = 0; a < NF; a++)
#pragma unroll
                for (int b = 0; b < MF; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
        }

        for (int kc = 0; kc < nk; kc++) {
            mbar_wait(bar0 + 8 * s, ph);
            // Release the PREVIOUS stage only now.  ptxas is free to schedule an arrive right behind the last
            // LDS of a stage (the remaining DMMAs only touch registers), i.e. before that load has returned,
            // and the producer's TMA would then overwrite shared memory under an in-flight read.  Behind this
            // wait every DMMA of the previous stage has issued, so all of its fragment loads have completed.
            if (pend >= 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (NST + pend));
            }
            const double *sa = stages + (size_t)s * (SA + SB), *sb = sa + SA;
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
            pend = s;
            if (++s == NST) { s = 0; ph ^= 1u; }
        }

        const double beta = preload ? 0.0 : p.beta;
        if (interior && beta == 0.0 && !(tp.dbg & 2)) {
            double *cb = C + (long long)(n0 + wn * WTN + fi) * p.ldc + m0 + wm * WTM + 2 * fk;
#pragma unroll
            for (int a = 0; a < NF; a++)
#pragma unroll
                for (int b = 0; b < MF; b++)
                    *reinterpret_cast<double2 *>(cb + (long long)(a * 8) * p.ldc + b * 8) =
                        make_double2(p.alpha * acc[a][b][0], p.alpha * acc[a][b][1]);
            continue;
        }
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
}

// ---- host side: tensor maps ---------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        EE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !sym) fatal("cuTensorMapEncodeTiled is not available", __FILE__, __LINE__);
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}
// operand with indices (idx, k): kcont -> element at base[k + idx*ld], else base[idx + k*ld]
static void make_operand_map(CUtensorMap *map, const double *base, long long nidx, long long nk, long long ld, bool kcont,
                             int tile)
{
    cuuint64_t dims[2], strides[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (kcont) { dims[0] = (cuuint64_t)nk; dims[1] = (cuuint64_t)nidx; box[0] = KC + 4; box[1] = (cuuint32_t)tile; }
    else { dims[0] = (cuuint64_t)nidx; dims[1] = (cuuint64_t)nk; box[0] = (cuuint32_t)tile + 4; box[1] = KC; }
    strides[0] = (cuuint64_t)ld * sizeof(double);
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu ld %lld box %u x %u", (int)r,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], ld, box[0], box[1]);
        fatal(buf, __FILE__, __LINE__);
    }
}

template <int BM, int BN>
static long long count_tiles(const GemmP &p, int tiles_m, int tiles_n)
{
    if (!p.tri) return (long long)tiles_m * tiles_n;
    long long tot = 0;
    for (int j = 0; j < tiles_n; j++) {
        long long gc = (long long)(std::min((j + 1) * BN, p.N) - 1) * p.py + p.y;
        if (gc < p.x) continue;
        long long nr = (gc - p.x) / ((long long)BM * p.px) + 1;
        tot += std::min(nr, (long long)tiles_m);
    }
    return tot;
}

// deepest ring (<= NSTMAX stages) that fits the 227 KB of shared memory of one CTA
template <int BM, int BN, int NSTMAX, int NG, bool AK, bool BK>
constexpr int fit_stages()
{
    constexpr long long stage = (long long)(tile_doubles<BM, AK>() + tile_doubles<BN, BK>()) * 8 + 16;
    constexpr long long fit = (232448 - 128) / (NG * stage);
    return (int)(fit < NSTMAX ? fit : NSTMAX);
}

template <int BM, int BN, int NSTMAX, int WM, int WN, int NG, bool AK, bool BK>
void launch_tma(cudaStream_t st, const GemmP &p)
{
    constexpr int NST = fit_stages<BM, BN, NSTMAX, NG, AK, BK>();
    static_assert(NST >= 2, "ring needs at least two stages");
    constexpr int SA = tile_doubles<BM, AK>(), SB = tile_doubles<BN, BK>();
    constexpr size_t smem = (size_t)NG * NST * (SA + SB) * sizeof(double) + (size_t)NG * 2 * NST * sizeof(unsigned long long) + 128;
    static_assert(smem <= 232448, "stage rings exceed the 227 KB of shared memory per CTA");
    auto kern = dgemm_tma_kernel<BM, BN, NST, WM, WN, NG, AK, BK>;
    static bool set = false;
    if (!set) { EE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
    TileP tp;
    tp.tiles_m = (p.M + BM - 1) / BM; tp.tiles_n = (p.N + BN - 1) / BN;
    const long long tiles = count_tiles<BM, BN>(p, tp.tiles_m, tp.tiles_n);
    if (tiles <= 0) return;
    if (tiles * p.ksplit > 0x7fffffffLL || (long long)p.N * p.py > 0x7fffffffLL || (long long)BM * p.px > 0x7fffffffLL)
        fatal("dgemm: problem too large for 32-bit tile arithmetic", __FILE__, __LINE__);
    tp.total_items = (int)(tiles * p.ksplit);
    // concurrently running tiles cover ~ (resident / group_w) x group_w : keep it near square in elements
    tp.group_w = std::max(1, std::min(tp.tiles_n, 16 * 64 / BN));
    // C read-modify-write with many tiles per group: stagger the groups over one tile period
    // (2 BM BN K flop at 1/NG of an SM's ~250 GFLOP/s)
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("EIGENEXA_B200_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
    tp.dbg = dbg;
    tp.stagger_ns = 0;
    if (!(dbg & 4) && p.beta != 0.0 && tp.total_items >= 4 * NG * ctx().sm_count) {
        const double tile_ns = 2.0 * BM * BN * (double)p.kper / 250.0 * NG;
        tp.stagger_ns = (unsigned)std::min(tile_ns, 400000.0);
    }
    CUtensorMap mapA, mapB;
    make_operand_map(&mapA, p.A, p.M, p.K, p.lda, AK, BM);
    make_operand_map(&mapB, p.B, p.N, p.K, p.ldb, BK, BN);
    const long long ctas = ((long long)tp.total_items + NG - 1) / NG;
    const unsigned grid = (unsigned)std::min(ctas, (long long)ctx().sm_count);
    kern<<<grid, tma_threads(NG, WM * WN), smem, st>>>(mapA, mapB, p, tp);
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
    const int cfg = gemm_cfg();
    if (!al16 || p.K <= 0 || cfg >= 10) {
        // operands TMA cannot describe: cp.async kernel (cfg >= 10 forces it, for A/B measurements)
        if (cfg == 11) launch_cfg<128, 128, 4, 4, AK, BK, 1>(st, p, al16);
        else launch_cfg<128, 64, 4, 4, AK, BK, 2>(st, p, al16);
        return;
    }
    // measured on B200 (tools/gemm_sweep.py, profiles/r01_gemm_sweep.md): both shapes reach cuBLAS on deep K;
    // two 128x64 groups per SM overlap the epilogue and quantise small problems better
    if (cfg == 2) launch_tma<128, 128, 6, 2, 4, 1, AK, BK>(st, p);    // one group, 8 consumer warps of 64x32
    else launch_tma<128, 64, 4, 2, 2, 2, AK, BK>(st, p);              // two groups, 4 consumer warps of 64x32 each
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
