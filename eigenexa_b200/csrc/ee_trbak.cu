// ee_trbak.cu -- compact-WY back-transformation  Z <- H_{n-1} ... H_1 Z  on B200.
//
// Replaces eigen_common_trbakwy / eigen_trbakwy_body / trbakwy_datacast
// (src/trbakwy4.F:77,227,655) and eigen_trbakwy_block_body(1,2) (src/trbakwy4_body.F:107-741).
// Per block of mb reflectors V = [u_i ... u_{i+mb-1}] (column u_j has j rows, zero below):
//     S = -V^T V (lower), S_jj <- S_jj/2 (0 -> 1)        (trbakwy4_body.F:206-213,302-313)
//     Z <- Z + V S^{-1} (V^T Z)                            (trbakwy4_body.F:604-608,687-725)
// Like the reference's blocked part, beta from the forward pass is not used; unlike the
// reference there is no separate unblocked "head" (trbakwy4.F:345-499): the leading
// (n-1) mod mb reflectors simply form a first, shorter block.
// All O(n^2 mb) work is FP64 tensor-core GEMM (ee_gemm.cu); on a P x Q grid the V panel is
// assembled replicated (pack of the owned pieces + one all-gather + unpack, prefetched one block
// ahead on the side stream: trbakwy4.F:508-602 prefetches two ahead) and V^T Z is summed over the
// x group, as in trbakwy4_body.F:235.  With a host destination, Z is processed in column chunks and
// every finished chunk leaves for the host behind the GEMMs of the next one.
#include "ee_common.cuh"
#include "ee_comm.h"
#include <chrono>

namespace ee {

namespace {

constexpr int MB_MAX = 256;   // nsm of the reference (src/eigen_devel.F:88-91)
constexpr int TB = 128;       // diagonal block handled by one tinv CTA
constexpr int KSPLIT = 64;
constexpr int SS_SPLIT_MAX = 6;

// V(g, c) = u_{i0+c}(g) for g < i0+c-(iblk-1), from the local cyclic pieces of a (zero elsewhere).
// iblk = 1: reflectors of eigen_trd (column gc has gc rows); iblk = 2: of eigen_prd (gc-1 rows),
// the nb argument of eigen_common_trbakwy (src/trbakwy4.F:77, src/eigen_sx.F:245-247).
__global__ void gather_v_kernel(const double *A, int lda, int px, int py, int x, int y, int i0, int mb, int rows,
                                int iblk, double *V, int ldv)
{
    const int c = blockIdx.y;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ldv) return;
    const int gc = i0 + c;
    double v = 0.0;
    if (c < mb && g < gc - (iblk - 1) && g < rows && (gc % py) == y && (g % px) == x) v = A[(size_t)(gc / py) * lda + g / px];
    V[(size_t)c * ldv + g] = v;
}

// multi-rank V panel: every rank packs the pieces it owns (rows g = x mod px, columns gc = y mod py of the block)
// into [lc][jl] (lc < cols_max, jl < rows_max, zero padded), one all-gather over the world, then every rank
// unpacks the P pieces into the replicated panel.  (The reference broadcasts the panel inside the process rows,
// src/trbakwy4.F:686-733; an all-reduce of a zero-padded full-length panel moves P times these bytes.)
__global__ void pack_v_kernel(const double *A, int lda, int px, int py, int x, int y, int i0, int mb, int rows, int iblk,
                              int lc0, int cols_max, int rows_max, double *out)
{
    const int lc = blockIdx.y;                 // local column index inside the block
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= rows_max) return;
    const int gc = (lc0 + lc) * py + y;        // global column
    const long long g = (long long)jl * px + x;
    double v = 0.0;
    if (gc >= i0 && gc < i0 + mb && g < gc - (iblk - 1) && g < rows) v = A[(size_t)(lc0 + lc) * lda + jl];
    out[(size_t)lc * rows_max + jl] = v;
}
__global__ void unpack_v_kernel(const double *in, int px, int py, int order_r, int i0, int mb, int rows, int cols_max,
                                int rows_max, double *V, int ldv)
{
    const int c = blockIdx.y;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ldv) return;
    double v = 0.0;
    if (c < mb && g < rows) {
        const int gc = i0 + c;
        const int xo = g % px, yo = gc % py;
        const int wr = order_r ? xo * py + yo : xo + yo * px;         // world rank of the owner
        const int lc = gc / py - cyc_count(i0, py, yo);
        v = in[((size_t)wr * cols_max + lc) * rows_max + g / px];
    }
    V[(size_t)c * ldv + g] = v;
}

// local rows of the replicated panel: Vx(jl, c) = V(jl*px + x, c)
__global__ void pick_rows_kernel(const double *V, int ldv, int px, int x, int nrl, double *Vx, int ldvx)
{
    const int c = blockIdx.y;
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= ldvx) return;
    Vx[(size_t)c * ldvx + jl] = (jl < nrl) ? V[(size_t)c * ldv + (size_t)jl * px + x] : 0.0;
}

// S = -(sum of split-K partials of V^T V), diagonal halved, 0 -> 1 (trbakwy4_body.F:206-213,302-313)
__global__ void reduce_s_kernel(const double *SMpart, int nsplit, int mb, double *S, double *T)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= mb * mb) return;
    const int i = idx % mb, k = idx / mb;
    double s = 0.0;
    for (int z = 0; z < nsplit; z++) s += SMpart[(size_t)z * mb * mb + idx];
    s = -s;
    if (i == k) s = (s == 0.0) ? 1.0 : 0.5 * s;
    S[idx] = s;
    T[idx] = 0.0;
}

// T_bb = S_bb^{-1} for the diagonal blocks (<= 128) of the lower triangular S; one CTA per block,
// thread j solves S x = e_j by forward substitution (column j of the block).
__global__ void __launch_bounds__(TB) tinv_kernel(const double *S, int lds, int mb, double *T)
{
    extern __shared__ double Sb[];  // bs x (bs+1)
    const int off = blockIdx.x * TB;
    const int bs = min(TB, mb - off);
    const int ld = bs + 1;
    for (int idx = threadIdx.x; idx < bs * bs; idx += blockDim.x) {
        int i = idx % bs, k = idx / bs;
        Sb[i * ld + k] = S[(size_t)(off + k) * lds + off + i];
    }
    __syncthreads();
    const int j = threadIdx.x;
    if (j < bs) {
        double *tc = T + (size_t)(off + j) * lds + off;  // column j of the block, L1/L2 resident
        tc[j] = 1.0 / Sb[j * ld + j];
        for (int i = j + 1; i < bs; i++) {
            double s0 = 0.0, s1 = 0.0;
            int k = j;
            for (; k + 1 < i; k += 2) {
                s0 = fma(Sb[i * ld + k], tc[k], s0);
                s1 = fma(Sb[i * ld + k + 1], tc[k + 1], s1);
            }
            if (k < i) s0 = fma(Sb[i * ld + k], tc[k], s0);
            tc[i] = -(s0 + s1) / Sb[i * ld + i];
        }
    }
}

// out = sum_z part[z]  (fixed order)
__global__ void reduce_parts_kernel(const double *part, long long stride, int nsplit, long long count, double *out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        double s = part[i];
        for (int z = 1; z < nsplit; z++) s += part[(long long)z * stride + i];
        out[i] = s;
    }
}

// split-K factor that fills the last wave of a deep-K GEMM (tiles x s work items over `slots`)
int choose_ksplit(long long tiles, int K, int slots)
{
    int best = 1; double beste = 0.0;
    for (int s = 1; s <= SS_SPLIT_MAX; s++) {
        if (s > 1 && K / s < 2048) break;
        double w = (double)tiles * s / slots;
        double e = w / ceil(w);
        if (e > beste + 0.02) { beste = e; best = s; }
    }
    return best;
}

}  // namespace

// Host destination of the back-transformed eigenvectors, set by the host-array entry point (ee_capi.cu): the
// columns of Z are independent, so Z is processed in column chunks and each finished chunk goes to the host on
// the side stream while the next one is transformed (single rank; on a grid each rank moves only 1/P of Z).
static double *g_zhost = nullptr;
static int g_ldzhost = 0;
void trbak_set_host_output(double *z_host, int ldz_host) { g_zhost = z_host; g_ldzhost = ldz_host; }

void trbak_dev(int n, int nvec, const double *a, int lda, double *z, int ldz, const double *e, int m_backward, int iblk)
{
    (void)e;
    if (iblk < 1) iblk = 1;
    Context &c = ctx();
    const Grid &g = c.g;
    cudaStream_t st = c.stream, s2 = c.stream2;
    double *zhost = g_zhost; const int ldzhost = g_ldzhost;
    g_zhost = nullptr;
    if (n <= iblk || nvec <= 0) {
        if (zhost) {   // nothing to transform: the tridiagonal eigenvectors are the result
            const int nrl0 = cyc_count(n, g.px, g.x), nvl0 = cyc_count(nvec, g.py, g.y);
            if (nrl0 > 0 && nvl0 > 0)
                EE_CUDA(cudaMemcpy2DAsync(zhost, (size_t)ldzhost * sizeof(double), z, (size_t)ldz * sizeof(double),
                                          (size_t)nrl0 * sizeof(double), nvl0, cudaMemcpyDeviceToHost, st));
        }
        return;
    }
    // m_backward is a blocking hint (reference default 128, cap nsm = 256).  The two GEMMs per
    // block run at K = mb resp. M = mb; wide blocks keep the C read-modify-write of Z hidden.
    int mb = m_backward < MB_MAX ? m_backward : MB_MAX;
    if (n >= 8192 && mb < MB_MAX) mb = MB_MAX;
    if (mb < 1) mb = 1;
    if (mb > n - iblk) mb = n - iblk;
    const int nvl = cyc_count(nvec, g.py, g.y);
    const bool multi = g.nnod > 1;
    const int P = g.nnod;
    const int ldv = (n + 15) & ~15;
    const int nrl_all = cyc_count(n, g.px, g.x);
    const int nrl_max = (n + g.px - 1) / g.px;
    const int ldvx = ((nrl_max > 0 ? nrl_max : 1) + 15) & ~15;
    auto wall = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double hw0 = wall();
    // block list: reflector columns i = iblk..n-1 ; first block takes the remainder (trbakwy4.F:292)
    struct Blk { int i0, cur, rows; };
    std::vector<Blk> blks;
    {
        int i0 = iblk;
        const int first = (n - iblk) % mb;
        while (i0 <= n - 1) {
            const int cur = (i0 == iblk && first != 0) ? first : mb;
            blks.push_back({i0, cur, i0 + cur - iblk});
            i0 += cur;
        }
    }
    const int nblk = (int)blks.size();
    // column chunks of Z (host output, single rank): finished chunks stream to the host behind the GEMMs
    int nchunk = 1;
    if (zhost && !multi && nvl >= 8192) nchunk = 4;
    {
        const char *ev = getenv("EIGENEXA_B200_TRBAK_CHUNKS");
        if (ev && atoi(ev) >= 1 && !multi) nchunk = atoi(ev);
        if (nchunk > nvl) nchunk = nvl > 0 ? nvl : 1;
    }
    const bool cacheT = nchunk > 1;
    // the V panel of block b+1 is assembled on the side stream while block b's GEMMs run (multi-rank):
    // two panel buffers, the reference keeps three and prefetches two blocks ahead (src/trbakwy4.F:508-602)
    const int nbuf = multi ? 2 : 1;
    const int cols_max = mb / g.py + 2, rows_max = ((nrl_max + 1) + 1) & ~1;
    double *V[2] = {nullptr, nullptr}, *Vx[2] = {nullptr, nullptr};
    double *pk = nullptr, *pg = nullptr;
    for (int b = 0; b < nbuf; b++) {
        V[b] = (double *)dev_alloc((size_t)ldv * mb * sizeof(double));
        Vx[b] = (g.px > 1) ? (double *)dev_alloc((size_t)ldvx * mb * sizeof(double)) : V[b];
    }
    if (multi) {
        pk = (double *)dev_alloc((size_t)cols_max * rows_max * sizeof(double));
        pg = (double *)dev_alloc((size_t)P * cols_max * rows_max * sizeof(double));
    }
    double *SMp = (double *)dev_alloc((size_t)KSPLIT * mb * mb * sizeof(double));
    double *S = (double *)dev_alloc((size_t)mb * mb * sizeof(double));
    double *T = (double *)dev_alloc((size_t)mb * mb * (cacheT ? nblk : 1) * sizeof(double));
    double *Wt = (double *)dev_alloc((size_t)mb * mb * sizeof(double));
    const int ldss = (mb + 1) & ~1;
    const int nvl_c = (nvl + nchunk - 1) / nchunk;                  // columns per chunk
    const size_t ss_elems = (size_t)ldss * (nvl_c > 0 ? nvl_c : 1);
    double *SSp = (double *)dev_alloc(ss_elems * SS_SPLIT_MAX * sizeof(double));
    double *SS = (double *)dev_alloc(ss_elems * sizeof(double));
    double *SS2 = (double *)dev_alloc(ss_elems * sizeof(double));
    EE_CUDA(cudaFuncSetAttribute(tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TB * (TB + 1) * sizeof(double))));
    cudaEvent_t evReady[2], evFree[2], evChunk;
    for (int b = 0; b < 2; b++) {
        EE_CUDA(cudaEventCreateWithFlags(&evReady[b], cudaEventDisableTiming));
        EE_CUDA(cudaEventCreateWithFlags(&evFree[b], cudaEventDisableTiming));
    }
    EE_CUDA(cudaEventCreateWithFlags(&evChunk, cudaEventDisableTiming));

    const double hw1 = wall();
    // profiling level 2: per-class device time (sync per class) -> timings[22..27]
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    double tcls[6] = {0, 0, 0, 0, 0, 0};
    if (c.profiling >= 2) { EE_CUDA(cudaEventCreate(&pe0)); EE_CUDA(cudaEventCreate(&pe1)); }
    auto pb = [&]() { if (c.profiling >= 2) EE_CUDA(cudaEventRecord(pe0, st)); };
    auto pe = [&](int cls) {
        if (c.profiling >= 2) {
            EE_CUDA(cudaEventRecord(pe1, st)); EE_CUDA(cudaEventSynchronize(pe1));
            float ms; EE_CUDA(cudaEventElapsedTime(&ms, pe0, pe1)); tcls[cls] += ms * 1e-3;
        }
    };
    // V panel of block b into buffer b % nbuf, on stream q (K15 + C12)
    auto assemble_v = [&](int b, cudaStream_t q) {
        const Blk &B = blks[b];
        const int buf = b % nbuf;
        const int nrl = cyc_count(B.rows, g.px, g.x);
        if (!multi) {
            dim3 grid((ldv + 255) / 256, B.cur);
            gather_v_kernel<<<grid, 256, 0, q>>>(a, lda, g.px, g.py, g.x, g.y, B.i0, B.cur, B.rows, iblk, V[buf], ldv);
            EE_CHECK_LAUNCH();
            return;
        }
        {
            dim3 grid((rows_max + 255) / 256, cols_max);
            pack_v_kernel<<<grid, 256, 0, q>>>(a, lda, g.px, g.py, g.x, g.y, B.i0, B.cur, B.rows, iblk,
                                               cyc_count(B.i0, g.py, g.y), cols_max, rows_max, pk);
            EE_CHECK_LAUNCH();
        }
        comm_allgather(pk, pg, (size_t)cols_max * rows_max, COMM_WORLD, q);
        {
            dim3 grid((ldv + 255) / 256, B.cur);
            unpack_v_kernel<<<grid, 256, 0, q>>>(pg, g.px, g.py, g.order == 'R' ? 1 : 0, B.i0, B.cur, B.rows, cols_max, rows_max,
                                                 V[buf], ldv);
            EE_CHECK_LAUNCH();
        }
        if (g.px > 1) {
            dim3 grid2((ldvx + 255) / 256, B.cur);
            pick_rows_kernel<<<grid2, 256, 0, q>>>(V[buf], ldv, g.px, g.x, nrl, Vx[buf], ldvx);
            EE_CHECK_LAUNCH();
        }
    };
    if (multi) {
        // side stream starts behind everything already queued on the main stream (a is final there)
        EE_CUDA(cudaEventRecord(evChunk, st));
        EE_CUDA(cudaStreamWaitEvent(s2, evChunk, 0));
        assemble_v(0, s2);
        EE_CUDA(cudaEventRecord(evReady[0], s2));
    }
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = ch * nvl_c, c1 = std::min(nvl, c0 + nvl_c);
        const int ncv = c1 - c0;
        double *zc = z + (size_t)c0 * ldz;
        for (int b = 0; b < nblk; b++) {
            const Blk &B = blks[b];
            const int cur = B.cur, rows = B.rows, buf = b % nbuf;
            const int nrl = cyc_count(rows, g.px, g.x);
            double *Tb = T + (cacheT ? (size_t)b * mb * mb : 0);
            // ---- V panel (K15) ---------------------------------------------------------------
            pb();
            if (multi) {
                if (b + 1 < nblk) {
                    // prefetch the next panel: its buffer is free once block b-1 has finished with it
                    if (b >= 1) EE_CUDA(cudaStreamWaitEvent(s2, evFree[(b + 1) % nbuf], 0));
                    assemble_v(b + 1, s2);
                    EE_CUDA(cudaEventRecord(evReady[(b + 1) % nbuf], s2));
                }
                EE_CUDA(cudaStreamWaitEvent(st, evReady[buf], 0));
            } else assemble_v(b, st);
            pe(0);
            const int ldx = (g.px > 1) ? ldvx : ldv;
            const double *Vb = V[buf], *Vxb = Vx[buf];
            pb();
            // ---- S = -V^T V from the replicated panel (split-K partials), T = S^{-1} (K12,K13) ---
            if (ch == 0 || !cacheT) {
                int ks = rows / 512; if (ks < 1) ks = 1; if (ks > KSPLIT) ks = KSPLIT;
                dgemm_ex(st, 'T', 'N', cur, cur, rows, 1.0, Vb, ldv, Vb, ldv, 0.0, SMp, cur, ks, (long long)cur * cur);
                reduce_s_kernel<<<(cur * cur + 255) / 256, 256, 0, st>>>(SMp, ks, cur, S, Tb);
                EE_CHECK_LAUNCH();
                const int nb = (cur + TB - 1) / TB;
                const int bs0 = cur < TB ? cur : TB;
                tinv_kernel<<<nb, TB, (size_t)bs0 * (bs0 + 1) * sizeof(double), st>>>(S, cur, cur, Tb);
                EE_CHECK_LAUNCH();
                if (nb == 2) {
                    // T21 = -T22 S21 T11
                    const int b1 = TB, b2 = cur - TB;
                    dgemm(st, 'N', 'N', b2, b1, b1, 1.0, S + b1, cur, Tb, cur, 0.0, Wt, b2);
                    dgemm(st, 'N', 'N', b2, b1, b2, -1.0, Tb + (size_t)b1 * cur + b1, cur, Wt, b2, 0.0, Tb + b1, cur);
                }
            }
            pe(1);
            if (ncv > 0) {
                pb();
                // ---- SS = Vx^T Z  (K12), split-K sized to fill the last wave ; x-group sum (C13) -----
                if (nrl > 0) {
                    const long long tiles = (long long)((cur + 127) / 128) * ((ncv + 63) / 64);
                    const int sk = choose_ksplit(tiles, nrl, 2 * c.sm_count);
                    if (sk == 1) dgemm(st, 'T', 'N', cur, ncv, nrl, 1.0, Vxb, ldx, zc, ldz, 0.0, SS, ldss);
                    else {
                        dgemm_ex(st, 'T', 'N', cur, ncv, nrl, 1.0, Vxb, ldx, zc, ldz, 0.0, SSp, ldss, sk, (long long)ss_elems);
                        reduce_parts_kernel<<<c.sm_count * 4, 256, 0, st>>>(SSp, (long long)ss_elems, sk, (long long)ldss * ncv, SS);
                        EE_CHECK_LAUNCH();
                    }
                } else EE_CUDA(cudaMemsetAsync(SS, 0, (size_t)ldss * ncv * sizeof(double), st));
                if (g.px > 1) comm_allreduce_sum(SS, (size_t)ldss * ncv, COMM_X, st);
                pe(2);
                // ---- SS2 = T SS ; Z += Vx SS2  (K14) ---------------------------------------------
                pb();
                dgemm(st, 'N', 'N', cur, ncv, cur, 1.0, Tb, cur, SS, ldss, 0.0, SS2, ldss);
                pe(3);
                pb();
                if (nrl > 0) dgemm(st, 'N', 'N', nrl, ncv, cur, 1.0, Vxb, ldx, SS2, ldss, 1.0, zc, ldz);
                pe(4);
            }
            if (multi) EE_CUDA(cudaEventRecord(evFree[buf], st));
        }
        if (zhost && nrl_all > 0 && ncv > 0) {
            // this chunk is final: to the host on the side stream (the main stream goes on with the next chunk)
            EE_CUDA(cudaEventRecord(evChunk, st));
            EE_CUDA(cudaStreamWaitEvent(s2, evChunk, 0));
            if (ldzhost == ldz) EE_CUDA(cudaMemcpyAsync(zhost + (size_t)c0 * ldzhost, zc, (size_t)ldz * ncv * sizeof(double), cudaMemcpyDeviceToHost, s2));
            else EE_CUDA(cudaMemcpy2DAsync(zhost + (size_t)c0 * ldzhost, (size_t)ldzhost * sizeof(double), zc, (size_t)ldz * sizeof(double),
                                           (size_t)nrl_all * sizeof(double), ncv, cudaMemcpyDeviceToHost, s2));
        }
    }
    const double hw2 = wall();
    EE_CUDA(cudaStreamSynchronize(st));
    if (multi) EE_CUDA(cudaStreamSynchronize(s2));     // (host output copies are awaited by the caller)
    const double hw3 = wall();
    c.timings[27] = hw1 - hw0; c.timings[28] = hw2 - hw1; c.timings[29] = hw3 - hw2;
    if (c.profiling >= 2) {
        for (int i = 0; i < 5; i++) c.timings[22 + i] = tcls[i];
        cudaEventDestroy(pe0); cudaEventDestroy(pe1);
    }
    for (int b = 0; b < 2; b++) { cudaEventDestroy(evReady[b]); cudaEventDestroy(evFree[b]); }
    cudaEventDestroy(evChunk);
    for (int b = 0; b < nbuf; b++) { if (g.px > 1) dev_free(Vx[b]); dev_free(V[b]); }
    if (pk) dev_free(pk);
    if (pg) dev_free(pg);
    dev_free(SMp); dev_free(S); dev_free(T); dev_free(Wt); dev_free(SSp); dev_free(SS); dev_free(SS2);
    c.timings[30] = wall() - hw3;
}

}  // namespace ee
