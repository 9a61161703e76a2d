// ee_trd.cu -- blocked Householder tri- and penta-diagonalisation on B200 (sm_100a).
//
// eigen_prd (two columns per step, src/eigen_prd.F) shares the SYMV tile kernel, the panel kernels and the
// trailing update with eigen_trd; its own kernels and their algebra are described further down.
//
// Replaces eigen_trd and its helpers (reference: src/eigen_trd.F:82-723,
// src/eigen_trd_t2.F (au: SYMV + scalars), _t4 (compute_u), _t5/_t5x (panel update),
// _t6_3 (compute_v), _t7 (panel load/restore), _t8 (init/final), src/eigen_t1.F (rank-2k)).
//
// B200-first design (not a translation):
//  * the local part of A stays in HBM in the caller's 2D cyclic layout; every vector of the
//    algorithm (current column, panel copy W, reflector panels U and V, p = A u) is kept
//    FULL LENGTH and replicated on every rank, so one column step needs exactly one
//    cross-rank exchange (of the partial p) instead of the reference's
//    redistribution + 2 all-reduces + broadcast (src/eigen_trd_t2.F:426-567).
//  * eigen_trd: ALL column steps of a panel run inside ONE persistent cooperative launch
//    (trd_panel_kernel, further down): per column a SYMV phase over the local strict upper
//    "staircase" (HBM-bound: each element feeds a column dot and a row axpy, K1 of SURVEY 2.4;
//    deterministic per-tile partials; the panel dot products U^T u, V^T u ride along as work
//    items), a p phase (sum of partials + diag*u - U s - V t, on a grid the exchange over NVLink
//    peer memory, partial u^T p) and a v phase (v = (p - alpha u)/beta, next column from the panel
//    copy, left-looking, and its norm), separated by grid barriers; the Householder scalars
//    (g, u_L, beta, alpha) are reduced redundantly by every CTA and never leave the device.
//  * the three-launch form of the same step (symv_kernel -> pvec_kernel -> vvec_kernel) is kept
//    as the debugging reference / fallback (EIGENEXA_B200_TRD_PERSIST=0, or no peer ring on a grid),
//    and eigen_prd still uses its four-launch form (next2 -> house2 -> symv2 -> pvec2).
//  * the trailing update A -= U V^T + V U^T is one FP64 tensor-core (DMMA) GEMM with K = 2m
//    on the upper staircase tiles (ee_gemm.cu).
// All reductions have a fixed order: results are bit-reproducible run to run.
#include "ee_common.cuh"
#include "ee_comm.h"
#include <chrono>

namespace ee {

namespace {

constexpr int TR = 128;     // tile rows (local)
constexpr int TC = 64;      // tile cols (local)
constexpr int SW = 4;       // max sub-tiles swept per CTA along a row strip (TrdP::sw <= SW)
constexpr int NCH = 32;     // row chunks for the panel dot products
constexpr int MAXM = 256;   // max panel width

struct TrdP {
    double *A; int lda;                 // local matrix, padded (lda % TR == 0, cols % TC == 0)
    int px, py, x, y;                   // grid
    int n, npad;                        // global size, leading dim of replicated panels
    int L;                              // reflector length = global column index i (0-based)
    int k, m0, ndone;                   // slot in panel, panel width, finished pairs (slots k+1..m0-1)
    int i_base;
    int sw;                             // sub-tiles per strip for this launch (1, 2 or 4)
    double *U, *V, *W;                  // replicated panels, npad x m
    double *ucur, *unext;               // current / next raw column (length npad)
    double *Prow, *Pcol; int ldprow, ldpcol;
    double *dots_part;                  // [NCH][2*MAXM]
    double *st;                         // [2*MAXM]  s_l = V_l^T u, t_l = U_l^T u (l = k+1+idx)
    double *pbuf;                       // p (length npad)
    double *part;                       // [2][maxblocks] partial sums (utp / norm)
    double *scal;                       // [0]=g [1]=u_n [2]=beta [3]=alpha
    unsigned int *tickets;              // [4]
    double *d_out, *e_out;
    int has_next;                       // vvec: build next column
    int first;                          // vvec: panel prologue only (no v to form)
    int piv; double upiv;               // persistent kernel: u(piv) is upiv (not yet visible in ucur), piv = -1: none
    int sv;                             // persistent kernel: tile rows per work item (column partials are summed over them)
};

// Loads of vectors that other CTAs write during the same launch (persistent panel kernel) must bypass the
// non-coherent L1 / read-only path: COH = true -> ld.global.cg.  The multi-launch kernels keep ld.global.nc.
template <bool COH>
__device__ __forceinline__ double ldv(const double *p) { return COH ? __ldcg(p) : __ldg(p); }
template <bool COH>
__device__ __forceinline__ double ldu(const TrdP &P, const double *u, long long g)
{
    if (COH && g == P.piv) return P.upiv;
    return ldv<COH>(u + g);
}

// --- staircase geometry shared by writer (symv) and reader (pvec) ------------------------
// number of local columns with global index < L
__device__ __forceinline__ int ncl_of(const TrdP &P) { return cyc_count(P.L, P.py, P.y); }
// number of active tile rows for sub-tile column t: tile rows br with some row g_row < max col g
__device__ __forceinline__ int ntile_rows(const TrdP &P, int t, int nclL)
{
    int clast = min((t + 1) * TC, nclL) - 1;         // last local col of the sub-tile within L
    if (clast < t * TC) return 0;
    long long cmax_g = (long long)clast * P.py + P.y;  // its global index
    // rows br*TR*px + x < cmax_g
    if (cmax_g <= P.x) return 0;
    return (int)((cmax_g - P.x - 1) / ((long long)TR * P.px)) + 1;
}
__device__ __forceinline__ int nstrips_of(const TrdP &P, int nclL) { return (nclL + P.sw * TC - 1) / (P.sw * TC); }
__device__ __forceinline__ int strip_rows(const TrdP &P, int sc, int nclL)
{
    int tlast = min((sc + 1) * P.sw, (nclL + TC - 1) / TC) - 1;
    return ntile_rows(P, tlast, nclL);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum (result valid in thread 0), fixed order
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sm)
{
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; i++) r += sm[i];
    }
    return r;
}

// ---------------------------------------------------------------------------------------
// SYMV over the local strict upper staircase + panel dot products
// ---------------------------------------------------------------------------------------
// NV = 1: one vector (eigen_trd_au).  NV = 2: the two reflectors of a column pair in ONE pass
// over the matrix (eigen_prd_au, src/eigen_prd_t2.F:153-206): half the bytes per column.
template <int NV>
struct SymvIO {
    const double *u[NV];     // replicated input vectors (entries >= P.L are ignored)
    double *prow[NV];        // row partials  [strip][row]
    double *pcol[NV];        // col partials  [tile row][col]
};

// COH (persistent kernel, NV = 1): the entries of u this strip needs -- SW*TC columns, TR rows -- are staged in
// shared memory (su) with coalesced ld.global.cg loads up front: a coherent load inside the column loop would
// be an L2 round trip per column on the critical path.
// scol != nullptr (persistent kernel): the column sums are added to scol[SW*TC] (every entry belongs to one thread)
// instead of being written to Pcol -- the caller sweeps several tile rows and writes one partial per column.
template <int NV, bool COH = false>
__device__ __forceinline__ void symv_strip(const TrdP &P, const SymvIO<NV> &io, int br, int sc, int nclL, double *smem,
                                           double *su = nullptr, double *scol = nullptr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = br * TR;
    // this thread's 4 rows
    int rloc[4] = {r0 + 2 * lane, r0 + 2 * lane + 1, r0 + 64 + 2 * lane, r0 + 64 + 2 * lane + 1};
    double ux[NV][4], acc_row[NV][4];
    // COH: the staging loads are issued here but only consumed (stored to shared memory) behind the loads of the
    // first tile, so the strip does not start with an exposed L2 round trip
    double stg[2] = {0.0, 0.0};
    bool staged = !COH;
    if (COH) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int idx = threadIdx.x + q * 256;
            if (idx < SW * TC + TR) {
                long long g;
                if (idx < SW * TC) g = (long long)(sc * P.sw * TC + idx) * P.py + P.y;    // column idx of the strip
                else g = (long long)(r0 + idx - SW * TC) * P.px + P.x;                     // row of the tile row
                stg[q] = (g < P.L) ? ldu<true>(P, io.u[0], g) : 0.0;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        long long g = (long long)rloc[q] * P.px + P.x;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            ux[v][q] = (!COH && g < P.L) ? ldu<COH>(P, io.u[v], g) : 0.0;
            acc_row[v][q] = 0.0;
        }
    }
    const long long rmax_g = (long long)(r0 + TR - 1) * P.px + P.x;
    for (int st = 0; st < P.sw; st++) {
        const int t = sc * P.sw + st;
        const int c0 = t * TC;
        if (c0 >= nclL) break;
        if (br >= ntile_rows(P, t, nclL)) continue;
        const int cw = c0 + 8 * warp;
        const double *__restrict__ base = P.A + (size_t)cw * P.lda + r0 + 2 * lane;
        double2 v0[8], v1[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            v0[c] = __ldcs(reinterpret_cast<const double2 *>(base + (size_t)c * P.lda));
            v1[c] = __ldcs(reinterpret_cast<const double2 *>(base + (size_t)c * P.lda + 64));
        }
        if (COH && !staged) {
            su[threadIdx.x] = stg[0];
            if (threadIdx.x + 256 < SW * TC + TR) su[threadIdx.x + 256] = stg[1];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; q++) ux[0][q] = su[SW * TC + rloc[q] - r0];
            staged = true;
        }
        const long long cmin_g = (long long)c0 * P.py + P.y;
        const long long cmax_g = (long long)(c0 + TC - 1) * P.py + P.y;
        const bool interior = (rmax_g < cmin_g) && (cmax_g < P.L);
        if (!interior) {
            long long gr[4];
#pragma unroll
            for (int q = 0; q < 4; q++) gr[q] = (long long)rloc[q] * P.px + P.x;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                long long gc = (long long)(cw + c) * P.py + P.y;
                bool cin = gc < P.L;
                if (!(cin && gr[0] < gc)) v0[c].x = 0.0;
                if (!(cin && gr[1] < gc)) v0[c].y = 0.0;
                if (!(cin && gr[2] < gc)) v1[c].x = 0.0;
                if (!(cin && gr[3] < gc)) v1[c].y = 0.0;
            }
        }
#pragma unroll
        for (int v = 0; v < NV; v++) {
            double acc_col[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const long long g = (long long)(cw + c) * P.py + P.y;
                const double uy = COH ? su[st * TC + 8 * warp + c] : ((g < P.L) ? ldu<COH>(P, io.u[v], g) : 0.0);
                acc_row[v][0] = fma(v0[c].x, uy, acc_row[v][0]);
                acc_row[v][1] = fma(v0[c].y, uy, acc_row[v][1]);
                acc_row[v][2] = fma(v1[c].x, uy, acc_row[v][2]);
                acc_row[v][3] = fma(v1[c].y, uy, acc_row[v][3]);
                double s = v0[c].x * ux[v][0];
                s = fma(v0[c].y, ux[v][1], s);
                s = fma(v1[c].x, ux[v][2], s);
                s = fma(v1[c].y, ux[v][3], s);
                acc_col[c] = s;
            }
            // reduce-scatter the 8 column sums over the 32 lanes (9 exchanges instead of 40)
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            double h[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double mine = b4 ? acc_col[q + 4] : acc_col[q];
                double other = b4 ? acc_col[q] : acc_col[q + 4];
                h[q] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
            }
            double h2[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                double mine = b3 ? h[q + 2] : h[q];
                double other = b3 ? h[q] : h[q + 2];
                h2[q] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
            }
            double mine = b2 ? h2[1] : h2[0];
            double other = b2 ? h2[0] : h2[1];
            double r = mine + __shfl_xor_sync(0xffffffffu, other, 4);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            if ((lane & 3) == 0) {
                int c = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
                if (COH && scol) scol[st * TC + 8 * warp + c] += r;
                else io.pcol[v][(size_t)br * P.ldpcol + cw + c] = r;
            }
        }
    }
    // cross-warp reduction of the row sums
    double *sm = smem;  // [8][TR]
#pragma unroll
    for (int v = 0; v < NV; v++) {
        __syncthreads();
        sm[warp * TR + 2 * lane] = acc_row[v][0];
        sm[warp * TR + 2 * lane + 1] = acc_row[v][1];
        sm[warp * TR + 64 + 2 * lane] = acc_row[v][2];
        sm[warp * TR + 64 + 2 * lane + 1] = acc_row[v][3];
        __syncthreads();
        if (threadIdx.x < TR) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) s += sm[w * TR + threadIdx.x];
            io.prow[v][(size_t)sc * P.ldprow + r0 + threadIdx.x] = s;
        }
    }
}

// chunk of the panel dot products  s_l = V_l^T u, t_l = U_l^T u  (finished slots first..first+nd-1)
// for NV vectors; results st[v*2*MAXM + (0..nd-1 | nd..2nd-1)]
template <int NV, bool COH = false>
__device__ void dots_chunk(const TrdP &P, const double *const *uvec, int first_slot, int ch)
{
    const int nd = P.ndone;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
    const int rows_per = ((P.L + NCH - 1) / NCH + 31) & ~31;
    const int j0 = ch * rows_per, j1 = min(P.L, j0 + rows_per);
    // each warp takes columns l = warp, warp+8, ... of the 2*nd vectors
    for (int c = warp; c < 2 * nd; c += 8) {
        const double *col = (c < nd) ? (P.V + (size_t)(first_slot + c) * P.npad)
                                     : (P.U + (size_t)(first_slot + c - nd) * P.npad);
        double s[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) s[v] = 0.0;
#pragma unroll 4
        for (int j = j0 + lane; j < j1; j += 32) {
            const double cj = ldv<COH>(col + j);
#pragma unroll
            for (int v = 0; v < NV; v++) s[v] = fma(cj, ldu<COH>(P, uvec[v], j), s[v]);
        }
#pragma unroll
        for (int v = 0; v < NV; v++) {
            double t = warp_sum(s[v]);
            if (lane == 0) P.dots_part[((size_t)ch * NV + v) * 2 * MAXM + c] = t;
        }
    }
    if (COH) return;   // persistent kernel: every CTA sums the NCH partials itself after the grid barrier
    // last chunk CTA reduces the partials in fixed order
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&P.tickets[0], 1u);
    __syncthreads();
    if (s_last == NCH - 1) {
        __threadfence();
        for (int c = threadIdx.x; c < 2 * nd * NV; c += blockDim.x) {
            const int v = c / (2 * nd), cc = c - v * 2 * nd;
            double t = 0.0;
            for (int q = 0; q < NCH; q++) t += __ldcg(P.dots_part + ((size_t)q * NV + v) * 2 * MAXM + cc);
            P.st[(size_t)v * 2 * MAXM + cc] = t;
        }
        if (threadIdx.x == 0) P.tickets[0] = 0u;
    }
}

// fold the triangle: CTA column bx sweeps strip (nsc-1-bx) and then strip bx, so that no CTA is empty
__device__ __forceinline__ bool fold_triangle(const TrdP &P, int bid, int gx, int nclL, int &sc, int &br)
{
    const int nsc = nstrips_of(P, nclL);
    const int bx = bid % gx, by = bid / gx;
    const int sc1 = nsc - 1 - bx, sc2 = bx;
    const int n1 = strip_rows(P, sc1, nclL);
    if (by < n1) { sc = sc1; br = by; return true; }
    if (sc2 == sc1) return false;
    br = by - n1; sc = sc2;
    return br < strip_rows(P, sc2, nclL);
}

// same with work items of P.sv tile rows: item (sc, bg) covers the tile rows bg*sv .. bg*sv+sv-1 of strip sc
__device__ __forceinline__ int strip_groups(const TrdP &P, int sc, int nclL) { return (strip_rows(P, sc, nclL) + P.sv - 1) / P.sv; }
__device__ __forceinline__ bool fold_triangle_g(const TrdP &P, int bid, int gx, int nclL, int &sc, int &bg)
{
    const int nsc = nstrips_of(P, nclL);
    const int bx = bid % gx, by = bid / gx;
    const int sc1 = nsc - 1 - bx, sc2 = bx;
    const int n1 = strip_groups(P, sc1, nclL);
    if (by < n1) { sc = sc1; bg = by; return true; }
    if (sc2 == sc1) return false;
    bg = by - n1; sc = sc2;
    return bg < strip_groups(P, sc2, nclL);
}

__global__ void __launch_bounds__(256, 2) symv_kernel(TrdP P, int gx, int ntile_blocks)
{
    __shared__ double smem[8 * TR];
    const int bid = blockIdx.x;
    if (bid >= ntile_blocks) {
        const double *uv[1] = {P.ucur};
        if (P.ndone > 0) dots_chunk<1>(P, uv, P.k + 1, bid - ntile_blocks);
        return;
    }
    const int nclL = ncl_of(P);
    int sc, br;
    if (!fold_triangle(P, bid, gx, nclL, sc, br)) return;
    SymvIO<1> io;
    io.u[0] = P.ucur; io.prow[0] = P.Prow; io.pcol[0] = P.Pcol;
    symv_strip<1>(P, io, br, sc, nclL, smem);
}

// ---------------------------------------------------------------------------------------
// p = A u (from partials) - U s - V t ; partial u^T p ; last CTA: alpha
// MODE 0: fused (single rank)   MODE 1: partial only (write p_partial)   MODE 2: post-allreduce
// Thread layout: 32 consecutive rows x 8 slices; a row's partial sums / panel corrections
// are split over the 8 slices (independent loads in flight) and combined in fixed order.
// ---------------------------------------------------------------------------------------
constexpr int VR = 32;   // rows per CTA in the vector kernels
constexpr int VS = 8;    // slices per row

// MODE 3 / 4 are MODE 1 / 2 with the cross-rank sum done over NVLink peer memory instead of an
// NCCL call between the launches: 3 stores the partial into slot[parity][rank] of EVERY rank and
// publishes an epoch flag; 4 waits for all P flags and sums the slots in rank order.
template <int MODE>
__global__ void __launch_bounds__(VR * VS) pvec_kernel(TrdP P, PeerView pv, unsigned long long epoch)
{
    __shared__ double s_st[2 * MAXM];
    __shared__ int s_nbr[1024];
    __shared__ double s_acc[VS][VR];
    __shared__ double s_red[VR * VS / 32];
    __shared__ unsigned int s_last;
    const int nd = P.ndone;
    const int nclL = ncl_of(P);
    const int nsc = nstrips_of(P, nclL);
    constexpr bool PARTIAL = (MODE == 0 || MODE == 1 || MODE == 3);   // sums the tile partials
    constexpr bool FINISH = (MODE == 0 || MODE == 2 || MODE == 4);    // corrections + u^T p
    const int par = (int)(epoch & 1ull);
    if (MODE == 4 && threadIdx.x == 0) {
        const volatile unsigned long long *fl = pv.flags[pv.r] + (size_t)par * pv.P;
        for (int q = 0; q < pv.P; q++) {
            unsigned long long spins = 0;
            while (fl[q] < epoch) {
                __nanosleep(64);
                if (++spins > (1ull << 27)) { *pv.err = 1; break; }   // ~10 s: peer never arrived
            }
        }
        __threadfence_system();
    }
    if (PARTIAL) {
        for (int s = threadIdx.x; s < nsc; s += blockDim.x) s_nbr[s] = strip_rows(P, s, nclL);
    }
    if (FINISH) {
        for (int c = threadIdx.x; c < 2 * nd; c += blockDim.x) s_st[c] = P.st[c];
    }
    __syncthreads();
    const int r = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int g = blockIdx.x * VR + r;
    double acc = 0.0;
    if (g < P.L) {
        if (PARTIAL) {
            const bool rown = (g % P.px) == P.x, coln = (g % P.py) == P.y;
            if (rown) {
                const int jl = g / P.px, br = jl / TR;
                for (int s = nsc - 1 - sl; s >= 0 && s_nbr[s] > br; s -= VS) acc += __ldcs(P.Prow + (size_t)s * P.ldprow + jl);
            }
            if (coln) {
                const int il = g / P.py;
                const int nb = ntile_rows(P, il / TC, nclL);
                for (int b = sl; b < nb; b += VS) acc += __ldcs(P.Pcol + (size_t)b * P.ldpcol + il);
                if (rown && sl == 0) acc = fma(P.A[(size_t)il * P.lda + g / P.px], P.ucur[g], acc);
            }
        } else if (sl == 0) {
            if (MODE == 4) {
                const double *base = pv.slots[pv.r] + (size_t)par * pv.P * pv.slot_doubles + g;
                for (int q = 0; q < pv.P; q++) acc += __ldcg(base + (size_t)q * pv.slot_doubles);
            } else acc = P.pbuf[g];
        }
        if (FINISH) {
            // corrections with the finished pairs of this panel
            for (int l = sl; l < nd; l += VS) {
                const size_t off = (size_t)(P.k + 1 + l) * P.npad + g;
                acc = fma(-__ldg(P.U + off), s_st[l], acc);
                acc = fma(-__ldg(P.V + off), s_st[nd + l], acc);
            }
        }
    }
    s_acc[sl][r] = acc;
    __syncthreads();
    double up = 0.0;
    if (sl == 0) {
        double p = 0.0;
#pragma unroll
        for (int q = 0; q < VS; q++) p += s_acc[q][r];
        if (g < P.L) {
            if (MODE == 3) {
                const size_t off = ((size_t)par * pv.P + pv.r) * pv.slot_doubles + g;
                for (int q = 0; q < pv.P; q++) pv.slots[q][off] = p;   // NVLink stores to every rank
            } else P.pbuf[g] = p;
            if (FINISH) up = P.ucur[g] * p;
        }
        up = warp_sum(up);
    }
    if (MODE == 1) return;
    if (MODE == 3) {
        // one system-scope release per CTA (bar.sync orders the CTA's peer stores before it; the fence is cumulative)
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence_system(); s_last = atomicAdd(&P.tickets[3], 1u); }
        __syncthreads();
        if (s_last == gridDim.x - 1 && threadIdx.x == 0) {
            __threadfence_system();
            for (int q = 0; q < pv.P; q++) {
                volatile unsigned long long *fl = pv.flags[q] + (size_t)par * pv.P + pv.r;
                *fl = epoch;
            }
            P.tickets[3] = 0u;
        }
        return;
    }
    if (threadIdx.x == 0) {
        P.part[blockIdx.x] = up;
        __threadfence();
        s_last = atomicAdd(&P.tickets[1], 1u);
    }
    __syncthreads();
    if (s_last == gridDim.x - 1) {
        __threadfence();
        double s = 0.0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) s += __ldcg(P.part + q);
        // fixed order: per-thread strided sums then block_sum (deterministic for a given grid)
        double tot = block_sum<VR * VS>(s, s_red);
        if (threadIdx.x == 0) {
            double beta = P.scal[2];
            P.scal[3] = tot / (2.0 * beta);  // alpha = u^T p / (2 beta)   (trd_t6_3.F:255-262)
            P.tickets[1] = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// v = (p - alpha u)/beta ; next raw column from the panel copy (left-looking) ; its norm;
// last CTA: Householder scalars of the next column (trd_t2.F:574-614)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VR * VS) vvec_kernel(TrdP P)
{
    __shared__ double s_ur[MAXM], s_vr[MAXM];  // row c of U and V for slots lo..m0-1
    __shared__ double s_acc[VS][VR];
    __shared__ double s_red[VR * VS / 32];
    __shared__ unsigned int s_last;
    const int L = P.L, k = P.k;
    const int c = L - 1;  // next column (global), also the last row of u
    double alpha = 0.0, beta = 1.0;
    int lo;               // pairs applied to the next column: slots lo..m0-1
    if (!P.first) { alpha = P.scal[3]; beta = P.scal[2]; lo = k; }
    else lo = k + 1;      // panel prologue: no pairs yet (k = m0 - 1 -> nl = 0)
    const int nl = P.m0 - lo;
    if (P.has_next) {
        for (int l = threadIdx.x; l < nl; l += blockDim.x) {
            int slot = lo + l;
            double ur, vr;
            if (!P.first && slot == k) {
                ur = P.ucur[c];
                vr = (P.pbuf[c] - alpha * ur) / beta;
            } else {
                ur = P.U[(size_t)slot * P.npad + c];
                vr = P.V[(size_t)slot * P.npad + c];
            }
            s_ur[l] = ur; s_vr[l] = vr;
        }
    }
    __syncthreads();
    const int r = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int g = blockIdx.x * VR + r;
    const int kn = P.first ? k : k - 1;  // slot of the next column
    const int top = P.first ? L : c;     // its diagonal row (= its reflector length)
    double acc = 0.0;
    if (g < L || (P.first && g <= L)) {
        double ug = 0.0, vg = 0.0;
        if (!P.first) {
            ug = P.ucur[g];
            vg = (P.pbuf[g] - alpha * ug) / beta;
            if (sl == 0) {
                P.U[(size_t)k * P.npad + g] = ug;
                P.V[(size_t)k * P.npad + g] = vg;
            }
        }
        if (P.has_next && g <= top) {
            if (sl == 0) acc = P.W[(size_t)kn * P.npad + g];
            for (int l = sl; l < nl; l += VS) {
                const int slot = lo + l;
                double uj, vj;
                if (!P.first && slot == k) { uj = ug; vj = vg; }
                else { uj = __ldg(P.U + (size_t)slot * P.npad + g); vj = __ldg(P.V + (size_t)slot * P.npad + g); }
                acc = fma(-uj, s_vr[l], acc);
                acc = fma(-vj, s_ur[l], acc);
            }
        }
    }
    if (!P.has_next) return;
    s_acc[sl][r] = acc;
    __syncthreads();
    double nrm = 0.0;
    if (sl == 0) {
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < VS; q++) a += s_acc[q][r];
        if (g <= top) {
            P.unext[g] = a;
            if (g < top) nrm = a * a;
        }
        nrm = warp_sum(nrm);
    }
    if (threadIdx.x == 0) {
        P.part[blockIdx.x] = nrm;
        __threadfence();
        s_last = atomicAdd(&P.tickets[2], 1u);
    }
    __syncthreads();
    if (s_last == gridDim.x - 1) {
        __threadfence();
        double s = 0.0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) s += __ldcg(P.part + q);
        double anorm2 = block_sum<VR * VS>(s, s_red);
        if (threadIdx.x == 0) {
            double a_n = __ldcg(P.unext + top - 1);
            double dia = __ldcg(P.unext + top);
            double g_n, u_n, bt;
            if (anorm2 != 0.0) {
                double nr = sqrt(anorm2);
                g_n = -copysign(nr, a_n);
                u_n = a_n - g_n;
                bt = -u_n * g_n;
            } else { g_n = 0.0; u_n = 0.0; bt = 1.0; }
            P.scal[0] = g_n; P.scal[1] = u_n; P.scal[2] = bt;
            P.unext[top - 1] = u_n;
            P.e_out[top] = g_n;
            P.d_out[top] = dia;
            P.tickets[2] = 0u;
        }
    }
}

// =========================================================================================
// Persistent panel kernel: ALL column steps of one panel in ONE cooperative launch.
//
// The three launches per column (symv -> pvec -> vvec) cost ~15-20 us of launch gaps and kernel tails per column
// on one GPU and ~45 us on a grid (plus the 11 us launch floor of a small SYMV): with 50000 dependent columns that is
// the strong-scaling limiter.  Here the CTAs stay resident (2 per SM) and the kernel boundaries become grid
// barriers on a monotonic counter; the Householder scalars are reduced redundantly, in the same fixed order, by
// every CTA, so no CTA ever waits for a "last block" to publish them.
//   per column:  [S] SYMV tiles + panel dot products, handed out by a ticket counter (dynamic balance)
//                --- grid barrier ---
//                [P] p = sum of tile partials (+ diag) ; on a grid: partial -> every rank's peer slot,
//                    --- grid barrier, epoch flag to every rank, wait for the P flags ---  p = sum of the P slots
//                    p -= U s + V t ; partial u^T p per CTA
//                --- grid barrier ---  alpha
//                [V] v = (p - alpha u)/beta -> panel ; next column from the panel copy (left-looking) ; norm partial
//                --- grid barrier ---  g, u_L, beta of the next column (u_L is substituted on the fly in [S], every
//                    other phase sees it through memory)
// Every vector another CTA writes during the launch is read with ld.global.cg (L1 is not coherent); only the
// matrix itself, which is constant for the whole panel, keeps the streaming loads.
// Every spin has a watchdog (__trap): a protocol error aborts the launch instead of hanging the GPU.
// =========================================================================================
struct PanelCtl {
    unsigned long long *bar;     // grid barrier counter, zeroed by the host before the launch
    unsigned long long *work;    // work ticket counter, zeroed by the host before the launch
    double *partA, *partB;       // [gridDim.x] per-CTA partial sums (u^T p / next-column norm)
    double *tacc;                // [8] accumulated nanoseconds (CTA 0): 0 SYMV phase, 1 p phase, 2 v phase; inside the p phase
                                 //     on a grid: 3 partial + peer stores, 4 barrier, 5 flag exchange, 6 sum + corrections; or nullptr
    double *pivraw;              // [2] raw pivot / diagonal element of the next column (written before the barrier)
    int k_stop;                  // last slot to process (2 for the first panel, else 0)
    unsigned long long epoch0;   // peer epoch of the first column of this launch (multi-rank)
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// all CTAs of the (co-resident) grid; target = number of arrivals that completes this barrier.
// SYS: the CTA's earlier stores include NVLink stores to peer memory -- thread 0 releases them at system scope.
// One fence per CTA: bar.sync orders the CTA's stores before thread 0's fence and the fence is cumulative; a
// fence.sc.sys executed by every thread serialises inside the SM and costs ~30 us per column.
template <bool SYS = false>
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        if (SYS) __threadfence_system(); else __threadfence();
        atomicAdd(bar, 1ull);
        unsigned int spins = 0;
        while (ld_acquire_gpu(bar) < target) {
            if (++spins > (1u << 27)) __trap();      // seconds: a CTA never arrived
        }
        __threadfence();
    }
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ int block_max_int(int v, int *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = sm[0];
#pragma unroll
    for (int i = 1; i < NT / 32; i++) r = max(r, sm[i]);
    return r;       // valid in every thread
}
// sum of per-CTA partials, identical in every thread of every CTA (fixed order)
template <int NT>
__device__ __forceinline__ double grid_sum(const double *part, int nparts, double *sm)
{
    double s = 0.0;
    for (int q = threadIdx.x; q < nparts; q += NT) s += __ldcg(part + q);
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < NT / 32; i++) r += sm[i];
    return r;
}

template <bool MULTI>
__global__ void __launch_bounds__(256, 2) trd_panel_kernel(TrdP P, PeerView pv, PanelCtl C)
{
    __shared__ double smem[8 * TR];          // SYMV row-sum exchange; reused by the vector phases
    __shared__ double s_st[2 * MAXM];
    __shared__ int s_nbr[1024];
    __shared__ double s_ur[MAXM], s_vr[MAXM];
    __shared__ double s_u[SW * TC + TR];     // entries of u of the strip in flight
    __shared__ double s_col[SW * TC];        // column sums of the work item in flight (summed over its tile rows)
    __shared__ double s_red[8];
    __shared__ int s_redi[8];
    __shared__ long long s_item;
    double(*s_acc)[VR] = reinterpret_cast<double(*)[VR]>(smem);   // [VS][VR]
    const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
    const int m0 = P.m0, i_base = P.i_base;
    unsigned long long nbar = 0;             // barriers passed so far
    unsigned long long work_base = 0;        // tickets consumed by the finished columns
    const bool timing = (C.tacc != nullptr) && bid == 0 && tid == 0;
    unsigned long long t_mark = 0, t_sub = 0;
    double *ucur = P.ucur, *unext = P.unext;
    double sc_g = 0.0, sc_un = 0.0, sc_beta = 1.0;   // Householder scalars of the current column (replicated per CTA)
    const int r = tid & 31, sl = tid >> 5;

    // ---- next column (slot kn, diagonal row top) from the panel copy; returns its scalars ------------------------
    // first: panel prologue (no finished pair to apply, no v to form)
    auto v_phase = [&](int k, int L, bool first, bool has_next, double alpha, double beta, double *u_c, double *u_n,
                       double &g_out, double &un_out, double &beta_out) {
        const int c = L - 1;
        const int lo = first ? k + 1 : k;
        const int nl = m0 - lo;
        if (has_next) {
            for (int l = tid; l < nl; l += 256) {
                const int slot = lo + l;
                double ur, vr;
                if (!first && slot == k) {
                    ur = __ldcg(u_c + c);
                    vr = (__ldcg(P.pbuf + c) - alpha * ur) / beta;
                } else {
                    ur = __ldcg(P.U + (size_t)slot * P.npad + c);
                    vr = __ldcg(P.V + (size_t)slot * P.npad + c);
                }
                s_ur[l] = ur; s_vr[l] = vr;
            }
        }
        __syncthreads();
        const int kn = first ? k : k - 1;
        const int top = first ? L : c;
        const int nrows = first ? L + 1 : L;
        double nrm_cta = 0.0;
        // loads + FMAs of one row block (no block barrier inside): the NEXT block's are issued before the current block is
        // reduced, so two blocks of L2 latency overlap
        struct VRow { double ug, vg, acc; };
        auto v_rows = [&](int rb) -> VRow {
            VRow o = {0.0, 0.0, 0.0};
            const int g = rb * VR + r;
            if (rb * VR < nrows && g < nrows) {
                if (!first) {
                    o.ug = __ldcg(u_c + g);
                    o.vg = (__ldcg(P.pbuf + g) - alpha * o.ug) / beta;
                }
                if (has_next && g <= top) {
                    if (sl == 0) o.acc = __ldcg(P.W + (size_t)kn * P.npad + g);
                    for (int l = sl; l < nl; l += VS) {
                        const int slot = lo + l;
                        double uj, vj;
                        if (!first && slot == k) { uj = o.ug; vj = o.vg; }
                        else { uj = __ldcg(P.U + (size_t)slot * P.npad + g); vj = __ldcg(P.V + (size_t)slot * P.npad + g); }
                        o.acc = fma(-uj, s_vr[l], o.acc);
                        o.acc = fma(-vj, s_ur[l], o.acc);
                    }
                }
            }
            return o;
        };
        VRow cur = v_rows(bid);
        for (int rb = bid; rb * VR < nrows; rb += G) {
            const int g = rb * VR + r;
            const VRow nx = v_rows(rb + G);
            if (!first && sl == 0 && g < nrows) {
                P.U[(size_t)k * P.npad + g] = cur.ug;
                P.V[(size_t)k * P.npad + g] = cur.vg;
            }
            if (has_next) {
                __syncthreads();
                s_acc[sl][r] = cur.acc;
                __syncthreads();
                if (sl == 0) {
                    double a = 0.0;
#pragma unroll
                    for (int q = 0; q < VS; q++) a += s_acc[q][r];
                    double nrm = 0.0;
                    if (g <= top) {
                        u_n[g] = a;
                        if (g < top) nrm = a * a;
                        if (g == top - 1) C.pivraw[0] = a;     // raw pivot element: every CTA reads it after the barrier
                        if (g == top) C.pivraw[1] = a;         // diagonal element
                    }
                    nrm = warp_sum(nrm);
                    nrm_cta += nrm;      // lane-uniform after warp_sum
                }
            }
            cur = nx;
        }
        if (!has_next) return;
        if (tid == 0) C.partB[bid] = nrm_cta;
        grid_barrier(C.bar, (++nbar) * (unsigned long long)G);
        const double anorm2 = grid_sum<256>(C.partB, G, s_red);
        // (not from u_n[top-1]: CTA 0 overwrites that entry with u_piv below while other CTAs may still be here)
        const double a_n = __ldcg(C.pivraw);
        double g_n, u_piv, bt;
        if (anorm2 != 0.0) {
            const double nr = sqrt(anorm2);
            g_n = -copysign(nr, a_n);
            u_piv = a_n - g_n;
            bt = -u_piv * g_n;
        } else { g_n = 0.0; u_piv = 0.0; bt = 1.0; }
        if (bid == 0 && tid == 0) {
            const double dia = __ldcg(C.pivraw + 1);
            u_n[top - 1] = u_piv;        // visible to the phases behind the next grid barrier; [S] substitutes it
            P.e_out[top] = g_n;
            P.d_out[top] = dia;
        }
        g_out = g_n; un_out = u_piv; beta_out = bt;
    };

    // ---- panel prologue: first column of the panel is the raw panel copy ---------------------------------------
    {
        const int k = m0 - 1, L = i_base + m0 - 1;
        v_phase(k, L, true, true, 0.0, 1.0, ucur, ucur, sc_g, sc_un, sc_beta);
    }

    for (int k = m0 - 1; k >= C.k_stop; k--) {
        const int L = i_base + k;
        const int ndone = m0 - 1 - k;
        const bool has_next = (k - 1 >= C.k_stop);
        TrdP Q = P;
        Q.k = k; Q.L = L; Q.ndone = ndone; Q.ucur = ucur; Q.unext = unext;
        Q.piv = L - 1; Q.upiv = sc_un;
        const int nclL = cyc_count(L, P.py, P.y);
        // work item = sw tiles along a row strip x sv tile rows.  Long items amortise the row-sum exchange and
        // shrink the partials the p phase has to sum (at L = 50000 on one GPU: 230 MB per column with 1-tile-row
        // items); short ones keep >= ~6 items per CTA for the ticket scheduler when the LOCAL trailing matrix is small
        const long long ntl = ((long long)cyc_count(L, P.px, P.x) / TR + 1) * ((long long)nclL / TC + 1) / 2;   // ~ local tiles
        const long long per = ntl / (6LL * G);
        const int sw = per >= 4 ? 4 : per >= 2 ? 2 : 1;
        const int sv = per >= 16 ? 4 : per >= 8 ? 2 : 1;
        Q.sw = sw;
        Q.sv = sv;
        const int nsc = (nclL + sw * TC - 1) / (sw * TC);
        const int gx = (nsc + 1) / 2;
        // groups of the biggest strip + groups of the smallest (fold_triangle_g), max over the CTA columns
        int gy = 0;
        for (int bx = tid; bx < gx; bx += 256) {
            const int s1 = nsc - 1 - bx, s2 = bx;
            int rr = strip_groups(Q, s1, nclL) + (s2 != s1 ? strip_groups(Q, s2, nclL) : 0);
            gy = max(gy, rr);
        }
        gy = block_max_int<256>(gy, s_redi);
        for (int s_ = tid; s_ < nsc && s_ < 1024; s_ += 256) s_nbr[s_] = strip_rows(Q, s_, nclL);
        const long long ntile = (long long)gx * gy;
        const long long ndots = ndone > 0 ? NCH : 0;
        const long long nitems = ntile + ndots;
        if (timing) t_mark = globaltimer_ns();
        // ================= [S] SYMV tiles + panel dot products =================================================
        {
            SymvIO<1> io;
            io.u[0] = ucur; io.prow[0] = P.Prow; io.pcol[0] = P.Pcol;
            const double *uv[1] = {ucur};
            long long next = 0;
            if (tid == 0) next = (long long)(atomicAdd(C.work, 1ull) - work_base);
            for (;;) {
                if (tid == 0) s_item = next;
                __syncthreads();
                const long long item = s_item;
                if (item >= nitems) break;
                if (tid == 0) next = (long long)(atomicAdd(C.work, 1ull) - work_base);   // in flight during the strip
                // the panel dot products come FIRST in ticket order: they are latency-bound (one CTA per row chunk)
                // and must run next to the tiles, not behind them
                if (item >= ndots) {
                    int sc, bg;
                    if (fold_triangle_g(Q, (int)(item - ndots), gx, nclL, sc, bg)) {
                        // this thread's column-sum slots: (lane & 3) == 0 owns column 8 warp + c of every tile
                        const int lane_ = tid & 31, warp_ = tid >> 5;
                        const int cown = ((lane_ & 16) ? 4 : 0) + ((lane_ & 8) ? 2 : 0) + ((lane_ & 4) ? 1 : 0);
                        if ((lane_ & 3) == 0)
                            for (int t_ = 0; t_ < sw; t_++) s_col[t_ * TC + 8 * warp_ + cown] = 0.0;
                        const int br_end = min(bg * Q.sv + Q.sv, strip_rows(Q, sc, nclL));
                        for (int br = bg * Q.sv; br < br_end; br++) symv_strip<1, true>(Q, io, br, sc, nclL, smem, s_u, s_col);
                        if ((lane_ & 3) == 0)
                            for (int t_ = 0; t_ < sw; t_++)
                                P.Pcol[(size_t)bg * P.ldpcol + (size_t)(sc * sw + t_) * TC + 8 * warp_ + cown] = s_col[t_ * TC + 8 * warp_ + cown];
                    }
                } else {
                    dots_chunk<1, true>(Q, uv, k + 1, (int)item);
                }
                __syncthreads();
            }
            work_base += (unsigned long long)nitems + (unsigned long long)G;   // every CTA draws exactly one ticket too many
        }
        grid_barrier(C.bar, (++nbar) * (unsigned long long)G);
        if (timing) { const unsigned long long t = globaltimer_ns(); C.tacc[0] += (double)(t - t_mark); t_mark = t; }
        // ================= [P] p, alpha ============================================================================
        for (int cc = tid; cc < 2 * ndone; cc += 256) {
            double t = 0.0;
            for (int q = 0; q < NCH; q++) t += __ldcg(P.dots_part + (size_t)q * 2 * MAXM + cc);
            s_st[cc] = t;
        }
        __syncthreads();
        const int par = MULTI ? (int)((C.epoch0 + (unsigned long long)(m0 - 1 - k)) & 1ull) : 0;
        const unsigned long long epoch = C.epoch0 + (unsigned long long)(m0 - 1 - k);
        double up_cta = 0.0;
        // partial sums of the tile partials of row block rb (all 256 threads) -> value in threads sl == 0
        auto partial_p = [&](int g) -> double {
            double acc = 0.0;
            if (g < L) {
                const bool rown = (g % P.px) == P.x, coln = (g % P.py) == P.y;
                if (rown) {
                    const int jl = g / P.px, brr = jl / TR;
                    // strips that reach this tile row: s_nbr is non-decreasing in s, so they are s >= s_lo
                    // (known trip count: the loads of the loop can be batched)
                    int s_lo = 0, s_hi = nsc;
                    while (s_lo < s_hi) {
                        const int mid = (s_lo + s_hi) >> 1;
                        if (s_nbr[mid] > brr) s_hi = mid; else s_lo = mid + 1;
                    }
#pragma unroll 4
                    for (int s_ = nsc - 1 - sl; s_ >= s_lo; s_ -= VS) acc += __ldcg(P.Prow + (size_t)s_ * P.ldprow + jl);
                }
                if (coln) {
                    const int il = g / P.py;
                    const int nb = (ntile_rows(Q, il / TC, nclL) + Q.sv - 1) / Q.sv;     // groups of tile rows
#pragma unroll 4
                    for (int b = sl; b < nb; b += VS) acc += __ldcg(P.Pcol + (size_t)b * P.ldpcol + il);
                    if (rown && sl == 0) acc = fma(P.A[(size_t)il * P.lda + g / P.px], ldu<true>(Q, ucur, g), acc);
                }
            }
            return acc;
        };
        auto corrections = [&](int g, double acc) -> double {
            if (g < L) {
                for (int l = sl; l < ndone; l += VS) {
                    const size_t off = (size_t)(k + 1 + l) * P.npad + g;
                    acc = fma(-__ldcg(P.U + off), s_st[l], acc);
                    acc = fma(-__ldcg(P.V + off), s_st[ndone + l], acc);
                }
            }
            return acc;
        };
        if (!MULTI) {
            double nxt_p = (bid * VR < L) ? corrections(bid * VR + r, partial_p(bid * VR + r)) : 0.0;
            for (int rb = bid; rb * VR < L; rb += G) {
                const int g = rb * VR + r;
                const double acc = nxt_p;
                // the next row block's loads are in flight while this one is reduced
                nxt_p = ((rb + G) * VR < L) ? corrections((rb + G) * VR + r, partial_p((rb + G) * VR + r)) : 0.0;
                __syncthreads();
                s_acc[sl][r] = acc;
                __syncthreads();
                if (sl == 0) {
                    double p = 0.0;
#pragma unroll
                    for (int q = 0; q < VS; q++) p += s_acc[q][r];
                    double up = 0.0;
                    if (g < L) { P.pbuf[g] = p; up = ldu<true>(Q, ucur, g) * p; }
                    up_cta += warp_sum(up);
                }
            }
        } else {
            // PULL exchange: the partial goes to this rank's OWN slot (local stores, device-scope barrier), an epoch flag
            // tells every peer it is there, and every rank then reads the P slots over NVLink.  (Pushing the partial
            // into every peer's memory needs a system-scope fence behind 400 KB of outstanding NVLink stores in
            // every CTA: 8 us per barrier + 12 us per flag round, profiles/r02_persist_notes.md.)
            double nxt_p = (bid * VR < L) ? partial_p(bid * VR + r) : 0.0;
            for (int rb = bid; rb * VR < L; rb += G) {
                const int g = rb * VR + r;
                const double acc = nxt_p;
                nxt_p = ((rb + G) * VR < L) ? partial_p((rb + G) * VR + r) : 0.0;    // in flight during the reduction
                __syncthreads();
                s_acc[sl][r] = acc;
                __syncthreads();
                if (sl == 0 && g < L) {
                    double p = 0.0;
#pragma unroll
                    for (int q = 0; q < VS; q++) p += s_acc[q][r];
                    pv.slots[pv.r][((size_t)par * pv.P + pv.r) * pv.slot_doubles + g] = p;
                }
            }
            if (timing) { const unsigned long long t = globaltimer_ns(); C.tacc[3] += (double)(t - t_mark); t_sub = t; }
            grid_barrier(C.bar, (++nbar) * (unsigned long long)G);
            if (timing) { const unsigned long long t = globaltimer_ns(); C.tacc[4] += (double)(t - t_sub); t_sub = t; }
            // one thread per peer: P release stores in a row by one thread cost ~1.5 us each (every release waits for
            // the previous store's acknowledgement), 12 us per column on 8 GPUs
            if (bid == 0 && tid < pv.P) {
                __threadfence_system();
                st_release_sys(pv.flags[tid] + (size_t)par * pv.P + pv.r, epoch);
            }
            // the panel corrections do not depend on the peers: they are formed while the flags travel (and while
            // this rank waits for a slower peer) and parked in pbuf
            double nxt_c = (bid * VR < L) ? corrections(bid * VR + r, 0.0) : 0.0;
            for (int rb = bid; rb * VR < L; rb += G) {
                const int g = rb * VR + r;
                const double acc = nxt_c;
                nxt_c = ((rb + G) * VR < L) ? corrections((rb + G) * VR + r, 0.0) : 0.0;
                __syncthreads();
                s_acc[sl][r] = acc;
                __syncthreads();
                if (sl == 0 && g < L) {
                    double cs = 0.0;
#pragma unroll
                    for (int q = 0; q < VS; q++) cs += s_acc[q][r];
                    P.pbuf[g] = cs;
                }
            }
            if (tid < pv.P) {       // one poller per peer
                const unsigned long long *fl = pv.flags[pv.r] + (size_t)par * pv.P + tid;
                unsigned int spins = 0;
                while (ld_acquire_sys(fl) < epoch) {
                    if (++spins > (1u << 27)) { *pv.err = 1; __trap(); }   // a peer never arrived
                }
                __threadfence_system();
            }
            __syncthreads();
            if (timing) { const unsigned long long t = globaltimer_ns(); C.tacc[5] += (double)(t - t_sub); t_sub = t; }
            // rank q's partial from rank q's memory; the slices take one rank each (independent NVLink loads), the loads
            // of the next row block are in flight while this one is reduced, and the fixed-order sum over the slices is
            // the same on every rank
            auto peer_part = [&](int rb) -> double {
                const int g = rb * VR + r;
                double a = 0.0;
                if (rb * VR < L && g < L)
                    for (int q = sl; q < pv.P; q += VS)
                        a += ld_relaxed_sys(pv.slots[q] + ((size_t)par * pv.P + q) * pv.slot_doubles + g);
                return a;
            };
            double a0 = peer_part(bid), a1 = peer_part(bid + G), a2 = peer_part(bid + 2 * G), a3 = peer_part(bid + 3 * G);
            for (int rb = bid; rb * VR < L; rb += G) {
                const int g = rb * VR + r;
                const double acc = a0;
                a0 = a1; a1 = a2; a2 = a3;
                a3 = peer_part(rb + 4 * G);      // four row blocks of NVLink loads in flight
                __syncthreads();
                s_acc[sl][r] = acc;
                __syncthreads();
                if (sl == 0) {
                    double p = 0.0;
#pragma unroll
                    for (int q = 0; q < VS; q++) p += s_acc[q][r];
                    double up = 0.0;
                    if (g < L) { p += P.pbuf[g]; P.pbuf[g] = p; up = ldu<true>(Q, ucur, g) * p; }
                    up_cta += warp_sum(up);
                }
            }
        }
        if (MULTI && timing) { const unsigned long long t = globaltimer_ns(); C.tacc[6] += (double)(t - t_sub); }
        if (tid == 0) C.partA[bid] = up_cta;
        grid_barrier(C.bar, (++nbar) * (unsigned long long)G);
        const double alpha = grid_sum<256>(C.partA, G, s_red) / (2.0 * sc_beta);   // u^T p / (2 beta)  (trd_t6_3.F:255-262)
        if (timing) { const unsigned long long t = globaltimer_ns(); C.tacc[1] += (double)(t - t_mark); t_mark = t; }
        // ================= [V] v, next column, its scalars =========================================================
        double g_n = 0.0, u_n = 0.0, b_n = 1.0;
        v_phase(k, L, false, has_next, alpha, sc_beta, ucur, unext, g_n, u_n, b_n);
        if (timing) { const unsigned long long t = globaltimer_ns(); C.tacc[2] += (double)(t - t_mark); t_mark = t; }
        sc_g = g_n; sc_un = u_n; sc_beta = b_n;
        double *tmp = ucur; ucur = unext; unext = tmp;
    }
    (void)sc_g;
}

// ---------------------------------------------------------------------------------------
// panel load: W(:,k) <- A(0..i_base+k, i_base+k)  (local pieces scattered to global rows)
// ---------------------------------------------------------------------------------------
__global__ void panel_load_kernel(TrdP P, int zero_first)
{
    const int k = blockIdx.y;
    const int gc = P.i_base + k;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.npad) return;
    double v = 0.0;
    if (k < P.m0 && g <= gc && (gc % P.py) == P.y && (g % P.px) == P.x) v = P.A[(size_t)(gc / P.py) * P.lda + g / P.px];
    P.W[(size_t)k * P.npad + g] = v;
    if (zero_first) { P.U[(size_t)k * P.npad + g] = 0.0; P.V[(size_t)k * P.npad + g] = 0.0; }
}

// panel restore: reflector columns back into A (rows 0..gc-1 of processed columns)
__global__ void panel_restore_kernel(TrdP P, int k_stop)
{
    const int k = blockIdx.y;
    const int gc = P.i_base + k;
    if ((gc % P.py) != P.y) return;
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    const long long g = (long long)jl * P.px + P.x;
    if (k >= k_stop) {
        if (g < gc) P.A[(size_t)(gc / P.py) * P.lda + jl] = P.U[(size_t)k * P.npad + g];
    } else {
        // first panel, columns 0 and 1: apply every pair of the panel to the panel copy
        if (g <= gc) {
            double a = P.W[(size_t)k * P.npad + g];
            for (int slot = k_stop; slot < P.m0; slot++) {
                a = fma(-P.U[(size_t)slot * P.npad + g], P.V[(size_t)slot * P.npad + gc], a);
                a = fma(-P.V[(size_t)slot * P.npad + g], P.U[(size_t)slot * P.npad + gc], a);
            }
            P.A[(size_t)(gc / P.py) * P.lda + jl] = a;
        }
    }
}

// pack syr2k operands: UVx[jl][0..m) = U(g,:), [m..2m) = V(g,:) ; VUy[il] = [V | U]
__global__ void pack_uv_kernel(TrdP P, double *UVx, int ldx, int nrl, double *VUy, int ldy, int ncl, int m)
{
    const int c = blockIdx.y;  // 0..2m-1
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const double *srcx = (c < m) ? P.U + (size_t)c * P.npad : P.V + (size_t)(c - m) * P.npad;
    const double *srcy = (c < m) ? P.V + (size_t)c * P.npad : P.U + (size_t)(c - m) * P.npad;
    if (idx < nrl) UVx[(size_t)c * ldx + idx] = srcx[(size_t)idx * P.px + P.x];
    if (idx < ncl) VUy[(size_t)c * ldy + idx] = srcy[(size_t)idx * P.py + P.y];
}

// eigen_trd_final (trd_t8.F:188-219): e(1) = -A(0,1), A(0,1) *= 2, d(0), d(1), e(0) = 0
__global__ void trd_final_kernel(TrdP P)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // on a multi-rank grid the owners differ; contributions are summed by the caller
    double t = 0.0, d0 = 0.0, d1 = 0.0;
    if (P.n >= 2 && (1 % P.py) == P.y && (0 % P.px) == P.x) {
        size_t off = (size_t)(1 / P.py) * P.lda + 0;
        t = P.A[off];
        P.A[off] = 2.0 * t;
    }
    if ((0 % P.py) == P.y && (0 % P.px) == P.x) d0 = P.A[0];
    if (P.n >= 2 && (1 % P.py) == P.y && (1 % P.px) == P.x) d1 = P.A[(size_t)(1 / P.py) * P.lda + 1 / P.px];
    P.scal[4] = -t; P.scal[5] = d0; P.scal[6] = d1;
}
__global__ void trd_final_store_kernel(TrdP P)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    P.e_out[0] = 0.0;
    P.d_out[0] = P.scal[5];
    if (P.n >= 2) { P.e_out[1] = P.scal[4]; P.d_out[1] = P.scal[6]; }
}

// zero the strictly lower part and the padding of the local matrix (trd_t8.F:84-94)
__global__ void zero_lower_kernel(double *A, int lda, int ncols, int n, int px, int py, int x, int y)
{
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= lda) return;
    for (int il = blockIdx.y; il < ncols; il += gridDim.y) {   // gridDim.y is capped at 32768 by the callers
        long long gr = (long long)jl * px + x, gc = (long long)il * py + y;
        if (gr > gc || gr >= n || gc >= n) A[(size_t)il * lda + jl] = 0.0;
    }
}

// max |a| over the upper triangle + non-finite flag (eigen_scaling.F:92-107)
__global__ void absmax_kernel(const double *A, int lda, int n, int px, int py, int x, int y, int nrl, int ncl,
                              double *out /* [0]=max, [1]=bad */)
{
    double mx = 0.0; int bad = 0;
    for (int il = blockIdx.y; il < ncl; il += gridDim.y) {
        long long gc = (long long)il * py + y;
        for (int jl = blockIdx.x * blockDim.x + threadIdx.x; jl < nrl; jl += gridDim.x * blockDim.x) {
            long long gr = (long long)jl * px + x;
            if (gr <= gc) {
                double t = A[(size_t)il * lda + jl];
                if (isfinite(t)) mx = fmax(mx, fabs(t)); else bad = 1;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        // non-negative doubles order like unsigned 64-bit integers
        atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(mx));
        if (bad) atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)__double_as_longlong(1.0));
    }
}
__global__ void scale_upper_kernel(double *A, int lda, int n, int px, int py, int x, int y, int nrl, int ncl, double s)
{
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= nrl) return;
    for (int il = blockIdx.y; il < ncl; il += gridDim.y) {
        long long gr = (long long)jl * px + x, gc = (long long)il * py + y;
        if (gr <= gc) A[(size_t)il * lda + jl] *= s;
    }
}

// =========================================================================================
// eigen_prd: reduction to penta-diagonal form, two columns per step
// (reference: src/eigen_prd.F:341-580, _t4x.F:114-353 compute_u, _t2.F:153-206 au,
//  _t6_3.F:160-458 compute_v, _t5.F local_2update, _t7.F panel load/store, _t8.F init/final).
//
// A column pair (a = column c2, b = column c2-1; L = c2-1 rows lie above the band) is reduced by
// two ordinary Householder reflectors H_a (length L, pivot row L-1) and H_b (length L-1, pivot row
// L-2) -- what the reference's Cholesky-QR + coupling matrix c produce
// (src/eigen_prd_t4x.F:262-353: u_b = (H_a x_b)(1:L-1), e(i-1,1) = (H_a x_b)(L)).  H_a is applied
// to column b explicitly (one dot product, no re-orthogonalisation needed), both reflectors are
// known before the matrix is touched, so ONE pass over the local staircase yields A u_a and A u_b
// (symv_strip<2>): 2/3 n^3 bytes instead of 4/3 n^3.  The second reflector's p is corrected
// algebraically for the first one:
//     v_a = (p_a - alpha_a u_a)/beta_a                 alpha_a = u_a^T p_a / (2 beta_a)
//     p_b' = p_b - u_a (v_a^T u_b) - v_a (u_a^T u_b)
//     v_b = (p_b' - alpha_b u_b)/beta_b                alpha_b = u_b^T p_b' / (2 beta_b)
// (u_a, v_a), (u_b, v_b) then enter the panel exactly like two tridiagonalisation steps, so the
// rank-2k update and the back-transformation are shared with eigen_trd.
// One pair = four launches:  next2 (finish previous pair, form the raw pair, H_a scalars) ->
// house2 (H_a on column b, H_b scalars) -> symv2 -> pvec2.
// =========================================================================================
struct PrdP : TrdP {
    double *xa, *xb;          // current pair: raw columns, then u_a (length L), u_b (length L-1)
    double *xa_n, *xb_n;      // next pair, written by next2_kernel
    double *pa, *pb;          // A u_a, A u_b minus the panel corrections (pb = pa + npad)
    double *Prow_b, *Pcol_b;  // tile partials of the second vector
    double *e2_out;           // second off-diagonal: e2(c) = T(c-2, c)
};
// scal[]: 0 g_a  1 beta_a  2 u_a^T x_b  3 g_b  4 beta_b  5 u_a^T u_b  6 alpha_a  7 v_a^T u_b  8 alpha_b
enum { S_GA = 0, S_BA = 1, S_SAB = 2, S_GB = 3, S_BB = 4, S_CAB = 5, S_ALA = 6, S_W = 7, S_ALB = 8 };

// v_a(g), v_b(g) of the pair whose vectors are in xa/xb/pa/pb
__device__ __forceinline__ void vpair(const PrdP &P, int g, double &ua, double &ub, double &va, double &vb)
{
    ua = P.xa[g]; ub = P.xb[g];
    va = (P.pa[g] - P.scal[S_ALA] * ua) / P.scal[S_BA];
    vb = (P.pb[g] - ua * P.scal[S_W] - va * P.scal[S_CAB] - P.scal[S_ALB] * ub) / P.scal[S_BB];
}

// P.k = slot of column a of the NEW pair (columns c2 = i_base + k, c1 = c2 - 1); the previous pair
// (slots k+2, k+1) is finished here unless P.first.  P.has_next = 0: panel end, only finish.
__global__ void __launch_bounds__(VR * VS) next2_kernel(PrdP P)
{
    __shared__ double s_ru[2][MAXM], s_rv[2][MAXM];   // rows c2 (0) and c1 (1) of U and V, slots lo..m0-1
    __shared__ double s_acc[2][VS][VR];
    __shared__ double s_red[VR * VS / 32];
    __shared__ unsigned int s_last;
    const int k = P.k;
    const int c2 = P.i_base + k, c1 = c2 - 1, L = c2 - 1;
    const int Lp = L + 2;                 // reflector length of the previous pair's column a
    const int lo = k + 1, nl = P.m0 - lo; // finished slots
    if (P.has_next) {
        for (int l = threadIdx.x; l < nl; l += blockDim.x) {
            const int slot = lo + l;
            double u2, v2, u1, v1;
            if (!P.first && slot <= k + 2) {
                double ua, ub, va, vb;
                vpair(P, c2, ua, ub, va, vb);
                u2 = (slot == k + 2) ? ua : ub; v2 = (slot == k + 2) ? va : vb;
                vpair(P, c1, ua, ub, va, vb);
                u1 = (slot == k + 2) ? ua : ub; v1 = (slot == k + 2) ? va : vb;
            } else {
                u2 = P.U[(size_t)slot * P.npad + c2]; v2 = P.V[(size_t)slot * P.npad + c2];
                u1 = P.U[(size_t)slot * P.npad + c1]; v1 = P.V[(size_t)slot * P.npad + c1];
            }
            s_ru[0][l] = u2; s_rv[0][l] = v2; s_ru[1][l] = u1; s_rv[1][l] = v1;
        }
    }
    __syncthreads();
    const int r = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int g = blockIdx.x * VR + r;
    double ua = 0.0, ub = 0.0, va = 0.0, vb = 0.0;
    if (!P.first && g < Lp) {
        vpair(P, g, ua, ub, va, vb);
        if (sl == 0) {
            P.U[(size_t)(k + 2) * P.npad + g] = ua; P.V[(size_t)(k + 2) * P.npad + g] = va;
            P.U[(size_t)(k + 1) * P.npad + g] = ub; P.V[(size_t)(k + 1) * P.npad + g] = vb;
        }
    }
    if (!P.has_next) return;
    double acc_a = 0.0, acc_b = 0.0;
    if (g <= c2) {
        if (sl == 0) {
            acc_a = P.W[(size_t)k * P.npad + g];
            if (g <= c1) acc_b = P.W[(size_t)(k - 1) * P.npad + g];
        }
        for (int l = sl; l < nl; l += VS) {
            const int slot = lo + l;
            double uj, vj;
            if (!P.first && slot == k + 2) { uj = ua; vj = va; }
            else if (!P.first && slot == k + 1) { uj = ub; vj = vb; }
            else { uj = __ldg(P.U + (size_t)slot * P.npad + g); vj = __ldg(P.V + (size_t)slot * P.npad + g); }
            acc_a = fma(-uj, s_rv[0][l], acc_a); acc_a = fma(-vj, s_ru[0][l], acc_a);
            acc_b = fma(-uj, s_rv[1][l], acc_b); acc_b = fma(-vj, s_ru[1][l], acc_b);
        }
    }
    s_acc[0][sl][r] = acc_a; s_acc[1][sl][r] = acc_b;
    __syncthreads();
    double n22 = 0.0, n12 = 0.0;
    if (sl == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int q = 0; q < VS; q++) { a += s_acc[0][q][r]; b += s_acc[1][q][r]; }
        if (g <= c2) {
            P.xa_n[g] = a;
            P.xb_n[g] = (g <= c1) ? b : 0.0;
            if (g < L) { n22 = a * a; n12 = a * b; }
        }
        n22 = warp_sum(n22); n12 = warp_sum(n12);
    }
    if (threadIdx.x == 0) {
        P.part[blockIdx.x] = n22; P.part[gridDim.x + blockIdx.x] = n12;
        __threadfence();
        s_last = atomicAdd(&P.tickets[2], 1u);
    }
    __syncthreads();
    if (s_last == gridDim.x - 1) {
        __threadfence();
        double s0 = 0.0, s1 = 0.0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) { s0 += __ldcg(P.part + q); s1 += __ldcg(P.part + gridDim.x + q); }
        const double t22 = block_sum<VR * VS>(s0, s_red);
        const double t12 = block_sum<VR * VS>(s1, s_red);
        if (threadIdx.x == 0) {
            // Householder scalars of column a (src/eigen_prd_t4x.F:262-275; sigma = 0 -> g = 0, beta = 1)
            const double a_piv = __ldcg(P.xa_n + L - 1);
            double g_a, u_piv, beta;
            if (t22 != 0.0) { g_a = -copysign(sqrt(t22), a_piv); u_piv = a_piv - g_a; beta = -u_piv * g_a; }
            else { g_a = 0.0; u_piv = 0.0; beta = 1.0; }
            P.scal[S_GA] = g_a; P.scal[S_BA] = beta;
            P.scal[S_SAB] = t12 - g_a * __ldcg(P.xb_n + L - 1);     // u_a^T x_b
            P.xa_n[L - 1] = u_piv;
            P.d_out[c2] = __ldcg(P.xa_n + c2);       // T(c2,c2)
            P.e_out[c2] = __ldcg(P.xa_n + c2 - 1);   // T(c2-1,c2)     (e(i,1) = u_t(8), prd_t4x.F:329)
            P.d_out[c1] = __ldcg(P.xb_n + c1);
            P.e2_out[c2] = g_a;                      // T(c2-2,c2)     (e(i,2) = sgm(2))
            P.tickets[2] = 0u;
        }
    }
}

// x_b <- H_a x_b on rows < L ; scalars of H_b (pivot row L-2) ; u_a^T u_b
__global__ void __launch_bounds__(256) house2_kernel(PrdP P)
{
    __shared__ double s_red[8];
    __shared__ unsigned int s_last;
    const int c2 = P.i_base + P.k, c1 = c2 - 1, L = c2 - 1;
    const double coef = P.scal[S_SAB] / P.scal[S_BA];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    double n11 = 0.0, cd = 0.0;
    if (g < L) {
        const double ua = P.xa[g];
        const double b = P.xb[g] - ua * coef;
        P.xb[g] = b;
        if (g < L - 1) { n11 = b * b; cd = ua * b; }
    }
    n11 = block_sum<256>(n11, s_red);
    cd = block_sum<256>(cd, s_red);
    if (threadIdx.x == 0) {
        P.part[blockIdx.x] = n11; P.part[gridDim.x + blockIdx.x] = cd;
        __threadfence();
        s_last = atomicAdd(&P.tickets[1], 1u);
    }
    __syncthreads();
    if (s_last == gridDim.x - 1) {
        __threadfence();
        double s0 = 0.0, s1 = 0.0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) { s0 += __ldcg(P.part + q); s1 += __ldcg(P.part + gridDim.x + q); }
        const double t11 = block_sum<256>(s0, s_red);
        const double tcd = block_sum<256>(s1, s_red);
        if (threadIdx.x == 0) {
            const double b_piv = __ldcg(P.xb + L - 2);
            double g_b, u_piv, beta;
            if (t11 != 0.0) { g_b = -copysign(sqrt(t11), b_piv); u_piv = b_piv - g_b; beta = -u_piv * g_b; }
            else { g_b = 0.0; u_piv = 0.0; beta = 1.0; }
            P.scal[S_GB] = g_b; P.scal[S_BB] = beta;
            P.scal[S_CAB] = tcd - g_b * __ldcg(P.xa + L - 2);
            P.e2_out[c1] = g_b;                       // T(c1-2,c1)    (e(i-1,2) = sgm(1))
            P.e_out[c1] = __ldcg(P.xb + L - 1);       // T(c1-1,c1) = (H_a x_b)(L-1)   (e(i-1,1) = sgm(2) r12)
            P.xb[L - 2] = u_piv;
            P.xb[L - 1] = 0.0;
            P.tickets[1] = 0u;
        }
    }
}

__global__ void __launch_bounds__(256, 2) symv2_kernel(PrdP P, int gx, int ntile_blocks)
{
    __shared__ double smem[8 * TR];
    const int bid = blockIdx.x;
    if (bid >= ntile_blocks) {
        const double *uv[2] = {P.xa, P.xb};
        if (P.ndone > 0) dots_chunk<2>(P, uv, P.k + 1, bid - ntile_blocks);
        return;
    }
    const int nclL = ncl_of(P);
    int sc, br;
    if (!fold_triangle(P, bid, gx, nclL, sc, br)) return;
    SymvIO<2> io;
    io.u[0] = P.xa; io.prow[0] = P.Prow; io.pcol[0] = P.Pcol;
    io.u[1] = P.xb; io.prow[1] = P.Prow_b; io.pcol[1] = P.Pcol_b;
    symv_strip<2>(P, io, br, sc, nclL, smem);
}

// p_a, p_b from the tile partials minus the panel corrections; u_a^T p_a, u_b^T p_a, u_b^T p_b;
// last CTA: alpha_a, v_a^T u_b, alpha_b.  MODE as pvec_kernel.
template <int MODE>
__global__ void __launch_bounds__(VR * VS) pvec2_kernel(PrdP P, PeerView pv, unsigned long long epoch)
{
    __shared__ double s_st[2][2 * MAXM];
    __shared__ int s_nbr[1024];
    __shared__ double s_acc[2][VS][VR];
    __shared__ double s_red[VR * VS / 32];
    __shared__ unsigned int s_last;
    const int nd = P.ndone;
    const int nclL = ncl_of(P);
    const int nsc = nstrips_of(P, nclL);
    constexpr bool PARTIAL = (MODE == 0 || MODE == 1 || MODE == 3);
    constexpr bool FINISH = (MODE == 0 || MODE == 2 || MODE == 4);
    const int par = (int)(epoch & 1ull);
    if (MODE == 4 && threadIdx.x == 0) {
        const volatile unsigned long long *fl = pv.flags[pv.r] + (size_t)par * pv.P;
        for (int q = 0; q < pv.P; q++) {
            unsigned long long spins = 0;
            while (fl[q] < epoch) {
                __nanosleep(64);
                if (++spins > (1ull << 27)) { *pv.err = 1; break; }
            }
        }
        __threadfence_system();
    }
    if (PARTIAL) {
        for (int s = threadIdx.x; s < nsc; s += blockDim.x) s_nbr[s] = strip_rows(P, s, nclL);
    }
    if (FINISH) {
        for (int c = threadIdx.x; c < 2 * nd; c += blockDim.x) { s_st[0][c] = P.st[c]; s_st[1][c] = P.st[2 * MAXM + c]; }
    }
    __syncthreads();
    const int r = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int g = blockIdx.x * VR + r;
    const int first_slot = P.k + 1;
    double acc_a = 0.0, acc_b = 0.0;
    if (g < P.L) {
        if (PARTIAL) {
            const bool rown = (g % P.px) == P.x, coln = (g % P.py) == P.y;
            if (rown) {
                const int jl = g / P.px, br = jl / TR;
                for (int s = nsc - 1 - sl; s >= 0 && s_nbr[s] > br; s -= VS) {
                    acc_a += __ldcs(P.Prow + (size_t)s * P.ldprow + jl);
                    acc_b += __ldcs(P.Prow_b + (size_t)s * P.ldprow + jl);
                }
            }
            if (coln) {
                const int il = g / P.py;
                const int nb = ntile_rows(P, il / TC, nclL);
                for (int b = sl; b < nb; b += VS) {
                    acc_a += __ldcs(P.Pcol + (size_t)b * P.ldpcol + il);
                    acc_b += __ldcs(P.Pcol_b + (size_t)b * P.ldpcol + il);
                }
                if (rown && sl == 0) {
                    const double dg = P.A[(size_t)il * P.lda + g / P.px];
                    acc_a = fma(dg, P.xa[g], acc_a); acc_b = fma(dg, P.xb[g], acc_b);
                }
            }
        } else if (sl == 0) {
            if (MODE == 4) {
                const double *base = pv.slots[pv.r] + (size_t)par * pv.P * pv.slot_doubles + g;
                for (int q = 0; q < pv.P; q++) {
                    acc_a += __ldcg(base + (size_t)q * pv.slot_doubles);
                    acc_b += __ldcg(base + (size_t)q * pv.slot_doubles + P.npad);
                }
            } else { acc_a = P.pa[g]; acc_b = P.pb[g]; }
        }
        if (FINISH) {
            for (int l = sl; l < nd; l += VS) {
                const size_t off = (size_t)(first_slot + l) * P.npad + g;
                const double ul = __ldg(P.U + off), vl = __ldg(P.V + off);
                acc_a = fma(-ul, s_st[0][l], acc_a); acc_a = fma(-vl, s_st[0][nd + l], acc_a);
                acc_b = fma(-ul, s_st[1][l], acc_b); acc_b = fma(-vl, s_st[1][nd + l], acc_b);
            }
        }
    }
    s_acc[0][sl][r] = acc_a; s_acc[1][sl][r] = acc_b;
    __syncthreads();
    double daa = 0.0, dba = 0.0, dbb = 0.0;
    if (sl == 0) {
        double p_a = 0.0, p_b = 0.0;
#pragma unroll
        for (int q = 0; q < VS; q++) { p_a += s_acc[0][q][r]; p_b += s_acc[1][q][r]; }
        if (g < P.L) {
            if (MODE == 3) {
                const size_t off = ((size_t)par * pv.P + pv.r) * pv.slot_doubles + g;
                for (int q = 0; q < pv.P; q++) { pv.slots[q][off] = p_a; pv.slots[q][off + P.npad] = p_b; }
            } else { P.pa[g] = p_a; P.pb[g] = p_b; }
            if (FINISH) {
                const double ua = P.xa[g], ub = P.xb[g];
                daa = ua * p_a; dba = ub * p_a; dbb = ub * p_b;
            }
        }
        daa = warp_sum(daa); dba = warp_sum(dba); dbb = warp_sum(dbb);
    }
    if (MODE == 1) return;
    if (MODE == 3) {
        // one system-scope release per CTA (bar.sync orders the CTA's peer stores before it; the fence is cumulative)
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence_system(); s_last = atomicAdd(&P.tickets[3], 1u); }
        __syncthreads();
        if (s_last == gridDim.x - 1 && threadIdx.x == 0) {
            __threadfence_system();
            for (int q = 0; q < pv.P; q++) {
                volatile unsigned long long *fl = pv.flags[q] + (size_t)par * pv.P + pv.r;
                *fl = epoch;
            }
            P.tickets[3] = 0u;
        }
        return;
    }
    if (threadIdx.x == 0) {
        P.part[blockIdx.x] = daa; P.part[gridDim.x + blockIdx.x] = dba; P.part[2 * gridDim.x + blockIdx.x] = dbb;
        __threadfence();
        s_last = atomicAdd(&P.tickets[1], 1u);
    }
    __syncthreads();
    if (s_last == gridDim.x - 1) {
        __threadfence();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
            s0 += __ldcg(P.part + q); s1 += __ldcg(P.part + gridDim.x + q); s2 += __ldcg(P.part + 2 * gridDim.x + q);
        }
        const double taa = block_sum<VR * VS>(s0, s_red);
        const double tba = block_sum<VR * VS>(s1, s_red);
        const double tbb = block_sum<VR * VS>(s2, s_red);
        if (threadIdx.x == 0) {
            const double beta_a = P.scal[S_BA], beta_b = P.scal[S_BB], cab = P.scal[S_CAB];
            const double alpha_a = taa / (2.0 * beta_a);
            const double w = (tba - alpha_a * cab) / beta_a;          // v_a^T u_b
            const double alpha_b = (tbb - 2.0 * cab * w) / (2.0 * beta_b);
            P.scal[S_ALA] = alpha_a; P.scal[S_W] = w; P.scal[S_ALB] = alpha_b;
            P.tickets[1] = 0u;
        }
    }
}

// eigen_prd_final (src/eigen_prd_t8.F:207-315): the leading nrem x nrem block holds the band entries
// of the first columns; they are zeroed in a so that the back-transformation sees empty reflectors.
// out[0..2] = d(0..2), out[3..4] = e1(1..2), out[5] = e2(2)   (owners only; summed by the caller)
__global__ void prd_final_kernel(PrdP P, int nrem)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double *out = P.scal + 16;
    for (int q = 0; q < 6; q++) out[q] = 0.0;
    auto own = [&](int gr, int gc) { return (gr % P.px) == P.x && (gc % P.py) == P.y; };
    auto at = [&](int gr, int gc) -> double & { return P.A[(size_t)(gc / P.py) * P.lda + gr / P.px]; };
    for (int j = 0; j < nrem && j < P.n; j++)
        if (own(j, j)) out[j] = at(j, j);
    for (int j = 1; j < nrem && j < P.n; j++)
        if (own(j - 1, j)) { out[2 + j] = at(j - 1, j); at(j - 1, j) = 0.0; }
    if (nrem == 3 && P.n >= 3 && own(0, 2)) { out[5] = at(0, 2); at(0, 2) = 0.0; }
}
__global__ void prd_final_store_kernel(PrdP P, int nrem)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double *out = P.scal + 16;
    for (int j = 0; j < nrem && j < P.n; j++) P.d_out[j] = out[j];
    P.e_out[0] = 0.0; P.e2_out[0] = 0.0;
    if (P.n >= 2) { P.e_out[1] = out[3]; P.e2_out[1] = 0.0; }
    if (nrem == 3 && P.n >= 3) { P.e_out[2] = out[4]; P.e2_out[2] = out[5]; }
}

}  // namespace

// =========================================================================================
// host drivers
// =========================================================================================
double scaling_dev(int n, double *a, int lda)
{
    Context &c = ctx();
    const Grid &g = c.g;
    const int nrl = cyc_count(n, g.px, g.x), ncl = cyc_count(n, g.py, g.y);
    double *out = (double *)dev_alloc(2 * sizeof(double));
    EE_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(double), c.stream));
    if (nrl > 0 && ncl > 0) {
        dim3 grid(min(64, (nrl + 255) / 256), min(ncl, 4096));
        absmax_kernel<<<grid, 256, 0, c.stream>>>(a, lda, n, g.px, g.py, g.x, g.y, nrl, ncl, out);
        EE_CHECK_LAUNCH();
    }
    comm_allreduce_max(out, 2, COMM_WORLD, c.stream);
    double h[2];
    EE_CUDA(cudaMemcpyAsync(h, out, sizeof h, cudaMemcpyDeviceToHost, c.stream));
    EE_CUDA(cudaStreamSynchronize(c.stream));
    dev_free(out);
    if (h[1] != 0.0) return NAN;
    // LAPACK dsyev convention, eigen_scaling.F:76-134
    const double SAFMIN = 2.2250738585072014e-308, PREC = 2.220446049250313e-16;
    const double SMLNUM = SAFMIN / PREC, BIGNUM = 1.0 / SMLNUM;
    const double RMIN = sqrt(SMLNUM);
    const double RMAX = fmin(sqrt(BIGNUM), 1.0 / sqrt(sqrt(SAFMIN)));
    const double anrm = h[0];
    double sigma = 1.0;
    if (anrm != 0.0 && anrm < RMIN) sigma = RMIN / anrm;
    else if (anrm > RMAX) sigma = RMAX / anrm;
    if (sigma != 1.0 && nrl > 0 && ncl > 0) {
        dim3 grid((nrl + 255) / 256, std::min(ncl, 32768));
        scale_upper_kernel<<<grid, 256, 0, c.stream>>>(a, lda, n, g.px, g.py, g.x, g.y, nrl, ncl, sigma);
        EE_CHECK_LAUNCH();
    }
    return sigma;
}

// Launch shape of the multi-launch SYMV kernels (symv_kernel / symv2_kernel) for a trailing matrix of order L:
// strip length sw (tiles) and the folded grid gx x gy (fold_triangle: CTA column bx sweeps strip nsc-1-bx, then strip bx;
// gy = most tile rows any CTA column needs).  Long strips amortise the row-sum reduction at large L, short ones give the
// latency-bound small trailing matrices more CTAs.  Returns sw.
static int symv_launch_shape(const Grid &g, int L, int nclL, int *gx_out, int *gy_out)
{
    const int sw = (L > 12288) ? 4 : (L > 6144) ? 2 : 1;
    const int nsc = (nclL + sw * TC - 1) / (sw * TC);
    int gx = (nsc + 1) / 2, gy = 0;
    auto strip_rows_h = [&](int sc) {       // host twin of strip_rows()
        int tlast = std::min((sc + 1) * sw, (nclL + TC - 1) / TC) - 1;
        int clast = std::min((tlast + 1) * TC, nclL) - 1;
        if (clast < tlast * TC) return 0;
        long long cmax_g = (long long)clast * g.py + g.y;
        if (cmax_g <= g.x) return 0;
        return (int)((cmax_g - g.x - 1) / ((long long)TR * g.px)) + 1;
    };
    for (int bx = 0; bx < gx; bx++) {
        int s1 = nsc - 1 - bx, s2 = bx;
        int r = strip_rows_h(s1) + (s2 != s1 ? strip_rows_h(s2) : 0);
        if (r > gy) gy = r;
    }
    *gx_out = gx; *gy_out = gy;
    return sw;
}

// local padded dims used by the trd kernels
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
int trd_lda_pad(int nrl) { return round_up(nrl > 0 ? nrl : 1, TR); }
int trd_ncl_pad(int ncl) { return round_up(ncl > 0 ? ncl : 1, TC * SW); }

// Internal panel width for wide trailing matrices (even, <= MAXM; EIGENEXA_B200_WIDE_W overrides for tuning).
// With the cp.async GEMM a K = 256 trailing update paid for a 128-wide panel; the TMA kernel runs K = 96 at
// nearly the same rate, and the replicated vector kernels cost ~ L x panel width per column, so m_forward
// itself is used (N = 32000, 1 GPU: eigen_trd 9.00 s at width 48, 9.03 at 64, 9.13 at 96, 9.23 at 128).
static int wide_panel_width()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("EIGENEXA_B200_WIDE_W");
        v = e ? atoi(e) : 2;
        if (v < 2) v = 2;
        if (v > MAXM) v = MAXM;
        v -= v % 2;
    }
    return v;
}

void trd_dev(int n, double *a_user, int lda_user, double *d_out, double *e_out, int m_forward)
{
    Context &c = ctx();
    const Grid &g = c.g;
    cudaStream_t st = c.stream;
    const int nrl = cyc_count(n, g.px, g.x), ncl = cyc_count(n, g.py, g.y);
    EE_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n, st));
    EE_CUDA(cudaMemsetAsync(e_out, 0, sizeof(double) * n, st));
    if (n == 1) {
        if (g.x == 0 && g.y == 0) EE_CUDA(cudaMemcpyAsync(d_out, a_user, sizeof(double), cudaMemcpyDeviceToDevice, st));
        comm_allreduce_sum(d_out, 1, COMM_WORLD, st);
        return;
    }
    int m = m_forward < n ? m_forward : n;
    if (m < 1) m = 1;
    if (m > MAXM) m = MAXM;
    // m_forward is the reference's blocking hint (default 48).  The trailing update is one
    // GEMM with K = 2*width; at K = 96 its C read-modify-write (16 B per 4*width flop) is not
    // hidden, so wide trailing matrices use a wider internal panel (same T up to rounding).
    // The panel-internal vector work grows with the width, hence only above WIDE_L.
    const int WIDE_W = wide_panel_width();
    constexpr int WIDE_L = 16384;
    const int mmax = (n > WIDE_L && m < WIDE_W) ? WIDE_W : m;

    // ---- internal padded copy of the local matrix ------------------------------------
    const int lda = trd_lda_pad(nrl), nclp = trd_ncl_pad(ncl);
    double *A = (double *)dev_alloc((size_t)lda * nclp * sizeof(double));
    EE_CUDA(cudaMemsetAsync(A, 0, (size_t)lda * nclp * sizeof(double), st));
    if (nrl > 0 && ncl > 0)
        EE_CUDA(cudaMemcpy2DAsync(A, (size_t)lda * sizeof(double), a_user, (size_t)lda_user * sizeof(double),
                                  (size_t)nrl * sizeof(double), ncl, cudaMemcpyDeviceToDevice, st));
    {
        dim3 grid((lda + 255) / 256, std::min(nclp, 32768));
        zero_lower_kernel<<<grid, 256, 0, st>>>(A, lda, nclp, n, g.px, g.py, g.x, g.y);
        EE_CHECK_LAUNCH();
    }

    // ---- workspace ---------------------------------------------------------------------
    const int npad = round_up(n + 1, 256);
    const int nstrip_max = (nclp + TC - 1) / TC;   // strips of one tile (sw = 1) need the most rows
    const int nbr_max = lda / TR;
    const int maxvb = (n + VR) / VR + 2;
    size_t wsz = 0;
    auto take = [&](size_t cnt) { size_t o = wsz; wsz += (cnt + 31) & ~(size_t)31; return o; };
    size_t oU = take((size_t)npad * mmax), oV = take((size_t)npad * mmax), oW = take((size_t)npad * mmax);
    size_t oU1 = take(npad), oU2 = take(npad), oP = take(npad);
    size_t oProw = take((size_t)nstrip_max * lda), oPcol = take((size_t)nbr_max * nclp);
    size_t oDots = take((size_t)NCH * 2 * MAXM), oSt = take(2 * MAXM), oPart = take(2 * (size_t)maxvb);
    size_t oScal = take(32), oTick = take(32);
    size_t oUVx = take((size_t)lda * 2 * mmax), oVUy = take((size_t)nclp * 2 * mmax);
    size_t oCtl = take(32), oPartAB = take(2 * 2048);
    double *ws = (double *)dev_alloc(wsz * sizeof(double));
    EE_CUDA(cudaMemsetAsync(ws, 0, wsz * sizeof(double), st));

    TrdP P;
    memset(&P, 0, sizeof P);
    P.piv = -1;
    P.A = A; P.lda = lda; P.px = g.px; P.py = g.py; P.x = g.x; P.y = g.y;
    P.n = n; P.npad = npad;
    P.U = ws + oU; P.V = ws + oV; P.W = ws + oW;
    P.ucur = ws + oU1; P.unext = ws + oU2; P.pbuf = ws + oP;
    P.Prow = ws + oProw; P.Pcol = ws + oPcol; P.ldprow = lda; P.ldpcol = nclp;
    P.dots_part = ws + oDots; P.st = ws + oSt; P.part = ws + oPart; P.scal = ws + oScal;
    P.tickets = reinterpret_cast<unsigned int *>(ws + oTick);
    P.d_out = d_out; P.e_out = e_out;
    double *UVx = ws + oUVx, *VUy = ws + oVUy;
    const bool multi = g.nnod > 1;
    PeerView pv;
    memset(&pv, 0, sizeof pv);
    // (two vectors per slot, as eigen_prd needs: a later eigen_sx of the same size then reuses the ring as it is)
    const bool use_peer = multi && comm_peer_setup((size_t)2 * npad, &pv);

    // ---- persistent panel kernel (one cooperative launch per panel) unless switched off / debugging ----------
    bool persist = (!multi || use_peer) && c.profiling < 2;
    {
        const char *e = getenv("EIGENEXA_B200_TRD_PERSIST");
        if (e && e[0] == '0') persist = false;
    }
    int pgrid = 0;
    PanelCtl ctl;
    memset(&ctl, 0, sizeof ctl);
    if (persist) {
        int per_sm = 0;
        if (multi) EE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trd_panel_kernel<true>, 256, 0));
        else EE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trd_panel_kernel<false>, 256, 0));
        pgrid = per_sm * c.sm_count;
        if (pgrid > 2048) pgrid = 2048;
        if (pgrid < 1) persist = false;
        ctl.bar = reinterpret_cast<unsigned long long *>(ws + oCtl);
        ctl.work = ctl.bar + 1;
        ctl.tacc = c.profiling ? ws + oCtl + 4 : nullptr;
        ctl.pivraw = ws + oCtl + 12;
        ctl.partA = ws + oPartAB; ctl.partB = ctl.partA + 2048;
    }

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float t_symv = 0.f, t_syr2k = 0.f, t_pvec = 0.f, t_vvec = 0.f;
    if (c.profiling) { EE_CUDA(cudaEventCreate(&ev0)); EE_CUDA(cudaEventCreate(&ev1)); }
    // level 1: event pairs recorded without synchronising (read after the final sync)
    std::vector<cudaEvent_t> pool_symv, pool_syr2k;
    size_t pool_next = 0;
    auto pool_get = [&]() {
        if (pool_next == c.ev_pool.size()) { cudaEvent_t e; EE_CUDA(cudaEventCreate(&e)); c.ev_pool.push_back(e); }
        return c.ev_pool[pool_next++];
    };
    c.symv_trace.clear();
    auto prof_begin = [&](int cls = 0) {
        if (c.profiling >= 2) EE_CUDA(cudaEventRecord(ev0, st));
        else if (c.profiling == 1 && cls) {
            cudaEvent_t e = pool_get(); EE_CUDA(cudaEventRecord(e, st));
            (cls == 1 ? pool_symv : pool_syr2k).push_back(e);
        }
    };
    auto prof_end = [&](float &acc, int cls = 0) {
        if (c.profiling >= 2) {
            EE_CUDA(cudaEventRecord(ev1, st));
            EE_CUDA(cudaEventSynchronize(ev1));
            float ms; EE_CUDA(cudaEventElapsedTime(&ms, ev0, ev1)); acc += ms;
            if (cls == 1) c.symv_trace.push_back(ms);
        } else if (c.profiling == 1 && cls) {
            cudaEvent_t e = pool_get(); EE_CUDA(cudaEventRecord(e, st));
            (cls == 1 ? pool_symv : pool_syr2k).push_back(e);
        }
    };
    auto wall = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tw0 = wall();
    if (c.profiling >= 2) EE_CUDA(cudaStreamSynchronize(st));
    double tw1 = wall();

    // panels from the right; boundaries at multiples of m (as the reference) while the
    // narrow width is in use, free-running for the wide panels
    int col_end = n;
    while (col_end > 2) {
        int i_base;
        if (col_end > WIDE_L && mmax > m) i_base = col_end - mmax;
        else i_base = ((col_end - 1) / m) * m;
        if (i_base < 0) i_base = 0;
        const int m0 = col_end - i_base;
        const int ib = (i_base > 0) ? 2 : 1;   // ib > 1 <=> a trailing matrix remains
        const int k_stop = (i_base == 0) ? 2 : 0;
        col_end = i_base;
        P.i_base = i_base; P.m0 = m0;
        // panel load (+ zero U,V)
        {
            dim3 grid((npad + 255) / 256, m0);
            TrdP Q = P;
            panel_load_kernel<<<grid, 256, 0, st>>>(Q, 1);
            EE_CHECK_LAUNCH();
            if (multi) comm_allreduce_sum(P.W, (size_t)npad * m0, COMM_WORLD, st);
        }
        if (persist) {
            // every column step of the panel in one cooperative launch
            TrdP Q = P;
            EE_CUDA(cudaMemsetAsync(ctl.bar, 0, 2 * sizeof(unsigned long long), st));
            PanelCtl C = ctl;
            C.k_stop = k_stop;
            C.epoch0 = multi ? comm_peer_reserve_epochs(m0 - k_stop) : 0ull;
            void *args[] = {(void *)&Q, (void *)&pv, (void *)&C};
            prof_begin(1);
            if (multi) EE_CUDA(cudaLaunchCooperativeKernel((void *)trd_panel_kernel<true>, dim3(pgrid), dim3(256), args, 0, st));
            else EE_CUDA(cudaLaunchCooperativeKernel((void *)trd_panel_kernel<false>, dim3(pgrid), dim3(256), args, 0, st));
            EE_CHECK_LAUNCH();
            prof_end(t_symv, 1);
        }
        // panel prologue: first column of the panel (slot m0-1) is the raw panel copy
        if (!persist) {
            TrdP Q = P;
            Q.k = m0 - 1; Q.L = i_base + m0 - 1; Q.first = 1; Q.has_next = 1; Q.ndone = 0;
            Q.unext = P.ucur;  // write straight into ucur
            int nb = (Q.L + 1 + VR - 1) / VR;
            vvec_kernel<<<nb, VR * VS, 0, st>>>(Q);
            EE_CHECK_LAUNCH();
        }
        for (int k = m0 - 1; k >= k_stop && !persist; k--) {
            const int i = i_base + k, L = i;
            if (c.debug_maxcols > 0 && (n - 1 - i) >= c.debug_maxcols) { col_end = 0; break; }
            TrdP Q = P;
            Q.k = k; Q.L = L; Q.ndone = m0 - 1 - k; Q.first = 0;
            Q.has_next = (k - 1 >= k_stop) ? 1 : 0;
            // ---- SYMV --------------------------------------------------------------
            const int nclL = cyc_count(L, g.py, g.y);
            // long strips amortise the row-sum reduction at large L; short ones give the
            // latency-bound small trailing matrices more CTAs
            int gx, gy;
            Q.sw = symv_launch_shape(g, L, nclL, &gx, &gy);
            const int ntile_blocks = gx * gy;
            const int nblocks = ntile_blocks + (Q.ndone > 0 ? NCH : 0);
            prof_begin(1);
            if (nblocks > 0) {
                symv_kernel<<<nblocks, 256, 0, st>>>(Q, gx > 0 ? gx : 1, ntile_blocks);
                EE_CHECK_LAUNCH();
            }
            prof_end(t_symv, 1);
            // ---- p, alpha --------------------------------------------------------------
            const int nvb = (L + VR - 1) / VR;
            prof_begin();
            if (!multi) {
                pvec_kernel<0><<<nvb, VR * VS, 0, st>>>(Q, pv, 0ull);
                EE_CHECK_LAUNCH();
            } else if (use_peer) {
                const unsigned long long epoch = comm_peer_next_epoch();
                pvec_kernel<3><<<nvb, VR * VS, 0, st>>>(Q, pv, epoch);
                EE_CHECK_LAUNCH();
                pvec_kernel<4><<<nvb, VR * VS, 0, st>>>(Q, pv, epoch);
                EE_CHECK_LAUNCH();
            } else {
                pvec_kernel<1><<<nvb, VR * VS, 0, st>>>(Q, pv, 0ull);
                EE_CHECK_LAUNCH();
                comm_allreduce_sum(P.pbuf, L, COMM_WORLD, st);
                pvec_kernel<2><<<nvb, VR * VS, 0, st>>>(Q, pv, 0ull);
                EE_CHECK_LAUNCH();
            }
            prof_end(t_pvec);
            prof_begin();
            // ---- v, next column ------------------------------------------------------
            vvec_kernel<<<nvb, VR * VS, 0, st>>>(Q);
            EE_CHECK_LAUNCH();
            prof_end(t_vvec);
            std::swap(P.ucur, P.unext);
        }
        // ---- panel end: reflectors back into A, trailing rank-2k update ----------------
        {
            TrdP Q = P;
            dim3 grid((lda + 255) / 256, m0);
            panel_restore_kernel<<<grid, 256, 0, st>>>(Q, k_stop);
            EE_CHECK_LAUNCH();
        }
        if (ib > 1) {
            const int nrl_b = cyc_count(i_base, g.px, g.x), ncl_b = cyc_count(i_base, g.py, g.y);
            if (nrl_b > 0 && ncl_b > 0) {
                TrdP Q = P;
                int mx = nrl_b > ncl_b ? nrl_b : ncl_b;
                dim3 grid((mx + 255) / 256, 2 * m0);
                pack_uv_kernel<<<grid, 256, 0, st>>>(Q, UVx, lda, nrl_b, VUy, nclp, ncl_b, m0);
                EE_CHECK_LAUNCH();
                TriSpec tri; tri.mode = 1; tri.px = g.px; tri.py = g.py; tri.x = g.x; tri.y = g.y;
                prof_begin(2);
                dgemm(st, 'N', 'T', nrl_b, ncl_b, 2 * m0, -1.0, UVx, lda, VUy, nclp, 1.0, A, lda, tri);
                prof_end(t_syr2k, 2);
            }
        }
        // profiling aid (ncu targets): stop after the panels that cover debug_maxcols columns
        if (persist && c.debug_maxcols > 0 && n - i_base >= c.debug_maxcols) { col_end = 0; break; }
    }
    double tw2 = wall();
    if (c.profiling >= 2) { EE_CUDA(cudaStreamSynchronize(st)); tw2 = wall(); }
    // ---- final 2x2 (trd_t8.F:188-219) -------------------------------------------------------
    {
        TrdP Q = P;
        trd_final_kernel<<<1, 32, 0, st>>>(Q);
        EE_CHECK_LAUNCH();
        if (multi) comm_allreduce_sum(P.scal + 4, 3, COMM_WORLD, st);
        trd_final_store_kernel<<<1, 32, 0, st>>>(Q);
        EE_CHECK_LAUNCH();
    }
    // reflectors back to the caller's array
    if (nrl > 0 && ncl > 0)
        EE_CUDA(cudaMemcpy2DAsync(a_user, (size_t)lda_user * sizeof(double), A, (size_t)lda * sizeof(double),
                                  (size_t)nrl * sizeof(double), ncl, cudaMemcpyDeviceToDevice, st));
    double tacc_h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (persist && ctl.tacc) EE_CUDA(cudaMemcpyAsync(tacc_h, ctl.tacc, sizeof tacc_h, cudaMemcpyDeviceToHost, st));
    EE_CUDA(cudaStreamSynchronize(st));
    if (use_peer) {
        int herr = 0;
        EE_CUDA(cudaMemcpy(&herr, pv.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) fatal("peer all-reduce timed out waiting for another rank", __FILE__, __LINE__);
    }
    double tw3 = wall();
    if (c.profiling == 1) {
        auto drain = [&](std::vector<cudaEvent_t> &pool, float &acc, bool trace) {
            for (size_t i = 0; i + 1 < pool.size(); i += 2) {
                float ms = 0.f; EE_CUDA(cudaEventElapsedTime(&ms, pool[i], pool[i + 1])); acc += ms;
                if (trace) c.symv_trace.push_back(ms);
            }
            pool.clear();
        };
        drain(pool_symv, t_symv, true); drain(pool_syr2k, t_syr2k, false);
    }
    dev_free(ws);
    dev_free(A);
    double tw4 = wall();
    if (c.profiling) {
        c.timings[5] = t_symv * 1e-3; c.timings[6] = t_syr2k * 1e-3;
        c.timings[7] = t_pvec * 1e-3; c.timings[8] = t_vvec * 1e-3;
        // persistent kernel: timings[5] is the event time of the panel kernels; the in-kernel split (globaltimer of
        // CTA 0 around the phases of every column) goes to [15] SYMV phase, [16] p phase, [31] v phase
        c.timings[15] = tacc_h[0] * 1e-9; c.timings[16] = tacc_h[1] * 1e-9; c.timings[31] = tacc_h[2] * 1e-9;
        for (int i = 0; i < 4; i++) c.timings[32 + i] = tacc_h[3 + i] * 1e-9;   // p phase on a grid: stores, barrier, flags, sum
        c.timings[9] = tw1 - tw0; c.timings[10] = tw2 - tw1; c.timings[11] = tw3 - tw2; c.timings[12] = tw4 - tw3;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
}

// -----------------------------------------------------------------------------------------
// eigen_prd driver: a (lda x ncl, local cyclic part) in/out; d, e1, e2 device arrays of length n
// (replicated).  e1(c) = T(c-1,c), e2(c) = T(c-2,c) (0-based c).
// -----------------------------------------------------------------------------------------
void prd_dev(int n, double *a_user, int lda_user, double *d_out, double *e1_out, double *e2_out, int m_forward)
{
    Context &c = ctx();
    const Grid &g = c.g;
    cudaStream_t st = c.stream;
    const int nrl = cyc_count(n, g.px, g.x), ncl = cyc_count(n, g.py, g.y);
    EE_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n, st));
    EE_CUDA(cudaMemsetAsync(e1_out, 0, sizeof(double) * n, st));
    EE_CUDA(cudaMemsetAsync(e2_out, 0, sizeof(double) * n, st));
    const int nrem = 2 + n % 2;     // MBAND + mod(n, MBAND) leading columns are never reduced (src/eigen_prd.F:392-401)
    int m = m_forward < n ? m_forward : n;
    m -= m % 2;                     // the pair loop needs an even width (manual 4.4)
    if (m < 2) m = 2;
    if (m > MAXM) m = MAXM;
    const int WIDE_W = wide_panel_width();
    constexpr int WIDE_L = 16384;
    const int mmax = (n > WIDE_L && m < WIDE_W) ? WIDE_W : m;

    const int lda = trd_lda_pad(nrl), nclp = trd_ncl_pad(ncl);
    double *A = (double *)dev_alloc((size_t)lda * nclp * sizeof(double));
    EE_CUDA(cudaMemsetAsync(A, 0, (size_t)lda * nclp * sizeof(double), st));
    if (nrl > 0 && ncl > 0)
        EE_CUDA(cudaMemcpy2DAsync(A, (size_t)lda * sizeof(double), a_user, (size_t)lda_user * sizeof(double),
                                  (size_t)nrl * sizeof(double), ncl, cudaMemcpyDeviceToDevice, st));
    {
        dim3 grid((lda + 255) / 256, std::min(nclp, 32768));
        zero_lower_kernel<<<grid, 256, 0, st>>>(A, lda, nclp, n, g.px, g.py, g.x, g.y);
        EE_CHECK_LAUNCH();
    }
    const int npad = round_up(n + 1, 256);
    const int nstrip_max = (nclp + TC - 1) / TC;
    const int nbr_max = lda / TR;
    const int maxvb = (n + VR) / VR + 2;
    size_t wsz = 0;
    auto take = [&](size_t cnt) { size_t o = wsz; wsz += (cnt + 31) & ~(size_t)31; return o; };
    size_t oU = take((size_t)npad * mmax), oV = take((size_t)npad * mmax), oW = take((size_t)npad * mmax);
    size_t oX = take((size_t)4 * npad), oP = take((size_t)2 * npad);
    size_t oProw = take((size_t)2 * nstrip_max * lda), oPcol = take((size_t)2 * nbr_max * nclp);
    size_t oDots = take((size_t)NCH * 4 * MAXM), oSt = take(4 * MAXM), oPart = take(3 * (size_t)maxvb);
    size_t oScal = take(32), oTick = take(32);
    size_t oUVx = take((size_t)lda * 2 * mmax), oVUy = take((size_t)nclp * 2 * mmax);
    double *ws = (double *)dev_alloc(wsz * sizeof(double));
    EE_CUDA(cudaMemsetAsync(ws, 0, wsz * sizeof(double), st));

    PrdP P;
    memset(&P, 0, sizeof P);
    P.A = A; P.lda = lda; P.px = g.px; P.py = g.py; P.x = g.x; P.y = g.y;
    P.n = n; P.npad = npad;
    P.U = ws + oU; P.V = ws + oV; P.W = ws + oW;
    P.xa = ws + oX; P.xb = P.xa + npad; P.xa_n = P.xb + npad; P.xb_n = P.xa_n + npad;
    P.pa = ws + oP; P.pb = P.pa + npad; P.pbuf = P.pa;
    P.Prow = ws + oProw; P.Prow_b = P.Prow + (size_t)nstrip_max * lda; P.ldprow = lda;
    P.Pcol = ws + oPcol; P.Pcol_b = P.Pcol + (size_t)nbr_max * nclp; P.ldpcol = nclp;
    P.dots_part = ws + oDots; P.st = ws + oSt; P.part = ws + oPart; P.scal = ws + oScal;
    P.tickets = reinterpret_cast<unsigned int *>(ws + oTick);
    P.d_out = d_out; P.e_out = e1_out; P.e2_out = e2_out;
    double *UVx = ws + oUVx, *VUy = ws + oVUy;
    const bool multi = g.nnod > 1;
    PeerView pv;
    memset(&pv, 0, sizeof pv);
    const bool use_peer = multi && comm_peer_setup((size_t)2 * npad, &pv);

    std::vector<cudaEvent_t> pool_symv, pool_syr2k;
    size_t pool_next = 0;
    auto pool_get = [&]() {
        if (pool_next == c.ev_pool.size()) { cudaEvent_t e; EE_CUDA(cudaEventCreate(&e)); c.ev_pool.push_back(e); }
        return c.ev_pool[pool_next++];
    };
    auto mark = [&](int cls) {
        if (c.profiling < 1) return;
        cudaEvent_t e = pool_get(); EE_CUDA(cudaEventRecord(e, st));
        (cls == 1 ? pool_symv : pool_syr2k).push_back(e);
    };
    c.symv_trace.clear();

    int col_end = n;
    while (col_end > nrem) {
        int i_base;
        if (col_end > WIDE_L && mmax > m) i_base = col_end - mmax;
        else i_base = nrem + ((col_end - nrem - 1) / m) * m;
        if (i_base < nrem) i_base = nrem;
        const int m0 = col_end - i_base;   // even: n - nrem and every width are even
        col_end = i_base;
        P.i_base = i_base; P.m0 = m0;
        {
            dim3 grid((npad + 255) / 256, m0);
            PrdP Q = P;
            panel_load_kernel<<<grid, 256, 0, st>>>(Q, 1);
            EE_CHECK_LAUNCH();
            if (multi) comm_allreduce_sum(P.W, (size_t)npad * m0, COMM_WORLD, st);
        }
        bool stop = false;
        for (int k = m0 - 1; k >= 1; k -= 2) {
            const int c2 = i_base + k, L = c2 - 1;
            if (c.debug_maxcols > 0 && (n - 1 - c2) >= c.debug_maxcols) { stop = true; break; }
            PrdP Q = P;
            Q.k = k; Q.L = L; Q.first = (k == m0 - 1); Q.has_next = 1; Q.ndone = m0 - 1 - k;
            // ---- previous pair -> panel, raw pair, H_a scalars ------------------------------------
            next2_kernel<<<(c2 + 1 + VR - 1) / VR, VR * VS, 0, st>>>(Q);
            EE_CHECK_LAUNCH();
            std::swap(P.xa, P.xa_n); std::swap(P.xb, P.xb_n);
            Q.xa = P.xa; Q.xb = P.xb; Q.xa_n = P.xa_n; Q.xb_n = P.xb_n;
            // ---- H_a on column b, H_b scalars ------------------------------------------------------
            house2_kernel<<<(L + 255) / 256, 256, 0, st>>>(Q);
            EE_CHECK_LAUNCH();
            // ---- [p_a p_b] = A [u_a u_b] in one pass --------------------------------------------------
            const int nclL = cyc_count(L, g.py, g.y);
            int gx, gy;
            Q.sw = symv_launch_shape(g, L, nclL, &gx, &gy);
            const int ntile_blocks = gx * gy;
            const int nblocks = ntile_blocks + (Q.ndone > 0 ? NCH : 0);
            mark(1);
            if (nblocks > 0) {
                symv2_kernel<<<nblocks, 256, 0, st>>>(Q, gx > 0 ? gx : 1, ntile_blocks);
                EE_CHECK_LAUNCH();
            }
            mark(1);
            // ---- p, alpha ----------------------------------------------------------------------------
            const int nvb = (L + VR - 1) / VR;
            if (!multi) {
                pvec2_kernel<0><<<nvb, VR * VS, 0, st>>>(Q, pv, 0ull);
                EE_CHECK_LAUNCH();
            } else if (use_peer) {
                const unsigned long long epoch = comm_peer_next_epoch();
                pvec2_kernel<3><<<nvb, VR * VS, 0, st>>>(Q, pv, epoch);
                EE_CHECK_LAUNCH();
                pvec2_kernel<4><<<nvb, VR * VS, 0, st>>>(Q, pv, epoch);
                EE_CHECK_LAUNCH();
            } else {
                pvec2_kernel<1><<<nvb, VR * VS, 0, st>>>(Q, pv, 0ull);
                EE_CHECK_LAUNCH();
                comm_allreduce_sum(P.pa, (size_t)npad + L, COMM_WORLD, st);
                pvec2_kernel<2><<<nvb, VR * VS, 0, st>>>(Q, pv, 0ull);
                EE_CHECK_LAUNCH();
            }
        }
        if (stop) { col_end = 0; break; }
        // ---- panel end: last pair -> panel, reflectors back into A, trailing rank-2k update ------
        {
            PrdP Q = P;
            Q.k = -1; Q.first = 0; Q.has_next = 0;
            next2_kernel<<<(i_base + 1 + VR - 1) / VR, VR * VS, 0, st>>>(Q);
            EE_CHECK_LAUNCH();
            dim3 grid((lda + 255) / 256, m0);
            panel_restore_kernel<<<grid, 256, 0, st>>>(Q, 0);
            EE_CHECK_LAUNCH();
        }
        const int nrl_b = cyc_count(i_base, g.px, g.x), ncl_b = cyc_count(i_base, g.py, g.y);
        if (nrl_b > 0 && ncl_b > 0) {
            PrdP Q = P;
            int mx = nrl_b > ncl_b ? nrl_b : ncl_b;
            dim3 grid((mx + 255) / 256, 2 * m0);
            pack_uv_kernel<<<grid, 256, 0, st>>>(Q, UVx, lda, nrl_b, VUy, nclp, ncl_b, m0);
            EE_CHECK_LAUNCH();
            TriSpec tri; tri.mode = 1; tri.px = g.px; tri.py = g.py; tri.x = g.x; tri.y = g.y;
            mark(2);
            dgemm(st, 'N', 'T', nrl_b, ncl_b, 2 * m0, -1.0, UVx, lda, VUy, nclp, 1.0, A, lda, tri);
            mark(2);
        }
    }
    // ---- leading nrem x nrem block (prd_t8.F:207-315) ------------------------------------------------
    {
        PrdP Q = P;
        prd_final_kernel<<<1, 32, 0, st>>>(Q, nrem);
        EE_CHECK_LAUNCH();
        if (multi) comm_allreduce_sum(P.scal + 16, 6, COMM_WORLD, st);
        prd_final_store_kernel<<<1, 32, 0, st>>>(Q, nrem);
        EE_CHECK_LAUNCH();
    }
    if (nrl > 0 && ncl > 0)
        EE_CUDA(cudaMemcpy2DAsync(a_user, (size_t)lda_user * sizeof(double), A, (size_t)lda * sizeof(double),
                                  (size_t)nrl * sizeof(double), ncl, cudaMemcpyDeviceToDevice, st));
    EE_CUDA(cudaStreamSynchronize(st));
    if (use_peer) {
        int herr = 0;
        EE_CUDA(cudaMemcpy(&herr, pv.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) fatal("peer all-reduce timed out waiting for another rank", __FILE__, __LINE__);
    }
    if (c.profiling >= 1) {
        float t_symv = 0.f, t_syr2k = 0.f;
        auto drain = [&](std::vector<cudaEvent_t> &pool, float &acc, bool trace) {
            for (size_t i = 0; i + 1 < pool.size(); i += 2) {
                float ms = 0.f; EE_CUDA(cudaEventElapsedTime(&ms, pool[i], pool[i + 1])); acc += ms;
                if (trace) c.symv_trace.push_back(ms);
            }
            pool.clear();
        };
        drain(pool_symv, t_symv, true); drain(pool_syr2k, t_syr2k, false);
        c.timings[5] = t_symv * 1e-3; c.timings[6] = t_syr2k * 1e-3;
        c.timings[15] = c.timings[16] = c.timings[31] = 0.0;   // (no persistent kernel on this path)
    }
    dev_free(ws);
    dev_free(A);
}

}  // namespace ee
