// ee_comm.cu -- NCCL over NVLink in place of comm_mod's MPI wrappers
// (reference: src/comm.F:726 bcast_dbl, :1065 bcastw_dbl, :1192 reduce_dbl (= all-reduce),
//  :1278 allgather_dbl, :1377-1730 datacast_dbl*; src/eigen_libs0.F:579-585 comm split).
//
// libnccl is opened lazily with dlopen so that single-GPU jobs (where every collective
// short-circuits, as in the reference with x_nnod = y_nnod = 1) need no NCCL at all and the
// library loads on a CPU-only box for the symbol checks.
#include "ee_comm.h"
#include <dlfcn.h>

namespace ee {

// minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id, enums
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load()
{
    if (g_nccl.h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return false; }
#define EE_SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, name);                               \
    if (!g_nccl.field) { set_error("libnccl lacks %s", name); return false; }
    EE_SYM(GetUniqueId, "ncclGetUniqueId")
    EE_SYM(CommInitRank, "ncclCommInitRank")
    EE_SYM(CommSplit, "ncclCommSplit")
    EE_SYM(CommDestroy, "ncclCommDestroy")
    EE_SYM(AllReduce, "ncclAllReduce")
    EE_SYM(Broadcast, "ncclBroadcast")
    EE_SYM(AllGather, "ncclAllGather")
    EE_SYM(Send, "ncclSend")
    EE_SYM(Recv, "ncclRecv")
    EE_SYM(GroupStart, "ncclGroupStart")
    EE_SYM(GroupEnd, "ncclGroupEnd")
    EE_SYM(GetErrorString, "ncclGetErrorString")
#undef EE_SYM
    g_nccl.h = h;
    return true;
}

#define EE_NCCL(call)                                                                    \
    do {                                                                                 \
        ncclResult_t r__ = (call);                                                       \
        if (r__ != 0) {                                                                  \
            char buf__[512];                                                             \
            snprintf(buf__, sizeof buf__, "NCCL error %s: %s", #call, g_nccl.GetErrorString(r__)); \
            ::ee::fatal(buf__, __FILE__, __LINE__);                                      \
        }                                                                                \
    } while (0)

struct Comm {
    ncclComm_t c[3] = {nullptr, nullptr, nullptr};
    int size[3] = {1, 1, 1};
    int rank[3] = {0, 0, 0};
    // peer ring
    bool peer_tried = false, peer_ok = false;
    size_t peer_slot = 0;
    void *peer_local = nullptr;
    void *peer_ptr[PEER_MAX] = {nullptr};
    unsigned long long epoch = 0;
    int *peer_err = nullptr;
};

int comm_get_unique_id(unsigned char *id128)
{
    if (!nccl_load()) return 1;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) { set_error("ncclGetUniqueId failed"); return 2; }
    memcpy(id128, id.internal, 128);
    return 0;
}

int comm_init(const unsigned char *unique_id, int rank, int nranks, const Grid &g)
{
    Context &c = ctx();
    c.comm = new Comm();
    c.comm->size[COMM_WORLD] = nranks; c.comm->rank[COMM_WORLD] = rank;
    c.comm->size[COMM_X] = g.px; c.comm->rank[COMM_X] = g.x;
    c.comm->size[COMM_Y] = g.py; c.comm->rank[COMM_Y] = g.y;
    if (nranks == 1) return 0;
    if (!nccl_load()) return 1;
    ncclUniqueId id;
    memcpy(id.internal, unique_id, 128);
    EE_NCCL(g_nccl.CommInitRank(&c.comm->c[COMM_WORLD], nranks, id, rank));
    // x group: ranks sharing a process column (same y), ordered by x; y group: same x
    // (MPI_Comm_split(comm, y_inod, x_inod) / (comm, x_inod, y_inod), eigen_libs0.F:579-585)
    // NCCL_SPLIT_NOCOLOR (-1) for 1-rank groups: every rank still calls split collectively
    EE_NCCL(g_nccl.CommSplit(c.comm->c[COMM_WORLD], g.px > 1 ? g.y : -1, g.x, &c.comm->c[COMM_X], nullptr));
    EE_NCCL(g_nccl.CommSplit(c.comm->c[COMM_WORLD], g.py > 1 ? g.x : -1, g.y, &c.comm->c[COMM_Y], nullptr));
    return 0;
}

static void peer_release(Comm *cm_)
{
    if (!cm_->peer_local) return;
    for (int q = 0; q < cm_->size[COMM_WORLD]; q++)
        if (q != cm_->rank[COMM_WORLD] && cm_->peer_ptr[q]) cudaIpcCloseMemHandle(cm_->peer_ptr[q]);
    cudaFree(cm_->peer_local);
    if (cm_->peer_err) cudaFree(cm_->peer_err);
    cm_->peer_local = nullptr; cm_->peer_err = nullptr; cm_->peer_ok = false; cm_->peer_slot = 0;
}

bool comm_peer_setup(size_t slot_doubles, PeerView *view)
{
    Comm *m = ctx().comm;
    if (!m || m->size[COMM_WORLD] <= 1 || m->size[COMM_WORLD] > PEER_MAX) return false;
    const char *off = getenv("EIGENEXA_B200_NO_PEER");
    if (off && off[0] == '1') return false;
    const int P = m->size[COMM_WORLD], r = m->rank[COMM_WORLD];
    cudaStream_t st = ctx().stream;
    if (m->peer_tried && !m->peer_ok && m->peer_slot >= slot_doubles) return false;
    if (!m->peer_ok || m->peer_slot < slot_doubles) {
        // (re)build: collective over all ranks (every rank sees the same sizes)
        EE_CUDA(cudaStreamSynchronize(st));
        comm_barrier(st);
        peer_release(m);
        m->peer_tried = true;
        m->peer_slot = slot_doubles;
        const size_t bytes = (size_t)2 * P * slot_doubles * sizeof(double) + 2 * PEER_MAX * sizeof(unsigned long long);
        if (cudaMalloc(&m->peer_local, bytes) != cudaSuccess) { cudaGetLastError(); m->peer_local = nullptr; }
        int okl = m->peer_local != nullptr;
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof h);
        if (okl && cudaIpcGetMemHandle(&h, m->peer_local) != cudaSuccess) { cudaGetLastError(); okl = 0; }
        if (okl) EE_CUDA(cudaMemsetAsync(m->peer_local, 0, bytes, st));
        EE_CUDA(cudaMalloc((void **)&m->peer_err, sizeof(int)));
        EE_CUDA(cudaMemsetAsync(m->peer_err, 0, sizeof(int), st));
        // exchange (ok flag + 64-byte handle) through NCCL as raw bytes
        const size_t rec = 8 + sizeof(cudaIpcMemHandle_t);
        unsigned char *d_all = (unsigned char *)dev_alloc(rec * P);
        std::vector<unsigned char> h_all(rec * P, 0);
        long long okll = okl;
        memcpy(h_all.data() + rec * r, &okll, 8);
        memcpy(h_all.data() + rec * r + 8, &h, sizeof h);
        EE_CUDA(cudaMemcpyAsync(d_all + rec * r, h_all.data() + rec * r, rec, cudaMemcpyHostToDevice, st));
        EE_NCCL(g_nccl.AllGather(d_all + rec * r, d_all, rec, /*ncclChar*/ 0, m->c[COMM_WORLD], st));
        EE_CUDA(cudaMemcpyAsync(h_all.data(), d_all, rec * P, cudaMemcpyDeviceToHost, st));
        EE_CUDA(cudaStreamSynchronize(st));
        dev_free(d_all);
        bool all_ok = true;
        for (int q = 0; q < P; q++) { long long v; memcpy(&v, h_all.data() + rec * q, 8); all_ok = all_ok && v == 1; }
        int opened = 1;
        if (all_ok) {
            for (int q = 0; q < P; q++) {
                if (q == r) { m->peer_ptr[q] = m->peer_local; continue; }
                cudaIpcMemHandle_t hq;
                memcpy(&hq, h_all.data() + rec * q + 8, sizeof hq);
                if (cudaIpcOpenMemHandle(&m->peer_ptr[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError(); m->peer_ptr[q] = nullptr; opened = 0;
                }
            }
        } else opened = 0;
        // everyone must agree
        double *d_ok = (double *)dev_alloc(sizeof(double));
        double hv = opened ? 0.0 : 1.0;
        EE_CUDA(cudaMemcpyAsync(d_ok, &hv, sizeof(double), cudaMemcpyHostToDevice, st));
        comm_allreduce_max(d_ok, 1, COMM_WORLD, st);
        EE_CUDA(cudaMemcpyAsync(&hv, d_ok, sizeof(double), cudaMemcpyDeviceToHost, st));
        EE_CUDA(cudaStreamSynchronize(st));
        dev_free(d_ok);
        m->peer_ok = (hv == 0.0);
        if (!m->peer_ok) {
            if (r == 0) fprintf(stderr, "[eigenexa_b200] CUDA IPC peer ring unavailable; using NCCL all-reduce per column\n");
            return false;
        }
    }
    view->P = P; view->r = r; view->slot_doubles = m->peer_slot; view->err = m->peer_err;
    for (int q = 0; q < P; q++) {
        view->slots[q] = (double *)m->peer_ptr[q];
        view->flags[q] = (unsigned long long *)((double *)m->peer_ptr[q] + (size_t)2 * P * m->peer_slot);
    }
    return true;
}

unsigned long long comm_peer_next_epoch() { return ++ctx().comm->epoch; }
unsigned long long comm_peer_reserve_epochs(int n)
{
    Comm *m = ctx().comm;
    const unsigned long long first = m->epoch + 1;
    m->epoch += (unsigned long long)(n > 0 ? n : 0);
    return first;
}

void comm_finalize()
{
    Context &c = ctx();
    if (!c.comm) return;
    peer_release(c.comm);
    for (int i = 2; i >= 0; i--)
        if (c.comm->c[i]) g_nccl.CommDestroy(c.comm->c[i]);
    delete c.comm;
    c.comm = nullptr;
}

static inline Comm *cm() { return ctx().comm; }

void comm_allreduce_sum(double *buf, size_t count, CommId w, cudaStream_t st)
{
    if (!cm() || cm()->size[w] <= 1 || count == 0) return;
    EE_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, cm()->c[w], st));
}
void comm_allreduce_max(double *buf, size_t count, CommId w, cudaStream_t st)
{
    if (!cm() || cm()->size[w] <= 1 || count == 0) return;
    EE_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclMax, cm()->c[w], st));
}
void comm_bcast(double *buf, size_t count, int root, CommId w, cudaStream_t st)
{
    if (!cm() || cm()->size[w] <= 1 || count == 0) return;
    EE_NCCL(g_nccl.Broadcast(buf, buf, count, ncclFloat64, root, cm()->c[w], st));
}
void comm_allgather(const double *send, double *recv, size_t count, CommId w, cudaStream_t st)
{
    if (!cm() || cm()->size[w] <= 1) {
        if (send != recv && count) EE_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return;
    }
    EE_NCCL(g_nccl.AllGather(send, recv, count, ncclFloat64, cm()->c[w], st));
}
void comm_send(const double *buf, size_t count, int peer, cudaStream_t st)
{
    EE_NCCL(g_nccl.Send(buf, count, ncclFloat64, peer, cm()->c[COMM_WORLD], st));
}
void comm_recv(double *buf, size_t count, int peer, cudaStream_t st)
{
    EE_NCCL(g_nccl.Recv(buf, count, ncclFloat64, peer, cm()->c[COMM_WORLD], st));
}
void comm_group_start() { if (cm() && cm()->size[COMM_WORLD] > 1) EE_NCCL(g_nccl.GroupStart()); }
void comm_group_end() { if (cm() && cm()->size[COMM_WORLD] > 1) EE_NCCL(g_nccl.GroupEnd()); }
void comm_barrier(cudaStream_t st)
{
    if (!cm() || cm()->size[COMM_WORLD] <= 1) return;
    double *d = (double *)dev_alloc(sizeof(double));
    EE_CUDA(cudaMemsetAsync(d, 0, sizeof(double), st));
    EE_NCCL(g_nccl.AllReduce(d, d, 1, ncclFloat64, ncclSum, cm()->c[COMM_WORLD], st));
    EE_CUDA(cudaStreamSynchronize(st));
    dev_free(d);
}

}  // namespace ee
