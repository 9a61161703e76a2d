// ee_comm.h -- NCCL plumbing that replaces comm_mod (src/comm.F) and
// MPI_Allreduce_group (src/MPI_Allreduce_group.F90).  One rank per GPU; every collective
// short-circuits on a 1-rank group exactly like the reference does
// (src/comm.F:781-783,1403-1408).
#pragma once
#include "ee_common.cuh"

namespace ee {

enum CommId { COMM_WORLD = 0, COMM_X = 1, COMM_Y = 2 };

// returns 0 on success; loads libnccl lazily (dlopen) only when nranks > 1
int comm_init(const unsigned char *unique_id, int rank, int nranks, const Grid &g);
void comm_finalize();
int comm_get_unique_id(unsigned char *id128);

void comm_allreduce_sum(double *buf, size_t count, CommId which, cudaStream_t st);  // reduce_dbl
void comm_allreduce_max(double *buf, size_t count, CommId which, cudaStream_t st);  // eigen_scaling.F:114
void comm_bcast(double *buf, size_t count, int root, CommId which, cudaStream_t st);  // bcast_dbl
void comm_allgather(const double *send, double *recv, size_t count_per_rank, CommId which, cudaStream_t st);
void comm_send(const double *buf, size_t count, int peer, cudaStream_t st);  // world ranks
void comm_recv(double *buf, size_t count, int peer, cudaStream_t st);
void comm_group_start();
void comm_group_end();
void comm_barrier(cudaStream_t st);  // MPI_Barrier(TRD_COMM_WORLD)

// ---- one-shot all-reduce over NVLink peer memory (fused into the trd vector kernels) --------
// Every rank owns a buffer [2 parities][P slots][slot_doubles] + flags, mapped into all peers
// with CUDA IPC.  A producer kernel stores its partial vector into slot[parity][my rank] of
// EVERY rank and then publishes an epoch flag; the consumer kernel waits for the P flags and
// sums the P slots in rank order (identical bits on every rank, as the reference's hand-made
// reductions guarantee, src/comm.F:2100-2406).
constexpr int PEER_MAX = 16;
struct PeerView {
    int P, r;
    unsigned long long slot_doubles;
    double *slots[PEER_MAX];                 // base of every rank's slot area (peer mapped)
    unsigned long long *flags[PEER_MAX];     // base of every rank's flag area [2][P]
    int *err;                                // local error flag (spin timeout)
};
// returns true when the peer ring is usable with at least slot_doubles per slot
bool comm_peer_setup(size_t slot_doubles, PeerView *view);
unsigned long long comm_peer_next_epoch();
// reserves n consecutive epochs (one per column of a persistent panel launch); returns the first
unsigned long long comm_peer_reserve_epochs(int n);

}  // namespace ee
