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

}  // namespace ee
