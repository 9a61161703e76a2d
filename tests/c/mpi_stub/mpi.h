/* Minimal MPI declarations for a syntax check of bindings/eigen_init_mpi.c (the image has no MPI). */
#ifndef EE_MPI_STUB_H
#define EE_MPI_STUB_H
typedef int MPI_Comm;
typedef int MPI_Fint;
typedef int MPI_Info;
#define MPI_COMM_TYPE_SHARED 0
#define MPI_INFO_NULL 0
#define MPI_BYTE 0
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm *);
int MPI_Comm_free(MPI_Comm *);
int MPI_Bcast(void *, int, int, int, MPI_Comm);
MPI_Comm MPI_Comm_f2c(MPI_Fint);
#endif
