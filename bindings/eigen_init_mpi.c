/*
 * eigen_init_mpi.c -- the MPI-side shim a reference maintainer compiles next to libeigenexa_b200.so
 * (INTEGRATION.md section 1).  NOT built in this repository: the image has no MPI.
 *
 * Replaces the body of the reference's C/EigenExa.c:8-15 (eigen_init(MPI_Comm, order) -> MPI_Comm_c2f ->
 * eigen_libs_eigen_init_): the communicator is turned into the rank record + NCCL bootstrap id the library takes.
 *   mpicc -c eigen_init_mpi.c -I../include && link with -leigenexa_b200
 */
#include <mpi.h>
#include "eigenexa_b200.h"

/* drop-in for the reference's  void eigen_init(MPI_Comm comm, const char *order)  */
void eigen_init_mpi(MPI_Comm comm, const char *order)
{
    eigenexa_b200_comm_t c;
    MPI_Comm local;
    int lrank = 0, i;
    for (i = 0; i < EIGENEXA_B200_UNIQUE_ID_BYTES; i++) c.unique_id[i] = 0;
    c.reserved = 0;
    MPI_Comm_rank(comm, &c.rank);
    MPI_Comm_size(comm, &c.nranks);
    /* one rank per GPU of the node: local rank = CUDA device ordinal */
    MPI_Comm_split_type(comm, MPI_COMM_TYPE_SHARED, c.rank, MPI_INFO_NULL, &local);
    MPI_Comm_rank(local, &lrank);
    MPI_Comm_free(&local);
    c.device = lrank;
    if (c.nranks > 1) {
        if (c.rank == 0) eigenexa_b200_get_unique_id(c.unique_id);
        MPI_Bcast(c.unique_id, EIGENEXA_B200_UNIQUE_ID_BYTES, MPI_BYTE, 0, comm);
    }
    eigen_init(&c, order);   /* grid chosen as src/eigen_libs0.F:526-540; order 'C' (default) or 'R' */
}

/* Fortran face: comm_f is an MPI_Fint (what `use mpi` hands out), order a NUL-terminated string */
void eigen_init_mpi_f(int comm_f, const char *order)
{
    eigen_init_mpi(MPI_Comm_f2c((MPI_Fint)comm_f), order);
}
