// TEST INFRASTRUCTURE ONLY (oracle).  In-process stand-in for the handful of
// MPI calls the reference's mpi_coordinator uses (src/mpi_coordinator.cc:4-69,
// src/mpi_coordinator.h:21), so that the reference's search_worker.cc and
// mpi_coordinator.cc compile and run UNMODIFIED in this container, which has no
// OpenMPI.  An MPI "rank" is a thread: the driver (oracle/ref_driver.cc) sets
// the world size, then starts one thread per rank with a thread-local rank.
// Collectives rendezvous on a pthread barrier; Gather/Gatherv copy in rank
// order, exactly what the reference relies on (src/mpi_coordinator.cc:50-59).
#ifndef VC_ORACLE_SHIM_MPI_H
#define VC_ORACLE_SHIM_MPI_H

#include <iostream>   // the reference leans on OpenMPI's transitive <iostream> (mpi_coordinator.cc:22)
#include <cstddef>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_INT        4
#define MPI_LONG_LONG  8
#define MPI_BOR        1
#define MPI_SUCCESS    0

extern "C" {
// driver-side controls (not MPI)
void vc_shim_mpi_set_world(int size);
void vc_shim_mpi_set_rank(int rank);

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype,
               void* recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype,
                void* recvbuf, const int* recvcounts, const int* displs,
                MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm);
}

#endif
