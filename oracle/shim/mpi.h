/* ORACLE TEST INFRASTRUCTURE — not product code.
 * Single-rank loop-back stand-in for <mpi.h>, enough for the reference's hot path
 * (src/comm_field.c:15-46, src/comm_plasma.c:887-984, src/sim.c:93-94,578,
 * src/solver.c:482, src/cpic.c:82-96). Rank 0 of 1: a send to self is copied into a
 * tagged mailbox and the matching receive pops it. */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H

#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Errhandler;
typedef int MPI_Info;
typedef int MPI_Win;
typedef long MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_THREAD_MULTIPLE 3
#define MPI_TASK_MULTIPLE 4
#define MPI_ERRORS_RETURN 1
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL 0

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype type, int dst, int tag, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dst, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count);
int MPI_Abort(MPI_Comm comm, int code);

#endif
