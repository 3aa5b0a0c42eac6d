/*
 * oracle/shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Prototype-only stand-in for <mpi.h> so that the UNMODIFIED reference sources under
 * /root/reference/src compile in an image that has no MPI installed.  Implemented by
 * mpi_shm.c (fork + POSIX shared memory rings).  Only the 29 functions and the handful
 * of constants the reference touches are declared.
 */
#ifndef CFDP_ORACLE_SHIM_MPI_H
#define CFDP_ORACLE_SHIM_MPI_H

#include <stddef.h>

#define MPI_VERSION 3
#define MPI_SUCCESS 0

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Win;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long MPI_Aint;

typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; int count; } MPI_Status;

#define MPI_COMM_WORLD      0
#define MPI_BYTE            1
#define MPI_CHAR            1
#define MPI_REQUEST_NULL    (-1)
#define MPI_UNDEFINED       (-32766)
#define MPI_STATUS_IGNORE   ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_INFO_NULL       0

#define MPI_THREAD_SINGLE     0
#define MPI_THREAD_FUNNELED   1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE   3

#define MPI_MODE_NOSTORE    1
#define MPI_MODE_NOPRECEDE  2
#define MPI_MODE_NOSUCCEED  4
#define MPI_WIN_MODEL       1
#define MPI_WIN_UNIFIED     1
#define MPI_WIN_SEPARATE    2

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Barrier(MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st);
int MPI_Testany(int n, MPI_Request *reqs, int *index, int *flag, MPI_Status *st);

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr);
int MPI_Info_create(MPI_Info *info);
int MPI_Info_set(MPI_Info info, const char *key, const char *value);
int MPI_Win_allocate(MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, void *baseptr, MPI_Win *win);
int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win);
int MPI_Win_get_attr(MPI_Win win, int keyval, void *attr, int *flag);
int MPI_Win_free(MPI_Win *win);
int MPI_Win_fence(int assert_, MPI_Win win);
int MPI_Win_post(MPI_Group g, int assert_, MPI_Win win);
int MPI_Win_start(MPI_Group g, int assert_, MPI_Win win);
int MPI_Win_complete(MPI_Win win);
int MPI_Win_wait(MPI_Win win);
int MPI_Put(const void *origin, int ocount, MPI_Datatype odt, int target, MPI_Aint disp,
            int tcount, MPI_Datatype tdt, MPI_Win win);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *g);
int MPI_Group_incl(MPI_Group g, int n, const int ranks[], MPI_Group *out);
int MPI_Group_free(MPI_Group *g);

#endif
