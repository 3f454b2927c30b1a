/* ORACLE TEST INFRASTRUCTURE — not product code.
 * Loop-back MPI for one rank. Messages are matched by tag in FIFO order. */
#include "mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct msg {
	int tag;
	size_t bytes;
	void *data;
	struct msg *next;
} msg_t;

static msg_t *head, *tail;

static size_t type_size(MPI_Datatype t) { return (size_t) t; }

int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return MPI_SUCCESS; }
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{ (void) argc; (void) argv; *provided = required; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void) c; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void) c; *size = 1; return MPI_SUCCESS; }
int MPI_Comm_set_errhandler(MPI_Comm c, MPI_Errhandler e) { (void) c; (void) e; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm c) { (void) c; return MPI_SUCCESS; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void) b; (void) n; (void) t; (void) root; (void) c; return MPI_SUCCESS; }

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dst, int tag, MPI_Comm comm)
{
	msg_t *m;
	(void) comm;
	if(dst != 0) { fprintf(stderr, "shim MPI: send to rank %d\n", dst); abort(); }
	m = malloc(sizeof(*m));
	if(!m) abort();
	m->tag = tag;
	m->bytes = (size_t) count * type_size(type);
	m->data = malloc(m->bytes ? m->bytes : 1);
	if(!m->data) abort();
	memcpy(m->data, buf, m->bytes);
	m->next = NULL;
	if(tail) tail->next = m; else head = m;
	tail = m;
	return MPI_SUCCESS;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dst, int tag, MPI_Comm comm, MPI_Request *req)
{
	if(req) *req = MPI_REQUEST_NULL;
	return MPI_Send(buf, count, type, dst, tag, comm);
}

int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *status)
{
	msg_t *m, *prev = NULL;
	size_t cap = (size_t) count * type_size(type);
	(void) comm; (void) src;
	for(m = head; m; prev = m, m = m->next)
		if(tag == MPI_ANY_TAG || m->tag == tag) break;
	if(!m)
	{
		fprintf(stderr, "shim MPI: receive with tag %x would block forever\n", tag);
		abort();
	}
	if(m->bytes > cap)
	{
		fprintf(stderr, "shim MPI: message of %zu bytes truncated to %zu\n", m->bytes, cap);
		abort();
	}
	memcpy(buf, m->data, m->bytes);
	if(status)
	{
		status->MPI_SOURCE = 0;
		status->MPI_TAG = m->tag;
		status->MPI_ERROR = MPI_SUCCESS;
		status->count = (int) m->bytes;
	}
	if(prev) prev->next = m->next; else head = m->next;
	if(tail == m) tail = prev;
	free(m->data);
	free(m);
	return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{ (void) status; if(req) *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }

int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count)
{ *count = (int) ((size_t) status->count / type_size(type)); return MPI_SUCCESS; }

int MPI_Abort(MPI_Comm c, int code) { (void) c; (void) code; abort(); }
