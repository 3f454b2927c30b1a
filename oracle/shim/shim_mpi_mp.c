/* ORACLE TEST INFRASTRUCTURE — not product code.
 *
 * Multi-process stand-in for <mpi.h>, so that the UNMODIFIED reference runs its own rank
 * decomposition (Y slabs, src/sim.c:116-130, 177-187; ring neighbours src/plasma.c) on the
 * host cores of one box without an MPI installation:
 *
 *   CPIC_SHIM_NPROCS=P ./cpic_ref_mp conf      (P = 1 behaves like the loop-back shim)
 *
 * MPI_Init_thread (called by the reference's main, src/cpic.c:82-96) creates one UNIX socket
 * pair per pair of ranks and one shared anonymous mapping, then forks P-1 children: the
 * caller returns as rank 0, the children as ranks 1..P-1. Point-to-point messages travel
 * over the sockets (non-blocking writes; whoever cannot make progress drains its incoming
 * sockets into a private list of pending messages, so no send/receive order can deadlock);
 * receives match (source, tag) in FIFO order, as MPI does. Barriers and the shared FFT work
 * array (shim_fftw_mp.c) live in the shared mapping.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#define MAXP 256

typedef struct msg {
	int tag;
	size_t bytes;
	char *data;
	struct msg *next;
} msg_t;

typedef struct peer {
	int fd;
	/* reassembly of the message being read */
	struct { int tag; int pad; size_t bytes; } hdr;
	size_t hdr_got, data_got;
	msg_t *cur;
	msg_t *head, *tail;      /* complete messages, oldest first */
} peer_t;

struct shared {
	atomic_int bar_count;
	atomic_int bar_sense;
	atomic_int failed;
	size_t arena_size;
	_Alignas(64) char arena[];
};

static int g_rank = 0, g_size = 1, g_init = 0;
static peer_t g_peer[MAXP];
static struct shared *g_sh;
static size_t g_arena_used;      /* the same on every rank: collective bump allocation */
static int g_local_sense = 0;
static pid_t g_child[MAXP];

int shim_mp_rank(void) { return g_rank; }
int shim_mp_size(void) { return g_size; }

static size_t type_size(MPI_Datatype t) { return (size_t) t; }

static void
die_(const char *what)
{
	fprintf(stderr, "shim MPI rank %d: %s: %s\n", g_rank, what, strerror(errno));
	if(g_sh) atomic_store(&g_sh->failed, 1);
	abort();
}

/* Reads whatever has arrived from peer r; complete messages go to its pending list */
static int
drain(int r)
{
	peer_t *p = &g_peer[r];
	int got = 0;
	for(;;)
	{
		ssize_t n;
		if(p->hdr_got < sizeof(p->hdr))
		{
			n = read(p->fd, (char *) &p->hdr + p->hdr_got, sizeof(p->hdr) - p->hdr_got);
			if(n < 0) { if(errno == EAGAIN || errno == EINTR) return got; die_("read"); }
			if(n == 0) return got;          /* peer closed: it has finished */
			p->hdr_got += (size_t) n;
			if(p->hdr_got < sizeof(p->hdr)) continue;
			p->cur = malloc(sizeof(msg_t));
			if(!p->cur) die_("malloc");
			p->cur->tag = p->hdr.tag;
			p->cur->bytes = p->hdr.bytes;
			p->cur->data = malloc(p->hdr.bytes ? p->hdr.bytes : 1);
			p->cur->next = NULL;
			if(!p->cur->data) die_("malloc");
			p->data_got = 0;
		}
		if(p->data_got < p->cur->bytes)
		{
			n = read(p->fd, p->cur->data + p->data_got, p->cur->bytes - p->data_got);
			if(n < 0) { if(errno == EAGAIN || errno == EINTR) return got; die_("read"); }
			if(n == 0) return got;
			p->data_got += (size_t) n;
			if(p->data_got < p->cur->bytes) continue;
		}
		if(p->tail) p->tail->next = p->cur; else p->head = p->cur;
		p->tail = p->cur;
		p->cur = NULL;
		p->hdr_got = 0;
		got++;
	}
}

static void
progress(int wait_ms)
{
	int any = 0;
	for(int r = 0; r < g_size; r++)
		if(r != g_rank) any += drain(r);
	if(g_sh && atomic_load(&g_sh->failed)) { fprintf(stderr, "shim MPI rank %d: a peer failed\n", g_rank); _exit(3); }
	if(!any && wait_ms > 0)
	{
		struct pollfd pf[MAXP];
		int n = 0;
		for(int r = 0; r < g_size; r++)
			if(r != g_rank) { pf[n].fd = g_peer[r].fd; pf[n].events = POLLIN; pf[n].revents = 0; n++; }
		poll(pf, (nfds_t) n, wait_ms);
	}
}

static void
enqueue_self(const void *buf, size_t bytes, int tag)
{
	peer_t *p = &g_peer[g_rank];
	msg_t *m = malloc(sizeof(*m));
	if(!m) die_("malloc");
	m->tag = tag; m->bytes = bytes; m->next = NULL;
	m->data = malloc(bytes ? bytes : 1);
	if(!m->data) die_("malloc");
	memcpy(m->data, buf, bytes);
	if(p->tail) p->tail->next = m; else p->head = m;
	p->tail = m;
}

static void
send_bytes(int dst, const void *buf, size_t bytes)
{
	const char *c = buf;
	size_t done = 0;
	while(done < bytes)
	{
		ssize_t n = write(g_peer[dst].fd, c + done, bytes - done);
		if(n < 0)
		{
			if(errno == EAGAIN || errno == EINTR) { progress(0); continue; }
			die_("write");
		}
		done += (size_t) n;
	}
}

int
MPI_Send(const void *buf, int count, MPI_Datatype type, int dst, int tag, MPI_Comm comm)
{
	(void) comm;
	size_t bytes = (size_t) count * type_size(type);
	if(dst < 0 || dst >= g_size) { fprintf(stderr, "shim MPI: send to rank %d of %d\n", dst, g_size); abort(); }
	if(dst == g_rank) { enqueue_self(buf, bytes, tag); return MPI_SUCCESS; }
	struct { int tag; int pad; size_t bytes; } hdr = { tag, 0, bytes };
	send_bytes(dst, &hdr, sizeof(hdr));
	send_bytes(dst, buf, bytes);
	return MPI_SUCCESS;
}

int
MPI_Isend(const void *buf, int count, MPI_Datatype type, int dst, int tag, MPI_Comm comm, MPI_Request *req)
{
	if(req) *req = MPI_REQUEST_NULL;
	return MPI_Send(buf, count, type, dst, tag, comm);
}

static msg_t *
take(int src, int tag)
{
	peer_t *p = &g_peer[src];
	msg_t *m, *prev = NULL;
	for(m = p->head; m; prev = m, m = m->next)
		if(tag == MPI_ANY_TAG || m->tag == tag) break;
	if(!m) return NULL;
	if(prev) prev->next = m->next; else p->head = m->next;
	if(p->tail == m) p->tail = prev;
	return m;
}

int
MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *status)
{
	(void) comm;
	size_t cap = (size_t) count * type_size(type);
	msg_t *m = NULL;
	int from = src;
	for(long spin = 0; !m; spin++)
	{
		if(src == MPI_ANY_SOURCE)
		{
			for(int r = 0; r < g_size && !m; r++) if((m = take(r, tag))) from = r;
		}
		else m = take(src, tag);
		if(m) break;
		if(g_size == 1)
		{
			fprintf(stderr, "shim MPI: receive with tag %x would block forever\n", tag);
			abort();
		}
		progress(spin > 100 ? 1 : 0);
	}
	if(m->bytes > cap)
	{
		fprintf(stderr, "shim MPI: message of %zu bytes truncated to %zu\n", m->bytes, cap);
		abort();
	}
	memcpy(buf, m->data, m->bytes);
	if(status)
	{
		status->MPI_SOURCE = from;
		status->MPI_TAG = m->tag;
		status->MPI_ERROR = MPI_SUCCESS;
		status->count = (int) m->bytes;
	}
	free(m->data);
	free(m);
	return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{ (void) status; if(req) *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }

int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count)
{ *count = (int) ((size_t) status->count / type_size(type)); return MPI_SUCCESS; }

/* Sense-reversing barrier in the shared mapping; waiting ranks keep draining their sockets */
void
shim_mp_barrier(void)
{
	if(g_size == 1) return;
	g_local_sense = !g_local_sense;
	if(atomic_fetch_add(&g_sh->bar_count, 1) == g_size - 1)
	{
		atomic_store(&g_sh->bar_count, 0);
		atomic_store(&g_sh->bar_sense, g_local_sense);
	}
	else
	{
		long spin = 0;
		while(atomic_load(&g_sh->bar_sense) != g_local_sense)
		{
			if((++spin & 63) == 0) progress(0);
			if(spin > 200000) { progress(1); }
		}
	}
}

int MPI_Barrier(MPI_Comm c) { (void) c; shim_mp_barrier(); return MPI_SUCCESS; }

#define TAG_BCAST 0x7fffb000

int
MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
	if(g_size == 1) return MPI_SUCCESS;
	if(g_rank == root)
	{
		for(int r = 0; r < g_size; r++)
			if(r != root) MPI_Send(b, n, t, r, TAG_BCAST, c);
	}
	else MPI_Recv(b, n, t, root, TAG_BCAST, c, MPI_STATUS_IGNORE);
	return MPI_SUCCESS;
}

/* Collective allocation in the shared mapping: every rank calls with the same sizes in the
 * same order and gets the same address */
void *
shim_mp_shared_alloc(size_t bytes)
{
	bytes = (bytes + 63) & ~(size_t) 63;
	if(g_size == 1 && !g_sh)
	{
		void *p = NULL;
		if(posix_memalign(&p, 64, bytes ? bytes : 64)) abort();
		return p;
	}
	if(g_arena_used + bytes > g_sh->arena_size)
	{
		fprintf(stderr, "shim MPI: shared arena exhausted (%zu + %zu > %zu); raise CPIC_SHIM_ARENA_MB\n",
				g_arena_used, bytes, g_sh->arena_size);
		abort();
	}
	void *p = g_sh->arena + g_arena_used;
	g_arena_used += bytes;
	return p;
}

static void
start_ranks(void)
{
	const char *e = getenv("CPIC_SHIM_NPROCS");
	int P = e ? atoi(e) : 1;
	if(P < 1) P = 1;
	if(P > MAXP) P = MAXP;
	g_size = P;
	g_rank = 0;
	if(P == 1) return;

	const char *a = getenv("CPIC_SHIM_ARENA_MB");
	size_t arena = (size_t) (a ? atol(a) : 2048) << 20;
	g_sh = mmap(NULL, sizeof(struct shared) + arena, PROT_READ | PROT_WRITE,
			MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
	if(g_sh == MAP_FAILED) die_("mmap");
	atomic_store(&g_sh->bar_count, 0);
	atomic_store(&g_sh->bar_sense, 0);
	atomic_store(&g_sh->failed, 0);
	g_sh->arena_size = arena;

	/* fd[i][j]: rank i's end of the pair (i, j) */
	int (*fd)[MAXP] = malloc(sizeof(int[MAXP]) * (size_t) P);
	if(!fd) die_("malloc");
	for(int i = 0; i < P; i++)
		for(int j = i + 1; j < P; j++)
		{
			int sv[2];
			if(socketpair(AF_UNIX, SOCK_STREAM, 0, sv)) die_("socketpair (raise ulimit -n?)");
			int sz = 4 << 20;
			setsockopt(sv[0], SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
			setsockopt(sv[1], SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
			fd[i][j] = sv[0];
			fd[j][i] = sv[1];
		}
	fflush(NULL);
	for(int r = 1; r < P; r++)
	{
		pid_t pid = fork();
		if(pid < 0) die_("fork");
		if(pid == 0)
		{
			g_rank = r;
			prctl(PR_SET_PDEATHSIG, SIGKILL);
			break;
		}
		g_child[r] = pid;
	}
	for(int i = 0; i < P; i++)
		for(int j = 0; j < P; j++)
		{
			if(i == j) continue;
			if(i == g_rank)
			{
				g_peer[j].fd = fd[i][j];
				fcntl(fd[i][j], F_SETFL, fcntl(fd[i][j], F_GETFL) | O_NONBLOCK);
			}
			else close(fd[i][j]);
		}
	free(fd);
	signal(SIGPIPE, SIG_IGN);
}

int
MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{
	(void) argc; (void) argv;
	if(provided) *provided = required;
	if(!g_init) { g_init = 1; start_ranks(); }
	return MPI_SUCCESS;
}

int MPI_Init(int *argc, char ***argv) { int p; return MPI_Init_thread(argc, argv, 0, &p); }

int
MPI_Finalize(void)
{
	if(g_size > 1)
	{
		shim_mp_barrier();
		fflush(NULL);
		if(g_rank == 0)
		{
			for(int r = 1; r < g_size; r++)
			{
				int st = 0;
				waitpid(g_child[r], &st, 0);
			}
		}
		else _exit(0);
	}
	return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm c, int *rank) { (void) c; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void) c; *size = g_size; return MPI_SUCCESS; }
int MPI_Comm_set_errhandler(MPI_Comm c, MPI_Errhandler e) { (void) c; (void) e; return MPI_SUCCESS; }

int
MPI_Abort(MPI_Comm c, int code)
{
	(void) c; (void) code;
	if(g_sh) atomic_store(&g_sh->failed, 1);
	abort();
}
