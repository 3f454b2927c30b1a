/* simt-check: a stand-in for the nine NCCL entry points cpic_b200 resolves with dlopen, for
 * the CPU test suite (CPIC_B200_NCCL=.../libfake_nccl.so). Ranks are processes on one host;
 * messages are files in a directory named by the unique id. Point-to-point operations of one
 * (source, destination) pair match in posting order, as in NCCL; inside a group all sends are
 * posted before any receive completes. Test infrastructure only: it never runs on the GPU box
 * and is never loaded unless CPIC_B200_NCCL names it. */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclSystemError = 2, ncclInvalidArgument = 4 };

#define MAXR 64
#define TIMEOUT_S 120.0

struct ncclComm {
	char dir[112];
	int n, rank;
	long sseq[MAXR], rseq[MAXR];     /* messages sent to / received from each peer */
	long cseq;                        /* collectives so far */
};
typedef struct ncclComm *ncclComm_t;

struct op { int recv; void *buf; size_t bytes; int peer; struct ncclComm *c; };
static struct op queue[4 * MAXR + 16];
static int nqueue, depth;

static size_t
tsize(int dtype)
{
	switch(dtype)
	{
	case 0: case 1: return 1;
	case 2: case 3: case 7: return 4;
	case 4: case 5: case 8: return 8;
	case 6: return 2;
	}
	return 0;
}

static double
now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static int
put(const char *path, const void *buf, size_t bytes)
{
	char tmp[256];
	snprintf(tmp, sizeof(tmp), "%s.tmp%d", path, (int) getpid());
	FILE *f = fopen(tmp, "wb");
	if(!f) return -1;
	if(bytes && fwrite(buf, 1, bytes, f) != bytes) { fclose(f); return -1; }
	fclose(f);
	return rename(tmp, path);
}

static int
get(const char *path, void *buf, size_t bytes, int unlink_after)
{
	const double t0 = now();
	for(;;)
	{
		FILE *f = fopen(path, "rb");
		if(f)
		{
			size_t got = bytes ? fread(buf, 1, bytes, f) : 0;
			fclose(f);
			if(got != bytes) return -1;
			if(unlink_after) unlink(path);
			return 0;
		}
		if(now() - t0 > TIMEOUT_S) { fprintf(stderr, "fake_nccl: timed out waiting for %s\n", path); return -1; }
		usleep(50);
	}
}

static int
run(struct op *o)
{
	char path[256];
	struct ncclComm *c = o->c;
	if(!o->recv)
	{
		snprintf(path, sizeof(path), "%s/m_%d_%d_%ld", c->dir, c->rank, o->peer, c->sseq[o->peer]++);
		return put(path, o->buf, o->bytes);
	}
	snprintf(path, sizeof(path), "%s/m_%d_%d_%ld", c->dir, o->peer, c->rank, c->rseq[o->peer]++);
	return get(path, o->buf, o->bytes, 1);
}

static ncclResult_t
flush(void)
{
	int bad = 0;
	for(int pass = 0; pass < 2; pass++)          /* every send, then every receive */
		for(int i = 0; i < nqueue; i++)
			if(queue[i].recv == pass && run(&queue[i])) bad = 1;
	nqueue = 0;
	return bad ? ncclSystemError : ncclSuccess;
}

static ncclResult_t
post(int recv, void *buf, size_t count, int dtype, int peer, ncclComm_t c)
{
	if(!c || peer < 0 || peer >= c->n || !tsize(dtype)) return ncclInvalidArgument;
	if(nqueue >= (int) (sizeof(queue) / sizeof(queue[0]))) return ncclInvalidArgument;
	queue[nqueue++] = (struct op) { recv, buf, count * tsize(dtype), peer, c };
	return depth ? ncclSuccess : flush();
}

ncclResult_t
ncclGetUniqueId(ncclUniqueId *id)
{
	const char *tmp = getenv("TMPDIR");
	memset(id, 0, sizeof(*id));
	snprintf(id->internal, 100, "%s/cpic_b200_fake_nccl_%d_%ld", tmp && *tmp ? tmp : "/tmp", (int) getpid(), (long) time(NULL));
	if(mkdir(id->internal, 0700) && errno != EEXIST) return ncclSystemError;
	return ncclSuccess;
}

ncclResult_t
ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank)
{
	if(nranks < 1 || nranks > MAXR || rank < 0 || rank >= nranks) return ncclInvalidArgument;
	struct ncclComm *c = calloc(1, sizeof(*c));
	id.internal[sizeof(c->dir) - 1] = 0;
	strcpy(c->dir, id.internal);
	c->n = nranks;
	c->rank = rank;
	if(mkdir(c->dir, 0700) && errno != EEXIST) { free(c); return ncclSystemError; }
	*comm = c;
	return ncclSuccess;
}

ncclResult_t
ncclCommDestroy(ncclComm_t c)
{
	if(!c) return ncclSuccess;
	/* the last collective files of this rank; the directory goes when it is empty */
	char path[256];
	for(long k = c->cseq > 2 ? c->cseq - 2 : 0; k < c->cseq; k++)
	{
		snprintf(path, sizeof(path), "%s/c_%ld_%d", c->dir, k, c->rank);
		unlink(path);
	}
	rmdir(c->dir);
	free(c);
	return ncclSuccess;
}

ncclResult_t ncclSend(const void *buf, size_t count, int dtype, int peer, ncclComm_t c, void *stream)
{
	(void) stream;
	return post(0, (void *) buf, count, dtype, peer, c);
}

ncclResult_t ncclRecv(void *buf, size_t count, int dtype, int peer, ncclComm_t c, void *stream)
{
	(void) stream;
	return post(1, buf, count, dtype, peer, c);
}

ncclResult_t ncclGroupStart(void) { depth++; return ncclSuccess; }

ncclResult_t
ncclGroupEnd(void)
{
	if(depth <= 0) return ncclInvalidArgument;
	return --depth ? ncclSuccess : flush();
}

#define REDUCE(T) do { T *a = (T *) acc; const T *b = (const T *) tmp; \
	for(size_t i = 0; i < count; i++) \
		a[i] = op == 0 ? a[i] + b[i] : op == 1 ? a[i] * b[i] : op == 2 ? (a[i] > b[i] ? a[i] : b[i]) : (a[i] < b[i] ? a[i] : b[i]); \
	} while(0)

ncclResult_t
ncclAllReduce(const void *send, void *recv, size_t count, int dtype, int op, ncclComm_t c, void *stream)
{
	(void) stream;
	const size_t ts = tsize(dtype), bytes = count * ts;
	if(!c || !ts || op < 0 || op > 3 || !(dtype == 2 || dtype == 4 || dtype == 8)) return ncclInvalidArgument;
	char path[256];
	const long k = c->cseq++;
	snprintf(path, sizeof(path), "%s/c_%ld_%d", c->dir, k, c->rank);
	if(put(path, send, bytes)) return ncclSystemError;
	/* every rank has passed collective k-2 once its file of k-1 exists, which all ranks read below
	 * before writing k: the own file of k-2 can go */
	if(k >= 2)
	{
		snprintf(path, sizeof(path), "%s/c_%ld_%d", c->dir, k - 2, c->rank);
		unlink(path);
	}
	void *acc = malloc(bytes ? bytes : 1), *tmp = malloc(bytes ? bytes : 1);
	int bad = 0;
	for(int r = 0; r < c->n && !bad; r++)      /* rank order: the same result on every rank */
	{
		snprintf(path, sizeof(path), "%s/c_%ld_%d", c->dir, k, r);
		if(get(path, r == 0 ? acc : tmp, bytes, 0)) { bad = 1; break; }
		if(r == 0) continue;
		if(dtype == 2) REDUCE(int32_t);
		else if(dtype == 4) REDUCE(int64_t);
		else REDUCE(double);
	}
	if(!bad) memcpy(recv, acc, bytes);
	free(acc);
	free(tmp);
	return bad ? ncclSystemError : ncclSuccess;
}

const char *
ncclGetErrorString(ncclResult_t r)
{
	return r == ncclSuccess ? "no error" : r == ncclSystemError ? "fake_nccl: file exchange failed or timed out"
			: "fake_nccl: invalid argument";
}
