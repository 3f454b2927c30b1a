/* simt-check runtime: fibers, warp collectives, late-completing asynchronous copies and the
 * slice of the CUDA runtime / cuFFT API that cpic_b200 uses. See simt.h: test infrastructure,
 * never part of the product. */
#include "simt.h"

#include <stdarg.h>
#include <sys/mman.h>
#include <time.h>

#include <complex>
#include <deque>
#include <unordered_map>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
/* AddressSanitizer has to be told about every stack switch */
#include <sanitizer/common_interface_defs.h>
#define SIMT_ASAN 1
#else
#define SIMT_ASAN 0
#endif

#if !defined(__x86_64__)
#error "simt-check's context switch is written for x86-64"
#endif

/* void simt_switch(void **save_sp, void *load_sp): callee-saved registers only */
extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size simt_switch,.-simt_switch\n");

namespace simt {

uint3 tid, bid;
dim3 bdim, gdim;
long long n_launches, n_switches;

/* SIMT_CHECK_EAGER=1: asynchronous copies land at issue instead of at the wait */
static const bool eager = getenv("SIMT_CHECK_EAGER") && atoi(getenv("SIMT_CHECK_EAGER"));
static const uint64_t POISON = 0x7ff8dead0000beefULL;   /* a NaN */
static const size_t STACK = 256 * 1024;
static const size_t DYN_SMEM = 228 * 1024;

struct Slot {
	bool active, completed;
	int op;
	unsigned mask, arrived;
	int readers;
	uint64_t val[32];
	int aux[32];
	uint64_t res[32];
};
struct Warp {
	unsigned alive;
	Slot slot[4];
};
struct Copy8 { void *dst; const void *src; };
struct BarOp {
	uint64_t *bar;
	void *dst;
	const void *src;     /* bulk: source; tensor: the map */
	uint32_t bytes;
	int tensor, c0, c1;
};
struct Bar { int init, pending; long long tx; int phase; };

enum { RUNNABLE, DONE };
struct Fiber {
	void *sp;
	char *stack;
	int state;
	int warp, lane;
	uint3 tid;
	int wkind;           /* 0: runnable, 1: waits for wslot->completed, 2: waits for the CTA barrier */
	Slot *wslot;
	uint64_t wcta;
	std::vector<Copy8> open;
	std::deque<std::vector<Copy8>> groups;
};

static std::vector<Fiber> fibers;
static std::vector<Warp> warps;
static Fiber *cur;
static void *sched_sp;
static const std::function<void()> *body;
static uint64_t cta_gen;
static int cta_arrived, cta_alive;
static unsigned char *smem_buf;
static std::vector<BarOp> bar_ops;
static std::unordered_map<uint64_t *, Bar> bars;
static bool aborting;
static char abort_msg[512];
static cudaError_t last_error = cudaSuccess;

#if SIMT_ASAN
static const void *sched_stack_bottom;
static size_t sched_stack_size;
#endif

/* fiber -> scheduler */
static void
to_scheduler()
{
#if SIMT_ASAN
	void *fake = NULL;
	Fiber *me = cur;
	__sanitizer_start_switch_fiber(me->state == DONE ? NULL : &fake, sched_stack_bottom, sched_stack_size);
	simt_switch(&me->sp, sched_sp);
	__sanitizer_finish_switch_fiber(fake, &sched_stack_bottom, &sched_stack_size);
#else
	simt_switch(&cur->sp, sched_sp);
#endif
}

/* scheduler -> fiber f */
static void
to_fiber(Fiber &f)
{
#if SIMT_ASAN
	void *fake = NULL;
	__sanitizer_start_switch_fiber(&fake, f.stack, STACK);
	simt_switch(&sched_sp, f.sp);
	__sanitizer_finish_switch_fiber(fake, NULL, NULL);
#else
	simt_switch(&sched_sp, f.sp);
#endif
}

static void
fatal(const char *fmt, ...)
{
	char msg[384];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(msg, sizeof(msg), fmt, ap);
	va_end(ap);
	snprintf(abort_msg, sizeof(abort_msg), "simt-check: %s [block (%u,%u) thread %u]", msg, bid.x, bid.y,
			cur ? cur->tid.x : 0u);
	fprintf(stderr, "%s\n", abort_msg);
	aborting = true;
	last_error = 700;
	if(cur) { cur->state = DONE; to_scheduler(); }
	abort();
}

void
yield()
{
	to_scheduler();
}

/* ------------------------------------------------------------- collectives */

static void
finish(Slot *s)
{
	const unsigned P = s->arrived;
	for(int l = 0; l < 32; l++)
	{
		if(!(P >> l & 1)) continue;
		uint64_t r = 0;
		switch(s->op)
		{
		case OP_BALLOT:
			for(int k = 0; k < 32; k++) if((P >> k & 1) && s->val[k]) r |= 1u << k;
			break;
		case OP_SHFL:
		{
			const int src = s->aux[l] & 31;
			r = (P >> src & 1) ? s->val[src] : s->val[l];
			break;
		}
		case OP_SHFL_UP:
		{
			const int src = l - s->aux[l];
			r = (src >= 0 && (P >> src & 1)) ? s->val[src] : s->val[l];
			break;
		}
		case OP_SHFL_DOWN:
		{
			const int src = l + s->aux[l];
			r = (src < 32 && (P >> src & 1)) ? s->val[src] : s->val[l];
			break;
		}
		case OP_MATCH_ANY:
			for(int k = 0; k < 32; k++) if((P >> k & 1) && s->val[k] == s->val[l]) r |= 1u << k;
			break;
		case OP_REDUCE_MAX:
		{
			int64_t m = INT64_MIN;
			for(int k = 0; k < 32; k++) if(P >> k & 1) m = std::max(m, (int64_t) s->val[k]);
			r = (uint64_t) m;
			break;
		}
		default:
			break;
		}
		s->res[l] = r;
	}
	s->completed = true;
	s->readers = __builtin_popcount(P);
}

static void
try_complete(Warp &w, Slot *s)
{
	const unsigned need = s->mask & w.alive;
	if(s->arrived && (s->arrived & need) == need) finish(s);
}

uint64_t
collective(int op, unsigned mask, uint64_t v, int aux)
{
	Fiber *f = cur;
	Warp &w = warps[f->warp];
	const unsigned bit = 1u << f->lane;
	if(!(mask & bit)) fatal("a lane called a *_sync primitive with a mask (%08x) that does not name it (lane %d)", mask, f->lane);
	Slot *s = NULL;
	for(Slot &c : w.slot)
		if(c.active && !c.completed && c.op == op && c.mask == mask && !(c.arrived & bit)) { s = &c; break; }
	if(!s)
	{
		for(Slot &c : w.slot) if(!c.active) { s = &c; break; }
		if(!s) fatal("more than %d warp collectives pending in one warp: lanes disagree on masks", (int) (sizeof(w.slot) / sizeof(w.slot[0])));
		s->active = true;
		s->completed = false;
		s->op = op;
		s->mask = mask;
		s->arrived = 0;
	}
	s->val[f->lane] = v;
	s->aux[f->lane] = aux;
	s->arrived |= bit;
	try_complete(w, s);
	if(!s->completed)
	{
		f->wkind = 1;
		f->wslot = s;
		to_scheduler();
		f->wkind = 0;
	}
	const uint64_t r = s->res[f->lane];
	if(--s->readers == 0) s->active = false;
	return r;
}

void
sync_cta()
{
	Fiber *f = cur;
	if(++cta_arrived == cta_alive)
	{
		cta_arrived = 0;
		cta_gen++;
		return;
	}
	f->wkind = 2;
	f->wcta = cta_gen;
	to_scheduler();
	f->wkind = 0;
}

/* ------------------------------------------------------ asynchronous copies */

static void
run_group(std::vector<Copy8> &g)
{
	for(const Copy8 &c : g) memcpy(c.dst, c.src, 8);
	g.clear();
}

void
async_copy8(void *dst, const void *src)
{
	if(eager) { memcpy(dst, src, 8); return; }
	memcpy(dst, &POISON, 8);
	cur->open.push_back({ dst, src });
}

void
misuse(const char *what)
{
	fatal("%s", what);
}

void
async_commit()
{
	cur->groups.emplace_back();
	cur->groups.back().swap(cur->open);
}

void
async_wait(int keep_newest)
{
	while((int) cur->groups.size() > keep_newest)
	{
		run_group(cur->groups.front());
		cur->groups.pop_front();
	}
}

static void
bar_check(Bar &b)
{
	if(b.pending == 0 && b.tx == 0)
	{
		b.phase ^= 1;
		b.pending = b.init;
	}
}

void
bar_init(uint64_t *bar, int count)
{
	bars[bar] = Bar{ count, count, 0, 0 };
}

static Bar &
bar_of(uint64_t *bar)
{
	auto it = bars.find(bar);
	if(it == bars.end()) fatal("mbarrier %p used before mbarrier.init", (void *) bar);
	return it->second;
}

void
bar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	Bar &b = bar_of(bar);
	if(b.pending <= 0) fatal("mbarrier.arrive beyond the barrier's arrival count");
	b.tx += bytes;
	b.pending--;
	bar_check(b);
}

static void
poison(void *dst, size_t bytes)
{
	for(size_t i = 0; i + 8 <= bytes; i += 8) memcpy((char *) dst + i, &POISON, 8);
}

void
bar_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	if(bytes % 16 || (uintptr_t) dst % 16 || (uintptr_t) src % 16)
		fatal("cp.async.bulk needs 16-byte aligned addresses and size (dst %p src %p bytes %u)", dst, src, bytes);
	bar_of(bar);
	poison(dst, bytes);
	bar_ops.push_back({ bar, dst, src, bytes, 0, 0, 0 });
	if(eager) memcpy(dst, src, bytes), bar_ops.back().src = dst;
}

void
bar_tensor_load_2d(void *dst, const void *map, int c0, int c1, uint64_t *bar)
{
	const CUtensorMap *m = (const CUtensorMap *) map;
	if((uintptr_t) dst % 128) fatal("TMA tensor destination %p is not 128-byte aligned", dst);
	bar_of(bar);
	const uint32_t bytes = m->box[0] * m->box[1] * m->elem;
	poison(dst, bytes);
	bar_ops.push_back({ bar, dst, map, bytes, 1, c0, c1 });
}

int
bar_try_wait(uint64_t *bar, uint32_t parity)
{
	Bar &b = bar_of(bar);
	/* the copies that complete on this barrier land now, as late as allowed */
	for(size_t i = 0; i < bar_ops.size();)
	{
		BarOp &o = bar_ops[i];
		if(o.bar != bar) { i++; continue; }
		if(!o.tensor) memcpy(o.dst, o.src, o.bytes);
		else
		{
			const CUtensorMap *m = (const CUtensorMap *) o.src;
			for(uint32_t r = 0; r < m->box[1]; r++)
				for(uint32_t c = 0; c < m->box[0]; c++)
				{
					const long long gx = (long long) o.c0 + c, gy = (long long) o.c1 + r;
					uint64_t v = 0;     /* out-of-range elements are zero-filled */
					if(gx >= 0 && gy >= 0 && (uint64_t) gx < m->dims[0] && (uint64_t) gy < m->dims[1])
						memcpy(&v, (const char *) m->base + (size_t) gy * m->row_bytes + (size_t) gx * m->elem, 8);
					memcpy((char *) o.dst + ((size_t) r * m->box[0] + c) * 8, &v, 8);
				}
		}
		b.tx -= o.bytes;
		bar_ops[i] = bar_ops.back();
		bar_ops.pop_back();
	}
	bar_check(b);
	return (uint32_t) b.phase != parity;
}

/* ------------------------------------------------------------------ launch */

static void
fiber_exit()
{
	Fiber *f = cur;
	if(!f->open.empty() || !f->groups.empty())
	{
		/* copies still in flight at exit: complete them (the kernels wait for group 0 before leaving) */
		async_commit();
		async_wait(0);
	}
	f->state = DONE;
	Warp &w = warps[f->warp];
	w.alive &= ~(1u << f->lane);
	for(Slot &c : w.slot) if(c.active && !c.completed) try_complete(w, &c);
	cta_alive--;
	if(cta_arrived > 0 && cta_arrived == cta_alive)
	{
		cta_arrived = 0;
		cta_gen++;
	}
	to_scheduler();
	abort();    /* a finished fiber is never resumed */
}

static void
fiber_entry()
{
#if SIMT_ASAN
	__sanitizer_finish_switch_fiber(NULL, &sched_stack_bottom, &sched_stack_size);
#endif
	(*body)();
	fiber_exit();
}

static void
prepare(Fiber &f)
{
	if(!f.stack)
	{
		f.stack = (char *) mmap(NULL, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if(f.stack == MAP_FAILED) { perror("simt-check: mmap"); abort(); }
	}
	uintptr_t top = ((uintptr_t) f.stack + STACK) & ~(uintptr_t) 15;
	void **sp = (void **) top;
	*--sp = NULL;                      /* return address of fiber_entry: never used */
	*--sp = (void *) fiber_entry;      /* `ret` of the first switch jumps here */
	for(int i = 0; i < 6; i++) *--sp = NULL;
	f.sp = sp;
	f.state = RUNNABLE;
	f.wkind = 0;
	f.open.clear();
	f.groups.clear();
}

unsigned char *
dyn_smem()
{
	return smem_buf;
}

void
launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &fn)
{
	n_launches++;
	if(aborting) return;
	const int n = (int) (block.x * block.y * block.z);
	if(n <= 0 || n > 1024 || smem > 227 * 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535)
	{
		last_error = 9;     /* invalid configuration */
		return;
	}
	if(!smem_buf)
	{
		if(posix_memalign((void **) &smem_buf, 1024, DYN_SMEM)) abort();
	}
	if((int) fibers.size() < n) fibers.resize(n);
	const int nw = (n + 31) / 32;
	warps.resize(nw);
	body = &fn;
	gdim = grid;
	bdim = block;
	for(unsigned bz = 0; bz < grid.z && !aborting; bz++)
		for(unsigned by = 0; by < grid.y && !aborting; by++)
			for(unsigned bx = 0; bx < grid.x && !aborting; bx++)
			{
				bid = { bx, by, bz };
				poison(smem_buf, smem);
				bars.clear();
				bar_ops.clear();
				cta_gen = 0;
				cta_arrived = 0;
				cta_alive = n;
				for(int w = 0; w < nw; w++)
				{
					const int lanes = std::min(32, n - 32 * w);
					warps[w].alive = lanes == 32 ? 0xffffffffu : (1u << lanes) - 1;
					for(Slot &c : warps[w].slot) c.active = false;
				}
				for(int t = 0; t < n; t++)
				{
					Fiber &f = fibers[t];
					prepare(f);
					f.warp = t / 32;
					f.lane = t % 32;
					f.tid = { (unsigned) t % block.x, (unsigned) (t / block.x) % block.y, (unsigned) (t / (block.x * block.y)) };
				}
				int done = 0;
				while(done < n && !aborting)
				{
					bool progress = false;
					for(int t = 0; t < n && !aborting; t++)
					{
						Fiber &f = fibers[t];
						if(f.state == DONE) continue;
						if(f.wkind == 1 && !f.wslot->completed) continue;
						if(f.wkind == 2 && f.wcta == cta_gen) continue;
						cur = &f;
						tid = f.tid;
						n_switches++;
						to_fiber(f);
						if(f.state == DONE) done++;
						progress = true;
					}
					if(!progress && done < n && !aborting)
					{
						int stuck = 0;
						for(int t = 0; t < n; t++) if(fibers[t].state != DONE) { stuck = t; break; }
						cur = NULL;
						snprintf(abort_msg, sizeof(abort_msg),
								"simt-check: deadlock in block (%u,%u): thread %d waits for a %s that the other threads never reach",
								bx, by, stuck, fibers[stuck].wkind == 2 ? "__syncthreads" : "warp collective");
						fprintf(stderr, "%s\n", abort_msg);
						aborting = true;
						last_error = 700;
					}
				}
			}
	cur = NULL;
}

}  /* namespace simt */

/* Clears a reported fault (deadlock, misuse of a primitive) so that the next launch runs */
extern "C" void
simt_check_reset(void)
{
	simt::aborting = false;
	simt::last_error = cudaSuccess;
	simt::abort_msg[0] = 0;
}

extern "C" long long simt_check_launches(void) { return simt::n_launches; }
extern "C" long long simt_check_switches(void) { return simt::n_switches; }

/* ------------------------------------------------------------ runtime API */

cudaError_t
simt_malloc(void **p, size_t n)
{
	void *q = NULL;
	if(posix_memalign(&q, 256, n ? n : 1)) return cudaErrorMemoryAllocation;
	memset(q, 0xff, n);          /* fresh device memory holds garbage: NaNs and -1 here */
	*p = q;
	return cudaSuccess;
}

cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }

cudaError_t
cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
	for(size_t r = 0; r < height; r++) memmove((char *) d + r * dpitch, (const char *) s + r * spitch, width);
	return cudaSuccess;
}

cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
/* every rank of a multi-process test names its own "device" */
static int cur_device;
cudaError_t cudaSetDevice(int d) { if(d < 0 || d >= 16) return cudaErrorInvalidValue; cur_device = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = cur_device; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }

cudaError_t
cudaGetLastError(void)
{
	cudaError_t e = simt::last_error;
	if(!simt::aborting) simt::last_error = cudaSuccess;     /* a fault is sticky, like a device fault */
	return e;
}

const char *
cudaGetErrorString(cudaError_t e)
{
	if(e == cudaSuccess) return "no error";
	if(e == 700) return simt::abort_msg;
	if(e == 9) return "invalid configuration argument";
	if(e == cudaErrorMemoryAllocation) return "out of memory";
	return "invalid value";
}

struct simt_stream_ { int unused; };
struct simt_event_ { double t; };

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new simt_stream_(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return simt::aborting ? 700 : cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return simt::aborting ? 700 : cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new simt_event_(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new simt_event_(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }

cudaError_t
cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	e->t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
	return cudaSuccess;
}

cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (b->t - a->t); return cudaSuccess; }

/* The checks the driver makes on a tiled tensor map (alignment and box limits) */
static CUresult
encode_tiled(CUtensorMap *map, CUtensorMapDataType type, cuuint32_t rank, void *base, const cuuint64_t *dims,
		const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *estr, CUtensorMapInterleave,
		CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)
{
	if(type != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 || rank != 2) return CUDA_ERROR_INVALID_VALUE;
	if((uintptr_t) base % 16 || strides[0] % 16) return CUDA_ERROR_INVALID_VALUE;
	if(box[0] == 0 || box[0] > 256 || box[1] == 0 || box[1] > 256) return CUDA_ERROR_INVALID_VALUE;
	if((box[0] * 8) % 16) return CUDA_ERROR_INVALID_VALUE;
	if(estr[0] != 1 || estr[1] != 1 || dims[0] == 0 || dims[1] == 0) return CUDA_ERROR_INVALID_VALUE;
	memset(map, 0, sizeof(*map));
	map->base = base;
	map->dims[0] = dims[0];
	map->dims[1] = dims[1];
	map->row_bytes = strides[0];
	map->box[0] = box[0];
	map->box[1] = box[1];
	map->elem = 8;
	return CUDA_SUCCESS;
}

cudaError_t
cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *res)
{
	*fn = NULL;
	if(!strcmp(name, "cuTensorMapEncodeTiled")) *fn = (void *) encode_tiled;
	if(res) *res = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
	return cudaSuccess;
}

/* ------------------------------------------------------------------ cuFFT */

typedef std::complex<double> cplx;

struct Plan {
	int rank, n[2];
	int inembed[2], onembed[2];
	int istride, idist, ostride, odist, batch;
	cufftType type;
	bool live;
};
static std::vector<Plan> plans;

/* exp(sign * 2 pi i k / n) with the exact symmetries of the unit circle (the angle is reduced to
 * the first octant before sin/cos are taken, in long double): w[n-k] is exactly conj(w[k]),
 * w[n/2-k] exactly its mirror. Without them a point charge's potential is not symmetric to the
 * last bit and the self force of a lone particle drifts (the reference's constant-speed test). */
static cplx
unit_root(long long k, long long n, int sign)
{
	k %= n;
	if(k < 0) k += n;
	const long long o = 8 * k / n;                     /* octant 0..7 */
	const long double tau = 6.283185307179586476925286766559005768L;
	long double c, s;
	if(o % 2 == 0)
	{
		const long double a = tau * (long double) (8 * k - o * n) / (long double) (8 * n);
		c = cosl(a); s = sinl(a);
	}
	else
	{
		const long double a = tau * (long double) ((o + 1) * n - 8 * k) / (long double) (8 * n);
		c = sinl(a); s = cosl(a);                      /* mirrored at the octant's end */
	}
	double x, y;
	switch(o)
	{
	case 0: x = (double) c; y = (double) s; break;
	case 1: x = (double) c; y = (double) s; break;
	case 2: x = -(double) s; y = (double) c; break;
	case 3: x = -(double) s; y = (double) c; break;
	case 4: x = -(double) c; y = -(double) s; break;
	case 5: x = -(double) c; y = -(double) s; break;
	case 6: x = (double) s; y = -(double) c; break;
	default: x = (double) s; y = -(double) c; break;
	}
	return cplx(x, sign * y);
}

/* in place, unnormalised; sign -1 forward, +1 inverse */
static void
fft1(cplx *a, int n, int stride, int sign)
{
	std::vector<cplx> t(n);
	for(int i = 0; i < n; i++) t[i] = a[(size_t) i * stride];
	if((n & (n - 1)) == 0)
	{
		for(int i = 1, j = 0; i < n; i++)
		{
			int bit = n >> 1;
			for(; j & bit; bit >>= 1) j ^= bit;
			j ^= bit;
			if(i < j) std::swap(t[i], t[j]);
		}
		for(int len = 2; len <= n; len <<= 1)
		{
			std::vector<cplx> w(len / 2);
			for(int k = 0; k < len / 2; k++) w[k] = unit_root(k, len, sign);
			for(int i = 0; i < n; i += len)
				for(int k = 0; k < len / 2; k++)
				{
					const cplx u = t[i + k], v = t[i + k + len / 2] * w[k];
					t[i + k] = u + v;
					t[i + k + len / 2] = u - v;
				}
		}
	}
	else
	{
		std::vector<cplx> o(n);
		for(int k = 0; k < n; k++)
		{
			cplx s = 0;
			for(int j = 0; j < n; j++) s += t[j] * unit_root((long long) j * k, n, sign);
			o[k] = s;
		}
		t.swap(o);
	}
	for(int i = 0; i < n; i++) a[(size_t) i * stride] = t[i];
}

cufftResult
cufftPlanMany(cufftHandle *plan, int rank, int *n, int *inembed, int istride, int idist, int *onembed, int ostride,
		int odist, cufftType type, int batch)
{
	if(rank < 1 || rank > 2) return CUFFT_NOT_SUPPORTED;
	Plan p = {};
	p.rank = rank;
	for(int i = 0; i < rank; i++)
	{
		p.n[i] = n[i];
		p.inembed[i] = inembed ? inembed[i] : n[i];
		p.onembed[i] = onembed ? onembed[i] : n[i];
	}
	p.istride = inembed ? istride : 1;
	p.ostride = onembed ? ostride : 1;
	p.idist = idist;
	p.odist = odist;
	p.batch = batch;
	p.type = type;
	p.live = true;
	plans.push_back(p);
	*plan = (int) plans.size() - 1;
	return CUFFT_SUCCESS;
}

cufftResult cufftSetStream(cufftHandle, cudaStream_t) { return CUFFT_SUCCESS; }

cufftResult
cufftDestroy(cufftHandle h)
{
	if(h < 0 || h >= (int) plans.size() || !plans[h].live) return CUFFT_INVALID_PLAN;
	plans[h].live = false;
	return CUFFT_SUCCESS;
}

static Plan *
plan_of(cufftHandle h, cufftType type)
{
	if(h < 0 || h >= (int) plans.size() || !plans[h].live || plans[h].type != type) return NULL;
	return &plans[h];
}

cufftResult
cufftExecD2Z(cufftHandle h, cufftDoubleReal *in, cufftDoubleComplex *out)
{
	Plan *p = plan_of(h, CUFFT_D2Z);
	if(!p) return CUFFT_INVALID_PLAN;
	for(int b = 0; b < p->batch; b++)
	{
		const double *src = in + (size_t) b * p->idist;
		cplx *dst = (cplx *) out + (size_t) b * p->odist;
		if(p->rank == 1)
		{
			const int n = p->n[0], nc = n / 2 + 1;
			std::vector<cplx> row(n);
			for(int i = 0; i < n; i++) row[i] = src[(size_t) i * p->istride];
			fft1(row.data(), n, 1, -1);
			for(int k = 0; k < nc; k++) dst[(size_t) k * p->ostride] = row[k];
			continue;
		}
		const int ny = p->n[0], nx = p->n[1], nc = nx / 2 + 1;
		std::vector<cplx> row(nx);
		for(int y = 0; y < ny; y++)
		{
			for(int x = 0; x < nx; x++) row[x] = src[((size_t) y * p->inembed[1] + x) * p->istride];
			fft1(row.data(), nx, 1, -1);
			for(int k = 0; k < nc; k++) dst[((size_t) y * p->onembed[1] + k) * p->ostride] = row[k];
		}
		for(int k = 0; k < nc; k++) fft1(dst + (size_t) k * p->ostride, ny, p->onembed[1] * p->ostride, -1);
	}
	return CUFFT_SUCCESS;
}

cufftResult
cufftExecZ2D(cufftHandle h, cufftDoubleComplex *in, cufftDoubleReal *out)
{
	Plan *p = plan_of(h, CUFFT_Z2D);
	if(!p) return CUFFT_INVALID_PLAN;
	for(int b = 0; b < p->batch; b++)
	{
		cplx *src = (cplx *) in + (size_t) b * p->idist;
		double *dst = out + (size_t) b * p->odist;
		if(p->rank == 1)
		{
			const int n = p->n[0], nc = n / 2 + 1;
			std::vector<cplx> row(n);
			for(int k = 0; k < nc; k++) row[k] = src[(size_t) k * p->istride];
			for(int k = nc; k < n; k++) row[k] = std::conj(row[n - k]);
			fft1(row.data(), n, 1, +1);
			for(int i = 0; i < n; i++) dst[(size_t) i * p->ostride] = row[i].real();
			continue;
		}
		const int ny = p->n[0], nx = p->n[1], nc = nx / 2 + 1;
		/* like cuFFT, the input is used as work space */
		for(int k = 0; k < nc; k++) fft1(src + (size_t) k * p->istride, ny, p->inembed[1] * p->istride, +1);
		std::vector<cplx> row(nx);
		for(int y = 0; y < ny; y++)
		{
			for(int k = 0; k < nc; k++) row[k] = src[((size_t) y * p->inembed[1] + k) * p->istride];
			for(int k = nc; k < nx; k++) row[k] = std::conj(row[nx - k]);
			fft1(row.data(), nx, 1, +1);
			for(int x = 0; x < nx; x++) dst[((size_t) y * p->onembed[1] + x) * p->ostride] = row[x].real();
		}
	}
	return CUFFT_SUCCESS;
}

cufftResult
cufftExecZ2Z(cufftHandle h, cufftDoubleComplex *in, cufftDoubleComplex *out, int dir)
{
	Plan *p = plan_of(h, CUFFT_Z2Z);
	if(!p) return CUFFT_INVALID_PLAN;
	if(p->rank != 1) return CUFFT_NOT_SUPPORTED;
	const int n = p->n[0];
	std::vector<cplx> row(n);
	for(int b = 0; b < p->batch; b++)
	{
		const cplx *src = (const cplx *) in + (size_t) b * p->idist;
		cplx *dst = (cplx *) out + (size_t) b * p->odist;
		for(int i = 0; i < n; i++) row[i] = src[(size_t) i * p->istride];
		fft1(row.data(), n, 1, dir < 0 ? -1 : +1);
		for(int i = 0; i < n; i++) dst[(size_t) i * p->ostride] = row[i];
	}
	return CUFFT_SUCCESS;
}
