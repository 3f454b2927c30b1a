/* simt-check: the six asynchronous-copy primitives of kernels.cuh (mbarrier, TMA bulk and
 * tensor loads, cp.async groups) without PTX. Copies land at the matching wait -- the latest
 * moment the programming model allows -- and their destinations are poisoned until then.
 * Test infrastructure only (see simt.h). */
#ifndef SIMT_ASYNC_H
#define SIMT_ASYNC_H

static inline void mbar_init(uint64_t *bar, int count) { simt::bar_init(bar, count); }
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { simt::bar_expect_tx(bar, bytes); }

static inline int
mbar_wait(uint64_t *bar, uint32_t parity)
{
	for(int spin = 0; spin < 4096; spin++)
	{
		if(simt::bar_try_wait(bar, parity)) return 0;
		simt::yield();
	}
	return 1;
}

static inline void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) { simt::bar_tensor_load_2d(dst, map, c0, c1, bar); }
static inline void tma_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { simt::bar_bulk_load(dst, src, bytes, bar); }
static inline void cp_async8(void *dst, const void *src) { simt::async_copy8(dst, src); }
static inline void cp_async16(void *dst, const void *src)
{
	if((uintptr_t) dst % 16 || (uintptr_t) src % 16) simt::misuse("cp.async 16 needs 16-byte aligned addresses");
	simt::async_copy8(dst, src);
	simt::async_copy8((char *) dst + 8, (const char *) src + 8);
}
static inline void cp_async_commit() { simt::async_commit(); }
template <int N> static inline void cp_async_wait() { simt::async_wait(N); }

#endif
