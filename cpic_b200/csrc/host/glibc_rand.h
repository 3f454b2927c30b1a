/* glibc's rand() as an explicit linear recurrence, so that the reference's initial conditions
 * (src/particle.c:17-21, 69-73: four rand() calls per particle, srand(seed + rank) per process,
 * src/sim.c:153) can be drawn on the device from any point of the stream.
 *
 * glibc (stdlib/random_r.c, TYPE_3, the default state of srand/rand): 34 seed words
 *   r[0] = seed (1 if 0), r[i] = 16807 * r[i-1] mod 2147483647 (i = 1..30, Schrage's form),
 *   r[i] = r[i-31] (i = 31..33), then r[i] = r[i-31] + r[i-3] mod 2^32 for all i >= 34;
 * the first 310 values after the seeding are discarded and call k (k = 0, 1, ...) returns
 * r[344 + k] >> 1. Being linear over Z/2^32, the recurrence can be advanced by n steps with the
 * n-th power of its 31 x 31 companion matrix (powers of two are cached). */
#ifndef CPIC_B200_GLIBC_RAND_H
#define CPIC_B200_GLIBC_RAND_H

#include <stdint.h>
#include <string.h>

#define GLIBC_RAND_DEG 31
#define GLIBC_RAND_SEP 3

/* state: the last 31 values, oldest first, such that the next call returns (st[0] + st[28]) >> 1 */
static inline void
glibc_rand_seed(uint32_t seed, uint32_t st[GLIBC_RAND_DEG])
{
	uint32_t r[344];
	if(seed == 0) seed = 1;
	r[0] = seed;
	for(int i = 1; i < 31; i++)
	{
		/* 16807 * r[i-1] % 2147483647 without overflow, as random_r.c does it */
		const int32_t hi = (int32_t) r[i - 1] / 127773, lo = (int32_t) r[i - 1] % 127773;
		int32_t word = 16807 * lo - 2836 * hi;
		if(word < 0) word += 2147483647;
		r[i] = (uint32_t) word;
	}
	for(int i = 31; i < 34; i++) r[i] = r[i - 31];
	for(int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
	memcpy(st, r + 344 - 31, 31 * sizeof(uint32_t));
}

static inline uint32_t
glibc_rand_next(uint32_t st[GLIBC_RAND_DEG])
{
	const uint32_t v = st[0] + st[GLIBC_RAND_DEG - GLIBC_RAND_SEP];
	memmove(st, st + 1, 30 * sizeof(uint32_t));
	st[30] = v;
	return v >> 1;
}

struct GlibcRandJump {
	/* pw[k] = M^(2^k), row-major, M the one-step matrix on the state vector */
	uint32_t pw[48][GLIBC_RAND_DEG][GLIBC_RAND_DEG];
	GlibcRandJump()
	{
		memset(pw, 0, sizeof(pw));
		for(int i = 0; i < 30; i++) pw[0][i][i + 1] = 1;      /* shift */
		pw[0][30][0] = 1;
		pw[0][30][GLIBC_RAND_DEG - GLIBC_RAND_SEP] = 1;        /* new = st[0] + st[28] */
		for(int k = 1; k < 48; k++) mul(pw[k - 1], pw[k - 1], pw[k]);
	}
	static void mul(const uint32_t a[31][31], const uint32_t b[31][31], uint32_t c[31][31])
	{
		uint32_t t[31][31];
		for(int i = 0; i < 31; i++)
			for(int j = 0; j < 31; j++)
			{
				uint32_t s = 0;
				for(int k = 0; k < 31; k++) s += a[i][k] * b[k][j];
				t[i][j] = s;
			}
		memcpy(c, t, sizeof(t));
	}
	static void apply(const uint32_t m[31][31], uint32_t st[31])
	{
		uint32_t t[31];
		for(int i = 0; i < 31; i++)
		{
			uint32_t s = 0;
			for(int k = 0; k < 31; k++) s += m[i][k] * st[k];
			t[i] = s;
		}
		memcpy(st, t, sizeof(t));
	}
	/* the matrix of n steps */
	void power(uint64_t n, uint32_t out[31][31]) const
	{
		memset(out, 0, 31 * 31 * sizeof(uint32_t));
		for(int i = 0; i < 31; i++) out[i][i] = 1;
		for(int k = 0; k < 48 && n; k++, n >>= 1)
			if(n & 1) mul(pw[k], out, out);
	}
	/* st advanced by n calls */
	void jump(uint32_t st[31], uint64_t n) const
	{
		for(int k = 0; k < 48 && n; k++, n >>= 1)
			if(n & 1) apply(pw[k], st);
	}
};

#endif
