/* ORACLE TEST INFRASTRUCTURE — not product code.
 *
 * Force-included (gcc -include) when building the "accumulate-correct" variant of the
 * reference's interpolate.c. The reference's vmat_add_xy (src/simd_avx2.h:245-249)
 * gathers four old values, adds, then stores lane by lane, so two lanes of one pack
 * that hit the same node keep only the last lane's deposit.
 * This header pulls in the reference's own definitions first (they are `#pragma once`)
 * and then redirects later *uses* of vmat_add_xy to a lane-serial accumulate. No
 * reference source file is modified or copied. */
#ifndef ORACLE_FIX_F1_H
#define ORACLE_FIX_F1_H

#include "simd.h"

static inline void
oracle_vmat_add_xy_serial(mat_t *m, vi64 ix, vi64 iy, vf64 x)
{
	size_t iv;
	vi64 idx = vmat_index_xy(m, ix, iy);
	for(iv = 0; iv < MAX_VEC; iv++)
		m->data[idx[iv]] += x[iv];
}

#define vmat_add_xy oracle_vmat_add_xy_serial

#endif
