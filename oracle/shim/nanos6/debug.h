/* ORACLE TEST INFRASTRUCTURE. Stand-in for <nanos6/debug.h> (src/solver.c:21, :292). */
#ifndef ORACLE_SHIM_NANOS6_DEBUG_H
#define ORACLE_SHIM_NANOS6_DEBUG_H
static inline unsigned int nanos6_get_num_cpus(void) { return 1; }
#endif
