/* Parser for cpic's `.conf` files (the libconfig grammar subset cpic uses).
 *
 * cpic reads its configuration through libconfig (reference: src/cpic.c:150-164,
 * src/sim.c:38-85, src/config.c, src/specie.c:15-36, src/particle.c:39-41,110-120,
 * src/output.c:67-110). libconfig is not available on the B200 image, so this is a
 * from-scratch reader of the same grammar: `#`, `//` and C comments, `name = value`
 * or `name : value` with optional `;` / `,` terminators, groups `{}`, lists `()`,
 * arrays `[]`, strings, booleans, ints (decimal/hex, `L` suffix), floats, and
 * `@include "file"` resolved against an include directory.
 */
#ifndef CPIC_B200_CONF_H
#define CPIC_B200_CONF_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum conf_type {
	CONF_NONE = 0,
	CONF_GROUP,
	CONF_INT,
	CONF_INT64,
	CONF_FLOAT,
	CONF_STRING,
	CONF_BOOL,
	CONF_ARRAY,
	CONF_LIST
};

typedef struct conf_node {
	int type;
	char *name;              /* NULL for list/array elements and the root */
	struct conf_node *parent;
	struct conf_node **child;
	int nchild, cap;
	long long ival;          /* CONF_INT, CONF_INT64, CONF_BOOL */
	double fval;             /* CONF_FLOAT */
	char *sval;              /* CONF_STRING */
	int line;
} conf_node_t;

/* Parse a file. `include_dir` may be NULL (then the file's own directory is used).
 * Returns the root group, or NULL with a message in `errbuf`. */
conf_node_t *conf_parse_file(const char *path, const char *include_dir,
		char *errbuf, size_t errlen);

/* Parse a NUL-terminated text. `include_dir` is used for @include (may be NULL). */
conf_node_t *conf_parse_text(const char *text, const char *include_dir,
		char *errbuf, size_t errlen);

void conf_free(conf_node_t *root);

/* Dotted path lookup from a group ("simulation.sampling_period.energy"). */
conf_node_t *conf_lookup(conf_node_t *from, const char *path);
conf_node_t *conf_member(conf_node_t *group, const char *name);
conf_node_t *conf_elem(conf_node_t *agg, int i);
int conf_length(conf_node_t *agg);

/* Typed getters with libconfig's (non auto-converting) rules:
 *   int    <- INT (and INT64 when it fits)
 *   int64  <- INT, INT64
 *   float  <- FLOAT only
 * Return 1 on success, 0 on type mismatch / NULL node. */
int conf_get_int(const conf_node_t *n, int *out);
int conf_get_int64(const conf_node_t *n, long long *out);
int conf_get_float(const conf_node_t *n, double *out);
int conf_get_string(const conf_node_t *n, const char **out);
int conf_get_bool(const conf_node_t *n, int *out);

#ifdef __cplusplus
}
#endif

#endif
