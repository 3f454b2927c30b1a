/* Reader for cpic's `.conf` grammar. See conf.h. */
#define _GNU_SOURCE
#include "conf.h"

#include <ctype.h>
#include <errno.h>
#include <libgen.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_INCLUDE_DEPTH 10

typedef struct parser {
	const char *s;       /* current position */
	const char *begin;
	int line;
	const char *incdir;
	int depth;
	char *err;
	size_t errlen;
	int failed;
} parser_t;

static void
fail(parser_t *p, const char *fmt, ...)
{
	va_list ap;
	int n;

	if(p->failed) return;
	p->failed = 1;
	if(!p->err || !p->errlen) return;
	n = snprintf(p->err, p->errlen, "line %d: ", p->line);
	if(n < 0 || (size_t) n >= p->errlen) return;
	va_start(ap, fmt);
	vsnprintf(p->err + n, p->errlen - (size_t) n, fmt, ap);
	va_end(ap);
}

static conf_node_t *
node_new(int type, const char *name, int line)
{
	conf_node_t *n = calloc(1, sizeof(*n));
	if(!n) abort();
	n->type = type;
	n->line = line;
	if(name) n->name = strdup(name);
	return n;
}

static void
node_add(conf_node_t *parent, conf_node_t *c)
{
	if(parent->nchild == parent->cap)
	{
		parent->cap = parent->cap ? parent->cap * 2 : 8;
		parent->child = realloc(parent->child,
				(size_t) parent->cap * sizeof(conf_node_t *));
		if(!parent->child) abort();
	}
	parent->child[parent->nchild++] = c;
	c->parent = parent;
}

void
conf_free(conf_node_t *n)
{
	int i;
	if(!n) return;
	for(i = 0; i < n->nchild; i++) conf_free(n->child[i]);
	free(n->child);
	free(n->name);
	free(n->sval);
	free(n);
}

static void
node_remove_member(conf_node_t *g, const char *name)
{
	int i, j;
	for(i = 0; i < g->nchild; i++)
	{
		if(g->child[i]->name && strcmp(g->child[i]->name, name) == 0)
		{
			conf_free(g->child[i]);
			for(j = i + 1; j < g->nchild; j++) g->child[j-1] = g->child[j];
			g->nchild--;
			return;
		}
	}
}

/* Skip blanks and the three comment styles */
static void
skip_ws(parser_t *p)
{
	for(;;)
	{
		char c = *p->s;
		if(c == '\n') { p->line++; p->s++; }
		else if(c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v') p->s++;
		else if(c == '#' || (c == '/' && p->s[1] == '/'))
		{
			while(*p->s && *p->s != '\n') p->s++;
		}
		else if(c == '/' && p->s[1] == '*')
		{
			p->s += 2;
			while(*p->s && !(p->s[0] == '*' && p->s[1] == '/'))
			{
				if(*p->s == '\n') p->line++;
				p->s++;
			}
			if(*p->s) p->s += 2;
			else { fail(p, "unterminated comment"); return; }
		}
		else return;
	}
}

static int
is_name_start(char c) { return isalpha((unsigned char) c) || c == '*' || c == '_'; }

static int
is_name_char(char c) { return isalnum((unsigned char) c) || c == '*' || c == '_' || c == '-'; }

static char *
parse_name(parser_t *p)
{
	const char *b = p->s;
	char *r;
	if(!is_name_start(*b)) { fail(p, "setting name expected near '%.12s'", b); return NULL; }
	while(is_name_char(*p->s)) p->s++;
	r = strndup(b, (size_t)(p->s - b));
	if(!r) abort();
	return r;
}

/* One quoted string piece, appended to *acc */
static int
parse_string_piece(parser_t *p, char **acc, size_t *len)
{
	p->s++; /* opening quote */
	for(;;)
	{
		char c = *p->s;
		char out;
		if(c == 0 || c == '\n') { fail(p, "unterminated string"); return -1; }
		if(c == '"') { p->s++; break; }
		if(c == '\\')
		{
			p->s++;
			switch(*p->s)
			{
				case 'n': out = '\n'; break;
				case 't': out = '\t'; break;
				case 'r': out = '\r'; break;
				case 'f': out = '\f'; break;
				case '\\': out = '\\'; break;
				case '"': out = '"'; break;
				case 'x':
				{
					char hex[3] = {0, 0, 0};
					if(isxdigit((unsigned char) p->s[1]) && isxdigit((unsigned char) p->s[2]))
					{
						hex[0] = p->s[1]; hex[1] = p->s[2];
						out = (char) strtol(hex, NULL, 16);
						p->s += 2;
					}
					else { fail(p, "bad \\x escape"); return -1; }
					break;
				}
				default: fail(p, "bad escape"); return -1;
			}
			p->s++;
		}
		else { out = c; p->s++; }
		*acc = realloc(*acc, *len + 2);
		if(!*acc) abort();
		(*acc)[(*len)++] = out;
		(*acc)[*len] = 0;
	}
	return 0;
}

static conf_node_t *parse_value(parser_t *p, const char *name);
static int parse_settings(parser_t *p, conf_node_t *group, int toplevel);

static conf_node_t *
parse_scalar(parser_t *p, const char *name)
{
	conf_node_t *n;
	char c = *p->s;
	int line = p->line;

	if(c == '"')
	{
		char *acc = strdup("");
		size_t len = 0;
		for(;;)
		{
			if(parse_string_piece(p, &acc, &len)) { free(acc); return NULL; }
			skip_ws(p); /* adjacent strings are concatenated */
			if(*p->s != '"') break;
		}
		n = node_new(CONF_STRING, name, line);
		n->sval = acc;
		return n;
	}

	if(isalpha((unsigned char) c))
	{
		if(strncasecmp(p->s, "true", 4) == 0 && !is_name_char(p->s[4]))
		{
			p->s += 4;
			n = node_new(CONF_BOOL, name, line); n->ival = 1; return n;
		}
		if(strncasecmp(p->s, "false", 5) == 0 && !is_name_char(p->s[5]))
		{
			p->s += 5;
			n = node_new(CONF_BOOL, name, line); n->ival = 0; return n;
		}
		fail(p, "value expected near '%.12s'", p->s);
		return NULL;
	}

	if(isdigit((unsigned char) c) || c == '+' || c == '-' || c == '.')
	{
		const char *b = p->s, *q = p->s;
		int is_float = 0, is_hex = 0;
		char *end;

		if(*q == '+' || *q == '-') q++;
		if(q[0] == '0' && (q[1] == 'x' || q[1] == 'X'))
		{
			is_hex = 1; q += 2;
			while(isxdigit((unsigned char) *q)) q++;
		}
		else
		{
			while(isdigit((unsigned char) *q)) q++;
			if(*q == '.') { is_float = 1; q++; while(isdigit((unsigned char) *q)) q++; }
			if(*q == 'e' || *q == 'E')
			{
				const char *r = q + 1;
				if(*r == '+' || *r == '-') r++;
				if(isdigit((unsigned char) *r))
				{
					is_float = 1;
					while(isdigit((unsigned char) *r)) r++;
					q = r;
				}
			}
		}

		if(q == b || (q == b + 1 && !isdigit((unsigned char) *b)))
		{
			fail(p, "number expected near '%.12s'", b);
			return NULL;
		}

		if(is_float)
		{
			n = node_new(CONF_FLOAT, name, line);
			n->fval = strtod(b, &end);
			p->s = q;
			return n;
		}

		errno = 0;
		{
			long long v = strtoll(b, &end, is_hex ? 16 : 10);
			int big = 0;
			p->s = q;
			if(*p->s == 'L')
			{
				big = 1;
				p->s++;
				if(*p->s == 'L') p->s++;
			}
			if(errno == ERANGE) { fail(p, "integer out of range"); return NULL; }
			if(v > INT_MAX || v < INT_MIN) big = 1;
			n = node_new(big ? CONF_INT64 : CONF_INT, name, line);
			n->ival = v;
			return n;
		}
	}

	fail(p, "value expected near '%.12s'", p->s);
	return NULL;
}

static conf_node_t *
parse_value(parser_t *p, const char *name)
{
	conf_node_t *n;
	int line = p->line;
	char c;

	skip_ws(p);
	c = *p->s;

	if(c == '{')
	{
		p->s++;
		n = node_new(CONF_GROUP, name, line);
		if(parse_settings(p, n, 0)) { conf_free(n); return NULL; }
		skip_ws(p);
		if(*p->s != '}') { fail(p, "'}' expected"); conf_free(n); return NULL; }
		p->s++;
		return n;
	}

	if(c == '(' || c == '[')
	{
		char close = (c == '(') ? ')' : ']';
		int is_list = (c == '(');
		p->s++;
		n = node_new(is_list ? CONF_LIST : CONF_ARRAY, name, line);
		for(;;)
		{
			conf_node_t *e;
			skip_ws(p);
			if(p->failed) { conf_free(n); return NULL; }
			if(*p->s == close) { p->s++; break; }
			if(is_list) e = parse_value(p, NULL);
			else e = parse_scalar(p, NULL);
			if(!e) { conf_free(n); return NULL; }
			if(!is_list && n->nchild > 0)
			{
				/* Arrays are homogeneous; int/int64 may mix (promoted) */
				int t0 = n->child[0]->type, t1 = e->type;
				int i0 = (t0 == CONF_INT || t0 == CONF_INT64);
				int i1 = (t1 == CONF_INT || t1 == CONF_INT64);
				if(t0 != t1 && !(i0 && i1))
				{
					fail(p, "mismatched element type in array");
					conf_free(e); conf_free(n); return NULL;
				}
			}
			node_add(n, e);
			skip_ws(p);
			if(*p->s == ',') { p->s++; continue; }
			if(*p->s == close) { p->s++; break; }
			fail(p, "',' or '%c' expected", close);
			conf_free(n);
			return NULL;
		}
		return n;
	}

	return parse_scalar(p, name);
}

static char *
read_whole_file(const char *path)
{
	FILE *f = fopen(path, "rb");
	char *buf;
	long sz;
	if(!f) return NULL;
	if(fseek(f, 0, SEEK_END)) { fclose(f); return NULL; }
	sz = ftell(f);
	if(sz < 0) { fclose(f); return NULL; }
	rewind(f);
	buf = malloc((size_t) sz + 1);
	if(!buf) abort();
	if(fread(buf, 1, (size_t) sz, f) != (size_t) sz) { free(buf); fclose(f); return NULL; }
	buf[sz] = 0;
	fclose(f);
	return buf;
}

static int
parse_include(parser_t *p, conf_node_t *group)
{
	char *acc = strdup("");
	size_t len = 0;
	char path[PATH_MAX];
	char *text;
	parser_t sub;
	int rc;

	skip_ws(p);
	if(*p->s != '"') { free(acc); fail(p, "@include needs a quoted path"); return -1; }
	if(parse_string_piece(p, &acc, &len)) { free(acc); return -1; }

	if(p->depth + 1 > MAX_INCLUDE_DEPTH)
	{
		free(acc); fail(p, "include nesting too deep"); return -1;
	}

	if(acc[0] == '/' || !p->incdir)
		snprintf(path, sizeof(path), "%s", acc);
	else
		snprintf(path, sizeof(path), "%s/%s", p->incdir, acc);

	text = read_whole_file(path);
	if(!text) { fail(p, "cannot open include file \"%s\"", path); free(acc); return -1; }
	free(acc);

	sub = *p;
	sub.s = sub.begin = text;
	sub.line = 1;
	sub.depth = p->depth + 1;
	rc = parse_settings(&sub, group, 1);
	if(!rc)
	{
		skip_ws(&sub);
		if(*sub.s) { fail(&sub, "unexpected '%c' in include file", *sub.s); rc = -1; }
	}
	p->failed = sub.failed;
	free(text);
	return rc;
}

static int
parse_settings(parser_t *p, conf_node_t *group, int toplevel)
{
	for(;;)
	{
		char *name;
		conf_node_t *v;

		skip_ws(p);
		if(p->failed) return -1;
		if(*p->s == 0) return toplevel ? 0 : (fail(p, "unexpected end of file"), -1);
		if(*p->s == '}') return toplevel ? (fail(p, "unexpected '}'"), -1) : 0;

		if(*p->s == '@')
		{
			if(strncmp(p->s, "@include", 8) != 0) { fail(p, "unknown directive"); return -1; }
			p->s += 8;
			if(parse_include(p, group)) return -1;
			continue;
		}

		name = parse_name(p);
		if(!name) return -1;
		skip_ws(p);
		if(*p->s != '=' && *p->s != ':')
		{
			fail(p, "'=' expected after \"%s\"", name);
			free(name);
			return -1;
		}
		p->s++;
		v = parse_value(p, name);
		if(!v) { free(name); return -1; }
		/* A later setting (e.g. after an @include) replaces an earlier one */
		node_remove_member(group, name);
		free(name);
		node_add(group, v);
		skip_ws(p);
		if(*p->s == ';' || *p->s == ',') p->s++;
	}
}

conf_node_t *
conf_parse_text(const char *text, const char *include_dir, char *errbuf, size_t errlen)
{
	parser_t p;
	conf_node_t *root;

	memset(&p, 0, sizeof(p));
	p.s = p.begin = text;
	p.line = 1;
	p.incdir = include_dir;
	p.err = errbuf;
	p.errlen = errlen;
	if(errbuf && errlen) errbuf[0] = 0;

	root = node_new(CONF_GROUP, NULL, 1);
	if(parse_settings(&p, root, 1) || p.failed)
	{
		conf_free(root);
		return NULL;
	}
	return root;
}

conf_node_t *
conf_parse_file(const char *path, const char *include_dir, char *errbuf, size_t errlen)
{
	char *text = read_whole_file(path);
	char *dup = NULL;
	conf_node_t *root;

	if(!text)
	{
		if(errbuf && errlen) snprintf(errbuf, errlen, "cannot open \"%s\"", path);
		return NULL;
	}
	if(!include_dir)
	{
		dup = strdup(path);
		include_dir = dirname(dup);
	}
	root = conf_parse_text(text, include_dir, errbuf, errlen);
	free(text);
	free(dup);
	return root;
}

conf_node_t *
conf_member(conf_node_t *g, const char *name)
{
	int i;
	if(!g || g->type != CONF_GROUP) return NULL;
	for(i = 0; i < g->nchild; i++)
		if(g->child[i]->name && strcmp(g->child[i]->name, name) == 0)
			return g->child[i];
	return NULL;
}

conf_node_t *
conf_elem(conf_node_t *a, int i)
{
	if(!a || i < 0 || i >= a->nchild) return NULL;
	if(a->type != CONF_GROUP && a->type != CONF_LIST && a->type != CONF_ARRAY) return NULL;
	return a->child[i];
}

int
conf_length(conf_node_t *a)
{
	if(!a) return 0;
	if(a->type != CONF_GROUP && a->type != CONF_LIST && a->type != CONF_ARRAY) return 0;
	return a->nchild;
}

conf_node_t *
conf_lookup(conf_node_t *from, const char *path)
{
	char buf[256];
	const char *s = path;
	conf_node_t *n = from;

	while(n && *s)
	{
		size_t k = strcspn(s, ".:/");
		if(k >= sizeof(buf)) return NULL;
		memcpy(buf, s, k);
		buf[k] = 0;
		if(k > 0)
		{
			if(buf[0] == '[')
				n = conf_elem(n, atoi(buf + 1));
			else
				n = conf_member(n, buf);
		}
		s += k;
		if(*s) s++;
	}
	return n;
}

int
conf_get_int(const conf_node_t *n, int *out)
{
	if(!n) return 0;
	if(n->type == CONF_INT || (n->type == CONF_INT64 && n->ival <= INT_MAX && n->ival >= INT_MIN))
	{
		*out = (int) n->ival;
		return 1;
	}
	return 0;
}

int
conf_get_int64(const conf_node_t *n, long long *out)
{
	if(!n) return 0;
	if(n->type == CONF_INT || n->type == CONF_INT64) { *out = n->ival; return 1; }
	return 0;
}

int
conf_get_float(const conf_node_t *n, double *out)
{
	if(!n || n->type != CONF_FLOAT) return 0;
	*out = n->fval;
	return 1;
}

int
conf_get_string(const conf_node_t *n, const char **out)
{
	if(!n || n->type != CONF_STRING) return 0;
	*out = n->sval;
	return 1;
}

int
conf_get_bool(const conf_node_t *n, int *out)
{
	if(!n || n->type != CONF_BOOL) return 0;
	*out = (int) n->ival;
	return 1;
}
