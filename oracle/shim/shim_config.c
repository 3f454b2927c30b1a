/* ORACLE TEST INFRASTRUCTURE — not product code.
 * libconfig entry points used by the reference, over the repo's `.conf` reader. */
#define _GNU_SOURCE
#include "libconfig.h"
#include "conf.h"

#include <stdlib.h>
#include <string.h>

void config_init(config_t *c) { memset(c, 0, sizeof(*c)); }

void config_destroy(config_t *c)
{
	conf_free(c->root);
	free(c->include_dir);
	memset(c, 0, sizeof(*c));
}

void config_set_include_dir(config_t *c, const char *dir)
{
	free(c->include_dir);
	c->include_dir = dir ? strdup(dir) : NULL;
}

static int finish(config_t *c, const char *file)
{
	if(c->root) return CONFIG_TRUE;
	c->error_file = file;
	c->error_line = 0;
	if(strncmp(c->error_text, "line ", 5) == 0) c->error_line = atoi(c->error_text + 5);
	return CONFIG_FALSE;
}

int config_read_file(config_t *c, const char *filename)
{
	conf_free(c->root);
	c->root = conf_parse_file(filename, c->include_dir, c->error_text, sizeof(c->error_text));
	return finish(c, filename);
}

int config_read_string(config_t *c, const char *str)
{
	conf_free(c->root);
	c->root = conf_parse_text(str, c->include_dir, c->error_text, sizeof(c->error_text));
	return finish(c, "<string>");
}

int config_read(config_t *c, FILE *f)
{
	size_t cap = 1 << 16, len = 0, n;
	char *buf = malloc(cap);
	int rc;
	if(!buf) abort();
	while((n = fread(buf + len, 1, cap - len - 1, f)) > 0)
	{
		len += n;
		if(len + 1 >= cap) { cap *= 2; buf = realloc(buf, cap); if(!buf) abort(); }
	}
	buf[len] = 0;
	rc = config_read_string(c, buf);
	free(buf);
	return rc;
}

const char *config_error_text(const config_t *c) { return c->error_text; }
const char *config_error_file(const config_t *c) { return c->error_file; }
int config_error_line(const config_t *c) { return c->error_line; }

config_setting_t *config_lookup(const config_t *c, const char *path)
{
	return conf_lookup(c->root, path);
}

int config_lookup_int(const config_t *c, const char *path, int *v)
{ return conf_get_int(conf_lookup(c->root, path), v); }
int config_lookup_int64(const config_t *c, const char *path, long long *v)
{ return conf_get_int64(conf_lookup(c->root, path), v); }
int config_lookup_float(const config_t *c, const char *path, double *v)
{ return conf_get_float(conf_lookup(c->root, path), v); }
int config_lookup_string(const config_t *c, const char *path, const char **v)
{ return conf_get_string(conf_lookup(c->root, path), v); }
int config_lookup_bool(const config_t *c, const char *path, int *v)
{ return conf_get_bool(conf_lookup(c->root, path), v); }

int config_setting_type(const config_setting_t *s) { return s ? s->type : CONFIG_TYPE_NONE; }
int config_setting_length(const config_setting_t *s) { return conf_length((conf_node_t *) s); }
const char *config_setting_name(const config_setting_t *s) { return s ? s->name : NULL; }
unsigned int config_setting_source_line(const config_setting_t *s) { return s ? (unsigned) s->line : 0; }

config_setting_t *config_setting_get_elem(const config_setting_t *s, unsigned int idx)
{ return conf_elem((conf_node_t *) s, (int) idx); }
config_setting_t *config_setting_get_member(const config_setting_t *s, const char *name)
{ return conf_member((conf_node_t *) s, name); }

int config_setting_get_int(const config_setting_t *s)
{ int v = 0; conf_get_int(s, &v); return v; }
long long config_setting_get_int64(const config_setting_t *s)
{ long long v = 0; conf_get_int64(s, &v); return v; }
double config_setting_get_float(const config_setting_t *s)
{ double v = 0; conf_get_float(s, &v); return v; }
const char *config_setting_get_string(const config_setting_t *s)
{ const char *v = NULL; conf_get_string(s, &v); return v; }

int config_setting_get_int_elem(const config_setting_t *s, int idx)
{ return config_setting_get_int(conf_elem((conf_node_t *) s, idx)); }
long long config_setting_get_int64_elem(const config_setting_t *s, int idx)
{ return config_setting_get_int64(conf_elem((conf_node_t *) s, idx)); }
double config_setting_get_float_elem(const config_setting_t *s, int idx)
{ return config_setting_get_float(conf_elem((conf_node_t *) s, idx)); }

int config_setting_lookup_int(const config_setting_t *s, const char *name, int *v)
{ return conf_get_int(conf_member((conf_node_t *) s, name), v); }
int config_setting_lookup_int64(const config_setting_t *s, const char *name, long long *v)
{ return conf_get_int64(conf_member((conf_node_t *) s, name), v); }
int config_setting_lookup_float(const config_setting_t *s, const char *name, double *v)
{ return conf_get_float(conf_member((conf_node_t *) s, name), v); }
int config_setting_lookup_string(const config_setting_t *s, const char *name, const char **v)
{ return conf_get_string(conf_member((conf_node_t *) s, name), v); }
