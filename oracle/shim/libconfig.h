/* ORACLE TEST INFRASTRUCTURE — not product code.
 * Declaration-level stand-in for <libconfig.h> so the unmodified reference sources
 * compile in a container without libconfig. Implemented in shim_config.c on top of
 * the repo's own `.conf` reader (cpic_b200/csrc/host/conf.c). Only the entry points
 * the reference calls are provided (src/sim.c:46-82, src/config.c, src/specie.c:15-21,
 * src/particle.c:39,110-119,188, src/output.c:67-105, src/cpic.c:150-162). */
#ifndef ORACLE_SHIM_LIBCONFIG_H
#define ORACLE_SHIM_LIBCONFIG_H

#include <stdio.h>

#define CONFIG_TRUE 1
#define CONFIG_FALSE 0

#define CONFIG_TYPE_NONE 0
#define CONFIG_TYPE_GROUP 1
#define CONFIG_TYPE_INT 2
#define CONFIG_TYPE_INT64 3
#define CONFIG_TYPE_FLOAT 4
#define CONFIG_TYPE_STRING 5
#define CONFIG_TYPE_BOOL 6
#define CONFIG_TYPE_ARRAY 7
#define CONFIG_TYPE_LIST 8

struct conf_node;
typedef struct conf_node config_setting_t;

typedef struct config_t {
	config_setting_t *root;
	char *include_dir;
	char error_text[256];
	const char *error_file;
	int error_line;
} config_t;

void config_init(config_t *config);
void config_destroy(config_t *config);
void config_set_include_dir(config_t *config, const char *include_dir);
int config_read_file(config_t *config, const char *filename);
int config_read(config_t *config, FILE *stream);
int config_read_string(config_t *config, const char *str);
const char *config_error_text(const config_t *config);
const char *config_error_file(const config_t *config);
int config_error_line(const config_t *config);

config_setting_t *config_lookup(const config_t *config, const char *path);
int config_lookup_int(const config_t *config, const char *path, int *value);
int config_lookup_int64(const config_t *config, const char *path, long long *value);
int config_lookup_float(const config_t *config, const char *path, double *value);
int config_lookup_string(const config_t *config, const char *path, const char **value);
int config_lookup_bool(const config_t *config, const char *path, int *value);

int config_setting_type(const config_setting_t *setting);
int config_setting_length(const config_setting_t *setting);
const char *config_setting_name(const config_setting_t *setting);
unsigned int config_setting_source_line(const config_setting_t *setting);
config_setting_t *config_setting_get_elem(const config_setting_t *setting, unsigned int idx);
config_setting_t *config_setting_get_member(const config_setting_t *setting, const char *name);

int config_setting_get_int(const config_setting_t *setting);
long long config_setting_get_int64(const config_setting_t *setting);
double config_setting_get_float(const config_setting_t *setting);
const char *config_setting_get_string(const config_setting_t *setting);
int config_setting_get_int_elem(const config_setting_t *setting, int idx);
long long config_setting_get_int64_elem(const config_setting_t *setting, int idx);
double config_setting_get_float_elem(const config_setting_t *setting, int idx);

int config_setting_lookup_int(const config_setting_t *setting, const char *name, int *value);
int config_setting_lookup_int64(const config_setting_t *setting, const char *name, long long *value);
int config_setting_lookup_float(const config_setting_t *setting, const char *name, double *value);
int config_setting_lookup_string(const config_setting_t *setting, const char *name, const char **value);

#endif
