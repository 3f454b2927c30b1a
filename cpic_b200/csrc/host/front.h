#ifndef CPIC_B200_FRONT_H
#define CPIC_B200_FRONT_H
#include "conf.h"
struct cpic_b200_conf { conf_node_t *root; };
void front_set_error(const char *fmt, ...);
#endif
