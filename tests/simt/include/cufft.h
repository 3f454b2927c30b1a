/* simt-check stand-in for <cufft.h>: see ../simt.h (test infrastructure only) */
#include "../simt.h"
