/* simt-check stand-in for <cuda.h>: see ../simt.h (test infrastructure only) */
#include "../simt.h"
