/* simt-check stand-in for <cuda_runtime.h>: see ../simt.h (test infrastructure only) */
#include "../simt.h"
