#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "br2cuda.h"
#include "br2_tables.cuh"
#include "br2_tile.cuh"
#ifndef K_NT
#define K_NT 64
#define K_MB 6
#define K_EX false
#define K_CL false
#define K_SPLIT false
#endif
void *keep() { return (void *)br2::br2_step_tile_kernel<2, K_NT, K_MB, K_EX, K_CL, false, K_SPLIT, false>; }
