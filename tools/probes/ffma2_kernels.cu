#include "ffma2_loop.cuh"
P_DEFINE_KERNELS(PSFX)
