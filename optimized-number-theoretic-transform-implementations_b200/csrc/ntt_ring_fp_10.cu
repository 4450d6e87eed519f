/* FP64 ring kernels for chunks of 2^10 coefficients (see ntt_ring_fp.cuh); one translation unit per chunk size. */
#define NTT_RING_L 10
#include "ntt_ring_fp_launch.inl"
