/*
 * csrc/ntt_launch.h -- internal glue between the translation units of the CUDA layer.
 *
 * ntt_kernels.cu holds the dispatch logic, the generic kernels and the helpers below; the ring kernels are
 * instantiated in their own translation units (ntt_ring_fp_{12,13,14}.cu, ntt_ring_int.cu) so that the library
 * builds in parallel.  Not part of the C-ABI.
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "ntt_cuda.h"

namespace nttb200 {

int nl_fail(const char *what, cudaError_t e);
int nl_fail_msg(const char *what);
int nl_sm_count(int device);
/* the coefficient array as rows of 128 bytes, one TMA box = `rows` rows, SWIZZLE_128B (cached per pointer/size) */
int nl_make_block_tmap(CUtensorMap *tm, uint64_t *d_a, size_t total_words, unsigned rows);
/* several small launches side by side (RNS limbs): chunks every CTA should at least get; 0 = use the whole GPU */
size_t nl_min_chunks_per_cta();

/* Parameters of several plans for ONE launch (RNS limbs: same N, one modulus and one set of tables each; limb l
 * owns polys_per_limb consecutive polynomials of the array).  Travels as a kernel argument (constant bank). */
constexpr int RING_MAX_LIMBS = 48;
template <bool MULTI>
struct RingLimbs {
  ntt_cuda_params_t e[MULTI ? RING_MAX_LIMBS : 1];
  uint32_t          polys_per_limb;
  uint32_t          ctas_per_limb; /* ring kernel: the grid is limbs x ctas_per_limb, a CTA serves one limb */
};
template <>
struct RingLimbs<false> {
  uint32_t polys_per_limb;
  uint32_t ctas_per_limb;
};

/* what the forward ring kernel may be asked to do on top of the transform */
struct RingOpts {
  const uint64_t *d_other    = nullptr; /* multiply pointwise by this transform-domain array before storing */
  size_t          other_mask = ~(size_t)0; /* chunk index mask into d_other: all ones = one operand per polynomial,
                                              2^(logn-L)-1 = ONE polynomial broadcast over the batch */
  bool            lazy_out   = false;   /* output may stay in [0,2q) (no final sign correction) */
};

/* one launcher per translation unit; `fwd` selects the direction */
int ring_fp_launch_10(bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st,
                      const RingOpts &o);
int ring_fp_launch_11(bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st,
                      const RingOpts &o);
int ring_fp_launch_12(bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st,
                      const RingOpts &o);
int ring_fp_launch_13(bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st,
                      const RingOpts &o);
int ring_fp_launch_14(bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st,
                      const RingOpts &o);
/* one-kernel negacyclic multiply, N = 2^13 (ntt_polymul_fp.cuh) */
int polymul_fp_launch(int device, const ntt_cuda_params_t &p, uint64_t *d_a, uint64_t *d_b, uint64_t *d_c,
                      size_t n_pairs, cudaStream_t st);
/* one launch of the L = 14 FP64 ring kernel over the chunks of several plans (same N, same range schedule) */
int ring_fp_launch_multi_14(bool fwd, int device, const ntt_cuda_params_t *const *plist, size_t n_limbs,
                            size_t polys_per_limb, uint64_t *d_a, cudaStream_t st);
/* strided passes in FP64 (ntt_strided_fp.cuh): global stages s0 .. s0+R-1 of `batch` polynomials; last: the inverse pass
 * that ends with global stage 0 (N^-1) */
int strided_fp_launch(bool fwd, bool last, int R, int device, const ntt_cuda_params_t &p, uint64_t *d_a, uint32_t s0,
                      size_t batch, cudaStream_t st);
int strided_fp_launch_multi(bool fwd, bool last, int R, int device, const RingLimbs<true> &lb, uint64_t *d_a, uint32_t s0,
                            size_t total_polys, cudaStream_t st);
int ring_int_launch(int L, bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks,
                    cudaStream_t st);

}  // namespace nttb200

#define NL_CU(call)                                             \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if(e_ != cudaSuccess) return nttb200::nl_fail(#call, e_);   \
  } while(0)
