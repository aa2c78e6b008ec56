// aps_hamming.cuh -- XOR + population count of packed binary descriptors (shared by the global and the
// batched pairwise Hamming kernels).
#pragma once
#include <cstdint>

// Hamming distance of two NW x 128-bit descriptors.  POPC issues at 16 lanes/clk/SM (the binding pipe of
// the plain 8-POPC form, profiles/r1_ncu_k_knn_hamming.txt), LOP3 at 64: carry-save adders trade three
// POPCs for six LOP3s per 256 bits -- popc(x0..x7) = popc(s2) + popc(x7) + 2*(popc(c0)+popc(c1)+popc(c2)).
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
  sum = a ^ b ^ c;                        // LOP3 0x96
  carry = (a & b) | (a & c) | (b & c);    // LOP3 0xE8
}
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
  const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
  uint32_t s0, c0, s1, c1, s2, c2;
  csa(x0, x1, x2, s0, c0);
  csa(x3, x4, x5, s1, c1);
  csa(s0, s1, x6, s2, c2);
  return __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
}
template <int NW>
__device__ __forceinline__ int hamming_words(const uint4 (&a)[NW], const uint4* __restrict__ b) {
  int h = 0;
  if (NW % 2 == 0) {
#pragma unroll
    for (int w = 0; w < NW; w += 2) h += hamming256(a[w], a[w + 1], b[w], b[w + 1]);
  } else {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint4 t = b[w];
      h += __popc(a[w].x ^ t.x) + __popc(a[w].y ^ t.y) + __popc(a[w].z ^ t.z) + __popc(a[w].w ^ t.w);
    }
  }
  return h;
}

