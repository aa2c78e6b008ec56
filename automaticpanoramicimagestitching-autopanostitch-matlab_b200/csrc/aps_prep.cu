// aps_prep.cu -- K1: pooling / layout, L2 normalisation, operand conversion.
//
// Replaces  PP/featureMatching/featureMatchingGlobal.m:70-86 (vertcat + single + L2 normalise, eps
//           INSIDE the sqrt), PP/featureMatching/matchFeaturesScratch.m:105-110,217-234 (normalise iff
//           max|.|>2, eps OUTSIDE the sqrt) and the column-major -> row-major copies of
//           PP/mex/flann_knn.cpp:99-116,243-252.
//
// HBM-bound streaming kernels: every element is read once and written once; algorithmic bytes per
// descriptor row = D*4 (read) + D*4 (xn) + Dp*2 (bf16 operand) + 16 (sq, invn, scale/bias).
#include "aps_common.cuh"

// ------------------------------------------------------------------------------------------------
// column-major [N x D] -> row-major [N x D], 32x32 shared-memory tiles, coalesced both ways.
template <class T>
__global__ void k_transpose_in(const T* __restrict__ src, int64_t N, int D, T* __restrict__ dst) {
  __shared__ T tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {  // j: column inside the tile
    int64_t r = r0 + threadIdx.x;
    int c = c0 + j;
    if (r < N && c < D) tile[j][threadIdx.x] = src[r + (int64_t)c * N];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {  // j: row inside the tile
    int64_t r = r0 + j;
    int c = c0 + threadIdx.x;
    if (r < N && c < D) dst[r * D + c] = tile[threadIdx.x][j];
  }
}

int aps_k_transpose_in(cudaStream_t s, const void* src_cm, int64_t N, int D, int elem_size, void* dst_rm) {
  if (N == 0 || D == 0) return APS_OK;
  dim3 grid((unsigned)aps_ceil_div(N, 32), (unsigned)aps_ceil_div(D, 32)), block(32, 8);
  if (elem_size == 4)
    k_transpose_in<uint32_t><<<grid, block, 0, s>>>((const uint32_t*)src_cm, N, D, (uint32_t*)dst_rm);
  else
    k_transpose_in<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)src_cm, N, D, (uint8_t*)dst_rm);
  APS_LAUNCHED();
  return APS_OK;
}

// row-major [N x k] results -> column-major [N x k] (the MEX output layout, flann_knn.cpp:243-252)
__global__ void k_transpose_out(const uint32_t* __restrict__ idx_rm, const float* __restrict__ dist_rm, int64_t N,
                                int k, uint32_t* __restrict__ idx_cm, float* __restrict__ dist_cm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * k) return;
  int64_t r = i % N;
  int c = (int)(i / N);
  idx_cm[i] = idx_rm[r * k + c];
  dist_cm[i] = dist_rm[r * k + c];
}

int aps_k_transpose_out_u32f32(cudaStream_t s, const uint32_t* idx_rm, const float* dist_rm, int64_t N, int k,
                               uint32_t* idx_cm, float* dist_cm) {
  if (N == 0) return APS_OK;
  k_transpose_out<<<(unsigned)aps_ceil_div(N * k, 256), 256, 0, s>>>(idx_rm, dist_rm, N, k, idx_cm, dist_cm);
  APS_LAUNCHED();
  return APS_OK;
}

// ------------------------------------------------------------------------------------------------
// pass 1.  One block = RB rows staged in shared memory; the per-row sums are SEQUENTIAL float32
// (one thread per row, one rounding per operation, no FMA) so that the normalised values carry
// the same bits as the oracle's / the reference's single-precision arithmetic.
__global__ void k_prepare_norm(const float* __restrict__ raw, int64_t F, int D, int RB, int norm_mode,
                               float* __restrict__ xn, float* __restrict__ sq, float* __restrict__ invn,
                               int32_t* __restrict__ flags) {
  extern __shared__ float tile[];  // [RB][D+1]
  __shared__ float s_norm[64];
  __shared__ int s_exact;
  __shared__ int s_maxdev, s_maxsq, s_maxabs;
  const int ld = D + 1;
  const int64_t r0 = (int64_t)blockIdx.x * RB;
  const int nr = (int)min((int64_t)RB, F - r0);
  if (threadIdx.x == 0) {
    s_exact = 1;
    s_maxdev = 0;
    s_maxsq = 0;
    s_maxabs = 0;
  }
  __syncthreads();
  int exact = 1;
  float maxabs = 0.f;
  for (int i = threadIdx.x; i < nr * D; i += blockDim.x) {
    int r = i / D, c = i - r * D;
    float v = raw[(r0 + r) * D + c];
    tile[r * ld + c] = v;
    exact &= (__bfloat162float(__float2bfloat16_rn(v)) == v);
    maxabs = fmaxf(maxabs, fabsf(v));
  }
  if (!exact) atomicAnd(&s_exact, 0);
  atomicMax(&s_maxabs, __float_as_int(maxabs));
  __syncthreads();
  if (threadIdx.x < nr) {
    const float* x = tile + threadIdx.x * ld;
    float sum = 0.f;
    for (int c = 0; c < D; ++c) sum = __fadd_rn(sum, __fmul_rn(x[c], x[c]));
    float n = 1.0f;
    if (norm_mode == APS_NORM_GLOBAL) n = __fsqrt_rn(__fadd_rn(sum, APS_EPS32));      // featureMatchingGlobal.m:83-84
    if (norm_mode == APS_NORM_PAIRWISE) n = __fadd_rn(__fsqrt_rn(sum), APS_EPS32);    // matchFeaturesScratch.m:232
    s_norm[threadIdx.x] = n;
  }
  __syncthreads();
  if (norm_mode != APS_NORM_NONE) {
    for (int i = threadIdx.x; i < nr * D; i += blockDim.x) {
      int r = i / D, c = i - r * D;
      tile[r * ld + c] = __fdiv_rn(tile[r * ld + c], s_norm[r]);
    }
    __syncthreads();
  }
  if (threadIdx.x < nr) {
    const float* x = tile + threadIdx.x * ld;
    float sum = 0.f;
    for (int c = 0; c < D; ++c) sum = __fadd_rn(sum, __fmul_rn(x[c], x[c]));
    sq[r0 + threadIdx.x] = sum;
    invn[r0 + threadIdx.x] = __fdiv_rn(1.0f, s_norm[threadIdx.x]);
    atomicMax(&s_maxdev, __float_as_int(fabsf(sum - 1.0f)));
    atomicMax(&s_maxsq, __float_as_int(sum));
  }
  if (xn != raw || norm_mode != APS_NORM_NONE)
    for (int i = threadIdx.x; i < nr * D; i += blockDim.x) {
      int r = i / D, c = i - r * D;
      xn[(r0 + r) * D + c] = tile[r * ld + c];
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!s_exact) atomicAnd(&flags[0], 0);
    atomicMax(&flags[1], s_maxdev);
    atomicMax(&flags[2], s_maxsq);
    atomicMax(&flags[3], s_maxabs);
  }
}

int aps_k_prepare_norm(cudaStream_t s, const float* raw, int64_t F, int D, int norm_mode, float* xn, float* sq,
                       float* invn, int32_t* flags) {
  if (F == 0) return APS_OK;
  int RB = 11000 / (D + 1);
  if (RB > 64) RB = 64;
  if (RB < 1) {
    aps_set_error(APS_ERR_DIM, "", "descriptor dimension %d too large", D);
    return APS_ERR_DIM;
  }
  size_t smem = (size_t)RB * (D + 1) * sizeof(float);
  k_prepare_norm<<<(unsigned)aps_ceil_div(F, RB), 256, smem, s>>>(raw, F, D, RB, norm_mode, xn, sq, invn, flags);
  APS_LAUNCHED();
  return APS_OK;
}

// ------------------------------------------------------------------------------------------------
// pass 2: bf16 operand rows [F x Dp] (K-major, zero padded to Dp) + per-row (scale,bias).
__global__ void k_prepare_operands(const float* __restrict__ raw, const float* __restrict__ xn,
                                   const float* __restrict__ sq, const float* __restrict__ invn, int64_t F, int D,
                                   int Dp, const int32_t* __restrict__ exact_flag, int bias_mode,
                                   __nv_bfloat16* __restrict__ xb, float* __restrict__ colscale,
                                   float* __restrict__ colbias) {
  const int exact = *exact_flag;
  const float* src = exact ? raw : xn;
  const int chunks = Dp / 8;
  const int64_t total = F * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / chunks;
    int c0 = (int)(i - r * chunks) * 8;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = c0 + j;
      v[j] = __float2bfloat16_rn(c < D ? src[r * D + c] : 0.0f);
    }
    *reinterpret_cast<uint4*>(xb + r * Dp + c0) = *reinterpret_cast<const uint4*>(v);
    if (c0 == 0) {
      colscale[r] = exact ? invn[r] : 1.0f;
      colbias[r] = bias_mode ? -0.5f * sq[r] : 0.0f;
    }
  }
}

int aps_k_prepare_operands(cudaStream_t s, const float* raw, const float* xn, const float* sq, const float* invn,
                           int64_t F, int D, int Dp, const int32_t* exact_flag, int bias_mode, __nv_bfloat16* xb,
                           float* colscale, float* colbias) {
  if (F == 0) return APS_OK;
  int64_t total = F * (Dp / 8);
  unsigned grid = (unsigned)aps_min64(aps_ceil_div(total, 256), 148 * 16);
  k_prepare_operands<<<grid, 256, 0, s>>>(raw, xn, sq, invn, F, D, Dp, exact_flag, bias_mode, xb, colscale, colbias);
  APS_LAUNCHED();
  return APS_OK;
}
