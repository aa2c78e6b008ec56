// aps_prep.cu -- K1: pooling / layout, L2 normalisation, operand conversion.
//
// Replaces  PP/featureMatching/featureMatchingGlobal.m:70-86 (vertcat + single + L2 normalise, eps
//           INSIDE the sqrt), PP/featureMatching/matchFeaturesScratch.m:105-110,217-234 (normalise iff
//           max|.|>2, eps OUTSIDE the sqrt) and the column-major -> row-major copies of
//           PP/mex/flann_knn.cpp:99-116,243-252.
//
// HBM-bound streaming kernels: every element is read once and written once; algorithmic bytes per
// descriptor row = D*4 (read) + D*4 (xn) + Dp*2 (bf16 operand) + 16 (sq, invn, scale/bias).
#include <cuda_fp16.h>

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
// pass 1.  WARP-centric: each warp stages 32 rows in its own shared-memory tile with coalesced (float4) loads; the
// divisions x/norm are done by all lanes on coalesced elements, and lane r walks row r for the sums of squares in
// column order: the per-row sums stay SEQUENTIAL float32 (one rounding per operation, no FMA) so that the
// normalised values carry the same bits as the oracle's / the reference's single-precision arithmetic.  No block-wide
// barrier; row stride D+1 floats makes the per-lane walks bank-conflict free.  Measured (ncu, C2: 163840 x 128): 118 us
// for 84 MB read + 84 MB written = 1.4 TB/s, 0.22 of the HBM peak -- INSTRUCTION bound, not bandwidth bound: IPC 1.46,
// ~80 warp instructions per element-row, most of them the correctly rounded IEEE divisions (__fdiv_rn) and the
// exactness test; x * (1/n) would be 4x cheaper but does not give the oracle's bits (profiles/r2_ncu_aux_kernels.txt).  HBM streaming: reads 4D, writes 4D + 8 per row.
// img_off != nullptr: blockIdx.y = image, rows [img_off[y], img_off[y+1]) with its own flag words flags + 8*y (the
// per-image magnitude test of the pairwise path in ONE launch)
constexpr int PN_WARPS = 4;   // warps per block; shared memory = PN_WARPS * 32 * (D+1) * 4 bytes
// DT = compile-time descriptor length (64, 128: the index arithmetic of the coalesced phases becomes shifts; it was 40 %
// of the instructions with a run-time D) or 0 = generic.
template <int DT>
__global__ void __launch_bounds__(32 * PN_WARPS) k_prepare_norm(const float* __restrict__ raw, int64_t F, int Drt,
                                                                int norm_mode, float* __restrict__ xn,
                                                                float* __restrict__ sq, float* __restrict__ invn,
                                                                int32_t* __restrict__ flags,
                                                                const int64_t* __restrict__ img_off, int fp16) {
  extern __shared__ float tile_all[];
  const int D = DT ? DT : Drt;
  const int ld = D + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = tile_all + (size_t)warp * 32 * ld;
  int64_t row_lo = 0;
  if (img_off) {
    row_lo = img_off[blockIdx.y];
    F = img_off[blockIdx.y + 1];
    flags += 8 * blockIdx.y;
  }
  int exact = 1;
  float maxabs = 0.f, maxdev = 0.f, maxsq = 0.f;
  const bool write_xn = (xn != raw) || norm_mode != APS_NORM_NONE;
  const bool vec4 = (D % 4 == 0) && ((((uintptr_t)raw) | ((uintptr_t)xn)) % 16 == 0);
  for (int64_t r0 = row_lo + ((int64_t)blockIdx.x * PN_WARPS + warp) * 32; r0 < F; r0 += (int64_t)gridDim.x * PN_WARPS * 32) {
    const int nr = (int)min((int64_t)32, F - r0);
    const int n_el = nr * D;
    const float* src = raw + r0 * D;
    // phase A (all lanes, coalesced): load the rows, keep value and square
    if (vec4) {
      const int d4 = D / 4, n4 = n_el / 4;
      for (int i0 = 0; i0 < n4; i0 += 32 * 8) {   // 8 independent 16-byte loads per lane in flight
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 + lane;
          v[u] = i < n4 ? *reinterpret_cast<const float4*>(src + 4 * (size_t)i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 + lane;
          if (i < n4) {
            const int r = i / d4, c = (i - r * d4) * 4;
            const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              tile[r * ld + c + j] = e[j];
              exact &= fp16 ? (__half2float(__float2half_rn(e[j])) == e[j]) : (__bfloat162float(__float2bfloat16_rn(e[j])) == e[j]);
              maxabs = fmaxf(maxabs, fabsf(e[j]));
            }
          }
        }
      }
    } else {
      const int step_r = 32 / D, step_c = 32 - step_r * D;
      for (int i = lane, r = lane / D, c = lane - (lane / D) * D; i < n_el; i += 32) {
        const float v = src[i];
        tile[r * ld + c] = v;
        exact &= fp16 ? (__half2float(__float2half_rn(v)) == v) : (__bfloat162float(__float2bfloat16_rn(v)) == v);
        maxabs = fmaxf(maxabs, fabsf(v));
        r += step_r; c += step_c; if (c >= D) { c -= D; ++r; }
      }
    }
    __syncwarp();
    // phase B (lane = row): the SEQUENTIAL sum of the squares, then the row's norm
    float n = 1.0f, sum = 0.f;
    if (lane < nr) {
      const float* x = tile + lane * ld;
#pragma unroll 8
      for (int c = 0; c < D; ++c) sum = __fadd_rn(sum, __fmul_rn(x[c], x[c]));
      if (norm_mode == APS_NORM_GLOBAL) n = __fsqrt_rn(__fadd_rn(sum, APS_EPS32));      // featureMatchingGlobal.m:83-84
      if (norm_mode == APS_NORM_PAIRWISE) n = __fadd_rn(__fsqrt_rn(sum), APS_EPS32);    // matchFeaturesScratch.m:232
    }
    if (norm_mode != APS_NORM_NONE) {
      // phase C (all lanes): divide by the row's norm, store the normalised value and its square
      for (int r = 0; r < nr; ++r) {
        const float nrm = __shfl_sync(0xffffffffu, n, r);
        for (int c = lane; c < D; c += 32) {
          tile[r * ld + c] = __fdiv_rn(tile[r * ld + c], nrm);
        }
      }
      __syncwarp();
      // phase D (lane = row): sequential sum of the normalised squares
      if (lane < nr) {
        const float* x = tile + lane * ld;
        sum = 0.f;
#pragma unroll 8
        for (int c = 0; c < D; ++c) sum = __fadd_rn(sum, __fmul_rn(x[c], x[c]));
      }
    }
    if (lane < nr) {
      sq[r0 + lane] = sum;
      invn[r0 + lane] = __fdiv_rn(1.0f, n);
      maxdev = fmaxf(maxdev, fabsf(sum - 1.0f));
      maxsq = fmaxf(maxsq, sum);
    }
    if (write_xn) {
      float* dst = xn + r0 * D;
      if (vec4) {
        const int d4 = D / 4, n4 = n_el / 4;
        for (int i = lane; i < n4; i += 32) {
          const int r = i / d4, c = (i - r * d4) * 4;
          const float* t = tile + r * ld + c;
          *reinterpret_cast<float4*>(dst + 4 * (size_t)i) = make_float4(t[0], t[1], t[2], t[3]);
        }
      } else {
        const int step_r = 32 / D, step_c = 32 - step_r * D;
        for (int i = lane, r = lane / D, c = lane - (lane / D) * D; i < n_el; i += 32) {
          dst[i] = tile[r * ld + c];
          r += step_r; c += step_c; if (c >= D) { c -= D; ++r; }
        }
      }
    }
    __syncwarp();
  }
  // one set of atomics per warp
  exact = __all_sync(0xffffffffu, exact);
  for (int o = 16; o > 0; o >>= 1) {
    maxabs = fmaxf(maxabs, __shfl_xor_sync(0xffffffffu, maxabs, o));
    maxdev = fmaxf(maxdev, __shfl_xor_sync(0xffffffffu, maxdev, o));
    maxsq = fmaxf(maxsq, __shfl_xor_sync(0xffffffffu, maxsq, o));
  }
  if (lane == 0) {
    // the maxima are monotone: a plain read first lets almost every warp skip its atomic on the four shared words
    volatile int32_t* f = flags;
    if (!exact && f[0] != 0) atomicAnd(&flags[0], 0);
    if (__float_as_int(maxdev) > f[1]) atomicMax(&flags[1], __float_as_int(maxdev));
    if (__float_as_int(maxsq) > f[2]) atomicMax(&flags[2], __float_as_int(maxsq));
    if (__float_as_int(maxabs) > f[3]) atomicMax(&flags[3], __float_as_int(maxabs));
  }
}

static int prepare_norm_launch(cudaStream_t s, const float* raw, int64_t F, int64_t rows_per_y, unsigned ny, int D,
                               int norm_mode, float* xn, float* sq, float* invn, int32_t* flags, const int64_t* img_off,
                               int fp16) {
  const size_t smem = (size_t)PN_WARPS * 32 * (D + 1) * sizeof(float);
  if (smem > 200 * 1024) {
    aps_set_error(APS_ERR_DIM, "", "descriptor dimension %d too large", D);
    return APS_ERR_DIM;
  }
  auto kern = D == 128 ? k_prepare_norm<128> : (D == 64 ? k_prepare_norm<64> : (D == 48 ? k_prepare_norm<48> : k_prepare_norm<0>));
  APS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t gx = aps_ceil_div(rows_per_y, (int64_t)PN_WARPS * 32);
  const int64_t cap = aps_ceil_div((int64_t)148 * 8, (int64_t)ny);   // a few CTAs per SM; warps loop over the rest
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)(gx < 1 ? 1 : gx), ny);
  kern<<<grid, 32 * PN_WARPS, smem, s>>>(raw, F, D, norm_mode, xn, sq, invn, flags, img_off, fp16);
  APS_LAUNCHED();
  return APS_OK;
}

int aps_k_prepare_norm(cudaStream_t s, const float* raw, int64_t F, int D, int norm_mode, float* xn, float* sq,
                       float* invn, int32_t* flags, int fp16) {
  if (F == 0) return APS_OK;
  return prepare_norm_launch(s, raw, F, F, 1, D, norm_mode, xn, sq, invn, flags, nullptr, fp16);
}

int aps_k_prepare_norm_images(cudaStream_t s, const float* raw, const int64_t* d_img_off, int n, int64_t maxcount, int D,
                              int norm_mode, float* xn, float* sq, float* invn, int32_t* flags, int fp16) {
  if (n == 0 || maxcount == 0) return APS_OK;
  return prepare_norm_launch(s, raw, 0, maxcount, (unsigned)n, D, norm_mode, xn, sq, invn, flags, d_img_off, fp16);
}

// ------------------------------------------------------------------------------------------------
// pass 2: bf16 operand rows [F x Dp] (K-major, zero padded to Dp) + per-row (scale,bias).
__global__ void k_prepare_operands(const float* __restrict__ raw, const float* __restrict__ xn,
                                   const float* __restrict__ sq, const float* __restrict__ invn, int64_t F, int D,
                                   int Dp, const int32_t* __restrict__ exact_flag, int bias_mode,
                                   __nv_bfloat16* __restrict__ xb, float* __restrict__ colscale,
                                   float* __restrict__ colbias, int fp16) {
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
      const float f = c < D ? src[r * D + c] : 0.0f;
      if (fp16) reinterpret_cast<__half*>(v)[j] = __float2half_rn(f);   // same 16-bit storage, kind::f16 with F16 inputs
      else v[j] = __float2bfloat16_rn(f);
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
                           float* colscale, float* colbias, int fp16) {
  if (F == 0) return APS_OK;
  int64_t total = F * (Dp / 8);
  unsigned grid = (unsigned)aps_min64(aps_ceil_div(total, 256), 148 * 16);
  k_prepare_operands<<<grid, 256, 0, s>>>(raw, xn, sq, invn, F, D, Dp, exact_flag, bias_mode, xb, colscale, colbias, fp16);
  APS_LAUNCHED();
  return APS_OK;
}

// ------------------------------------------------------------------------------------------------
// Train-side view of the tensor kernel.  The epilogue pre-filters RAW accumulators against a per-row
// threshold (theta - bias_max(tile)) / scale_max(tile); to make that bound tight the train rows are
// bucket-sorted by their scale (1/norm), so the 128 rows of a tile have nearly the same scale.  perm[pos] is
// the original row of sorted position pos (order inside a bucket is arbitrary: the exact re-rank works
// on original indices, so results do not depend on it).
constexpr int NBUCKET = 4096;

__global__ void k_scale_minmax(const float* __restrict__ sc, int64_t N, int32_t* __restrict__ mm) {
  float mn = 3.0e38f, mx = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = sc[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {  // scales are positive: integer order == float order
    atomicMin(&mm[0], __float_as_int(mn));
    atomicMax(&mm[1], __float_as_int(mx));
  }
}
__device__ __forceinline__ int scale_bucket(float v, const int32_t* mm) {
  const float mn = __int_as_float(mm[0]), mx = __int_as_float(mm[1]);
  const float range = mx - mn;
  if (!(range > 0.f)) return 0;
  int b = (int)((v - mn) / range * (float)(NBUCKET - 1));
  return b < 0 ? 0 : (b >= NBUCKET ? NBUCKET - 1 : b);
}
__global__ void k_bucket_hist(const float* __restrict__ sc, int64_t N, const int32_t* __restrict__ mm,
                              int32_t* __restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&hist[scale_bucket(sc[i], mm)], 1);
}
__global__ void k_bucket_scan(const int32_t* __restrict__ hist, int32_t* __restrict__ cursor) {
  __shared__ int32_t part[1024];
  const int per = NBUCKET / 1024, t = threadIdx.x;
  int sum = 0;
  for (int i = 0; i < per; ++i) sum += hist[t * per + i];
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) {
      const int v = part[i];
      part[i] = run;
      run += v;
    }
  }
  __syncthreads();
  int run = part[t];
  for (int i = 0; i < per; ++i) {
    cursor[t * per + i] = run;
    run += hist[t * per + i];
  }
}
__global__ void k_bucket_scatter(const float* __restrict__ sc, int64_t N, const int32_t* __restrict__ mm,
                                 int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    perm[atomicAdd(&cursor[scale_bucket(sc[i], mm)], 1)] = (int32_t)i;
}
__global__ void k_gather_train(const uint4* __restrict__ xb, const float* __restrict__ sc, const float* __restrict__ bi,
                               const int32_t* __restrict__ perm, int64_t N, int chunks, uint4* __restrict__ xb_t,
                               float* __restrict__ sc_t, float* __restrict__ bi_t) {
  const int64_t total = N * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = i / chunks;
    const int c = (int)(i - pos * chunks);
    const int64_t r = perm[pos];
    xb_t[i] = xb[r * chunks + c];
    if (c == 0) {
      sc_t[pos] = sc[r];
      bi_t[pos] = bi[r];
    }
  }
}

int aps_k_sort_train_by_scale(cudaStream_t s, const __nv_bfloat16* xb, const float* colscale, const float* colbias,
                              int64_t N, int Dp, int32_t* scratch /* 2 + 2*NBUCKET ints */, int32_t* perm,
                              __nv_bfloat16* xb_t, float* colscale_t, float* colbias_t) {
  if (N == 0) return APS_OK;
  int32_t* mm = scratch;
  int32_t* hist = scratch + 2;
  int32_t* cursor = hist + NBUCKET;
  static const int32_t init_mm[2] = {0x7f7fffff, 0};
  APS_CUDA(cudaMemcpyAsync(mm, init_mm, sizeof init_mm, cudaMemcpyHostToDevice, s));
  APS_CUDA(cudaMemsetAsync(hist, 0, NBUCKET * sizeof(int32_t), s));
  const unsigned grid = (unsigned)aps_min64(aps_ceil_div(N, 256), 148 * 8);
  k_scale_minmax<<<grid, 256, 0, s>>>(colscale, N, mm);
  APS_LAUNCHED();
  k_bucket_hist<<<grid, 256, 0, s>>>(colscale, N, mm, hist);
  APS_LAUNCHED();
  k_bucket_scan<<<1, 1024, 0, s>>>(hist, cursor);
  APS_LAUNCHED();
  k_bucket_scatter<<<grid, 256, 0, s>>>(colscale, N, mm, cursor, perm);
  APS_LAUNCHED();
  const int chunks = Dp * 2 / 16;
  const unsigned g2 = (unsigned)aps_min64(aps_ceil_div(N * chunks, 256), 148 * 16);
  k_gather_train<<<g2, 256, 0, s>>>((const uint4*)xb, colscale, colbias, perm, N, chunks, (uint4*)xb_t, colscale_t,
                                    colbias_t);
  APS_LAUNCHED();
  return APS_OK;
}
int aps_sort_scratch_ints() { return 2 + 2 * NBUCKET; }

// per 128-row tile of the train view: (1/scale_max, 1/scale_min) with a 1e-6 safety factor, bias_max
__global__ void k_tile_bounds(const float* __restrict__ sc, const float* __restrict__ bi, int64_t N, int tile_rows,
                              int64_t ntiles, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= ntiles) return;
  float smax = 0.f, smin = 3.0e38f, bmax = -3.0e38f;
  for (int r = lane; r < tile_rows; r += 32) {
    const int64_t j = t * tile_rows + r;
    if (j < N) {
      smax = fmaxf(smax, sc[j]);
      smin = fminf(smin, sc[j]);
      bmax = fmaxf(bmax, bi[j]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
  }
  if (lane == 0) {
    float4 v;
    v.x = smax > 0.f ? (1.0f / smax) * (1.0f - 1.0e-6f) : 0.f;      // multiplies (theta - bias_max) >= 0
    v.y = (smin > 0.f && smin < 3.0e38f) ? (1.0f / smin) * (1.0f + 1.0e-6f) : 3.0e38f;  // ... when it is negative
    v.z = bmax > -3.0e38f ? bmax : 0.f;
    v.w = 0.f;
    out[t] = v;
  }
}
int aps_k_tile_bounds(cudaStream_t s, const float* colscale, const float* colbias, int64_t N, int tile_rows,
                      float4* out) {
  const int64_t ntiles = aps_ceil_div(N, tile_rows);
  if (ntiles == 0) return APS_OK;
  k_tile_bounds<<<(unsigned)aps_ceil_div(ntiles, 8), 256, 0, s>>>(colscale, colbias, N, tile_rows, ntiles, out);
  APS_LAUNCHED();
  return APS_OK;
}
