// aps_pairwise.cu -- batched pairwise matching: every (query image i, train image j) pair of a batch in
// a handful of launches instead of a launch sequence per pair.
//
// Replaces  the parfor over image pairs of PP/featureMatching/featureMatchingPairwise.m:48-62 with
//           getMatches -> matchFeaturesScratch (Exhaustive, Unique=true) per pair
//           (PP/featureMatching/matchFeaturesScratch.m:116-126,169-211, PP/mex/nearest2HammingExhaustive*MEX.cpp).
//
// Bookkeeping (aps_pair_tables): the batch's pairs are laid out back to back; "entry" e in
// [eoff[p], eoff[p+1]) is query row qoff[p] + e - eoff[p] of pair p, searched in train rows
// [toff[p], toff[p]+tcnt[p]).  All per-entry arrays (2-NN tables, keys, winners, matches) use that index.
//   stage 1  2-NN per entry: float -> tcgen05 candidates (unit table) + exact re-rank (+ k_pair_exact2 for
//            unproven entries, or for all entries on the exact engine); binary -> k_knn_hamming_tab
//   stage 2  ratio / threshold keys, atomicMin per train row (uniqueness), winners, rank-by-counting
#include <math_constants.h>

#include "aps_common.cuh"
#include "aps_exact_math.cuh"
#include "aps_hamming.cuh"

namespace {

__device__ __forceinline__ int pair_of_entry(const aps_pair_tables& pt, int64_t e) {
  int lo = 0, hi = pt.npairs;  // last p with eoff[p] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pt.eoff[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- exact 2-NN of single entries: one warp per entry (fallback list, or every entry) ----------------------
template <int METRIC>
__global__ void __launch_bounds__(256) k_pair_exact2(const float* __restrict__ X, const float* __restrict__ sq,
                                                     const float* __restrict__ XT, const float* __restrict__ sqT, int D,
                                                     aps_pair_tables pt, const int32_t* __restrict__ rows,
                                                     const int32_t* __restrict__ nrows_dev, int64_t n_entries,
                                                     uint32_t* __restrict__ idx, float* __restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const int64_t total = rows ? (int64_t)(*nrows_dev) : n_entries;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total;
       w += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t e = rows ? (int64_t)rows[w] : w;
    const int p = pair_of_entry(pt, e);
    const int64_t q = (int64_t)pt.qoff[p] + (e - pt.eoff[p]);
    const int64_t t0 = pt.toff[p];
    const int tn = pt.tcnt[p];
    const float* a = X + q * D;
    const float a2 = METRIC == 1 ? sq[q] : 0.f;
    float d0 = CUDART_INF_F, d1 = CUDART_INF_F;  // this lane's two best (ascending j => ties keep the lower index)
    int j0 = -1, j1 = -1;
    for (int j = lane; j < tn; j += 32) {
      const float* b = XT + (t0 + j) * D;
      const float d = METRIC == 0   ? l2sq_flann(a, b, D)
                      : METRIC == 2 ? __fsqrt_rn(l2sq_seq(a, b, D))
                      : METRIC == 3 ? -dot_seq(a, b, D)
                                    : ssd_seq(a, b, D, a2, sqT[t0 + j]);
      if (d < d0) { d1 = d0; j1 = j0; d0 = d; j0 = j; }
      else if (d < d1) { d1 = d; j1 = j; }
    }
    for (int c = 0; c < 2; ++c) {  // warp merge by (distance, index)
      float bd = d0;
      int bj = j0 < 0 ? 0x7fffffff : j0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
      }
      const bool found = bj != 0x7fffffff;
      if (lane == 0) {
        idx[e * 2 + c] = found ? (uint32_t)(bj + 1) : 0u;
        dist[e * 2 + c] = found ? (METRIC == 2   ? __fmul_rn(bd, bd)
                                   : METRIC == 3 ? __fsub_rn(2.0f, __fmul_rn(2.0f, -bd))
                                                 : bd)
                                : CUDART_INF_F;
      }
      if (found && j0 == bj) { d0 = d1; j0 = j1; d1 = CUDART_INF_F; j1 = -1; }  // pop the winner
    }
  }
}

// ---- batched Hamming 2-NN: one CTA = up to 256 queries of ONE pair (block table) -------------------------
struct HamBlock {
  int32_t q0, nq, t0, t1;
  int64_t out_row;
};
constexpr int HQ = 256, HT = 512;

template <int NW>
__global__ void __launch_bounds__(HQ) k_knn_hamming_tab(const uint4* __restrict__ X, const HamBlock* __restrict__ tab,
                                                        uint32_t* __restrict__ idx, float* __restrict__ dist) {
  __shared__ uint4 ts[HT * NW];
  const HamBlock hb = tab[blockIdx.x];
  const int tid = threadIdx.x;
  const bool qvalid = tid < hb.nq;
  uint4 a[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) a[w] = qvalid ? X[(int64_t)(hb.q0 + tid) * NW + w] : make_uint4(0, 0, 0, 0);
  int b0 = 0x7fffffff, b1 = 0x7fffffff;
  uint32_t i0 = 0, i1 = 0;
  for (int j0 = hb.t0; j0 < hb.t1; j0 += HT) {
    const int nj = min(HT, hb.t1 - j0);
    __syncthreads();
    for (int f = tid; f < nj * NW; f += HQ) ts[f] = X[(int64_t)j0 * NW + f];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nj; ++j) {
      const int h = hamming_words<NW>(a, ts + j * NW);
      if (h < b1) {  // ascending scan + strict '<' : first index attaining the minimum, second with multiplicity
        const uint32_t id = (uint32_t)(j0 + j - hb.t0 + 1);
        if (h < b0) { b1 = b0; i1 = i0; b0 = h; i0 = id; }
        else { b1 = h; i1 = id; }
      }
    }
  }
  if (qvalid) {
    const int64_t o = (hb.out_row + tid) * 2;
    idx[o] = i0;
    idx[o + 1] = i1;
    dist[o] = i0 ? (float)b0 : CUDART_INF_F;
    dist[o + 1] = i1 ? (float)b1 : CUDART_INF_F;
  }
}

// ---- [E x 2] tables -> idx2 / d1 / d2 with the MEX edge rules -----------------------------------------------
__global__ void k_pairs_k2_to_nn(aps_pair_tables pt, int64_t E, int is_binary, int nb, const uint32_t* __restrict__ i2,
                                 const float* __restrict__ dd, uint32_t* __restrict__ idx2, float* __restrict__ d1,
                                 float* __restrict__ d2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  idx2[e] = i2[2 * e];
  d1[e] = dd[2 * e];
  float s = dd[2 * e + 1];
  if (is_binary) {  // nearest2HammingExhaustiveMEX.cpp:71-74 : N2 == 1 -> second = 8*nb
    const int p = pair_of_entry(pt, e);
    if (pt.tcnt[p] == 1 || i2[2 * e + 1] == 0u) s = (float)(nb * 8);
  }
  d2[e] = s;
}

// ---- stage 2 (matchFeaturesScratch.m:169-211), batched --------------------------------------------------------
__device__ __forceinline__ uint32_t f32_order_bits(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_order_bits(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}

__global__ void k_pairs_keys(aps_pair_tables pt, int64_t E, const uint32_t* __restrict__ idx2,
                             const float* __restrict__ d1, const float* __restrict__ d2, int is_binary, int nbits,
                             double match_threshold, double max_ratio, unsigned long long* __restrict__ best,
                             unsigned long long* __restrict__ keys) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  bool keep;
  float dB;
  if (is_binary) {
    float s = d2[e];
    if (!isfinite(s) || s == 0.0f) s = (float)nbits;                                // :318
    dB = __fmul_rn(__fdiv_rn(d1[e], (float)nbits), 100.0f);                         // :120
    const float dS = __fmul_rn(__fdiv_rn(s, (float)nbits), 100.0f);                 // :121
    const float rhs = __fmul_rn((float)max_ratio, dS);                              // :171 (single arithmetic)
    keep = (dB <= rhs) && (dB <= (float)match_threshold) && isfinite(dB) && isfinite(dS);
  } else {
    dB = d1[e];
    const double b = (double)d1[e], sd = (double)d2[e];
    const double r2 = max_ratio * max_ratio;                                        // :173
    keep = (b <= r2 * sd) && (b <= match_threshold) && isfinite(b) && isfinite(sd); // :174-178
  }
  unsigned long long key = ~0ull;
  if (keep) {
    const int p = pair_of_entry(pt, e);
    const uint32_t qloc = (uint32_t)(e - pt.eoff[p]);
    key = ((unsigned long long)f32_order_bits(dB) << 32) | (unsigned long long)qloc;  // order: (d, query)
    atomicMin(&best[pt.boff[p] + (idx2[e] - 1)], key);                                 // Unique: per train row
  }
  keys[e] = key;
}

__global__ void k_pairs_collect(aps_pair_tables pt, int64_t E, const uint32_t* __restrict__ idx2,
                                const unsigned long long* __restrict__ best,
                                const unsigned long long* __restrict__ keys, int32_t* __restrict__ count,
                                unsigned long long* __restrict__ winners) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const unsigned long long key = keys[e];
  if (key == ~0ull) return;
  const int p = pair_of_entry(pt, e);
  if (best[pt.boff[p] + (idx2[e] - 1)] != key) return;
  const int pos = atomicAdd(&count[p], 1);
  winners[pt.eoff[p] + pos] = key;
}

// one CTA per pair: rank the pair's winners by counting (unique keys), emit in ascending (d, query) order
__global__ void __launch_bounds__(256) k_pairs_rank_emit(aps_pair_tables pt, const unsigned long long* __restrict__ winners,
                                                         const int32_t* __restrict__ count,
                                                         const uint32_t* __restrict__ idx2,
                                                         uint32_t* __restrict__ matches, double* __restrict__ metric) {
  __shared__ unsigned long long tile[1024];
  const int p = blockIdx.x;
  const int W = count[p];
  const int64_t base = pt.eoff[p];
  for (int w0 = 0; w0 < W; w0 += blockDim.x) {
    const int w = w0 + threadIdx.x;
    const unsigned long long mine = (w < W) ? winners[base + w] : ~0ull;
    int r = 0;
    for (int t0 = 0; t0 < W; t0 += 1024) {
      __syncthreads();
      for (int t = threadIdx.x; t < 1024; t += blockDim.x) tile[t] = (t0 + t < W) ? winners[base + t0 + t] : ~0ull;
      __syncthreads();
      const int nt = min(1024, W - t0);
      for (int t = 0; t < nt; ++t) r += (tile[t] < mine);
    }
    if (w < W) {
      const uint32_t q = (uint32_t)(mine & 0xffffffffull);
      uint32_t t = idx2[base + q];
      if (pt.vmap && (int64_t)pt.toff[p] >= pt.vfirst)   // subset image: position in candB -> row of B (:406, idx2 = candB(I))
        t = (uint32_t)pt.vmap[(int64_t)pt.toff[p] - pt.vfirst + (int64_t)t - 1] + 1u;
      matches[2 * (base + r)] = q + 1;
      matches[2 * (base + r) + 1] = t;
      metric[base + r] = (double)f32_from_order_bits((uint32_t)(mine >> 32));
    }
  }
}

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------------------------
int aps_k_pair_exact2(cudaStream_t s, const float* X, const float* sq, const float* XT, const float* sqT, int D, int metric,
                      const aps_pair_tables& pt, const int32_t* rows, const int32_t* nrows_dev, int64_t n_entries,
                      uint32_t* idx, float* dist) {
  if (n_entries == 0) return APS_OK;
  const unsigned grid = (unsigned)aps_min64(aps_ceil_div(n_entries, 8), 148 * 8);
  if (metric == 0)
    k_pair_exact2<0><<<grid, 256, 0, s>>>(X, sq, XT, sqT, D, pt, rows, nrows_dev, n_entries, idx, dist);
  else if (metric == 2)
    k_pair_exact2<2><<<grid, 256, 0, s>>>(X, sq, XT, sqT, D, pt, rows, nrows_dev, n_entries, idx, dist);
  else if (metric == 3)
    k_pair_exact2<3><<<grid, 256, 0, s>>>(X, sq, XT, sqT, D, pt, rows, nrows_dev, n_entries, idx, dist);
  else
    k_pair_exact2<1><<<grid, 256, 0, s>>>(X, sq, XT, sqT, D, pt, rows, nrows_dev, n_entries, idx, dist);
  APS_LAUNCHED();
  return APS_OK;
}

int aps_k_pairs_hamming2(cudaStream_t s, const uint8_t* Xpad, int nb16, const std::vector<int64_t>& eoff,
                         const std::vector<int32_t>& qoff, const std::vector<int32_t>& toff,
                         const std::vector<int32_t>& tcnt, uint32_t* idx, float* dist) {
  std::vector<HamBlock> tab;
  const size_t np = qoff.size();
  for (size_t p = 0; p < np; ++p) {
    const int64_t nq = eoff[p + 1] - eoff[p];
    if (tcnt[p] == 0) continue;
    for (int64_t c = 0; c < nq; c += HQ) {
      HamBlock hb;
      hb.q0 = (int32_t)(qoff[p] + c);
      hb.nq = (int32_t)aps_min64(HQ, nq - c);
      hb.t0 = toff[p];
      hb.t1 = toff[p] + tcnt[p];
      hb.out_row = eoff[p] + c;
      tab.push_back(hb);
    }
  }
  if (tab.empty()) return APS_OK;
  DevBuf<HamBlock> dtab;
  APS_TRY(dtab.alloc(tab.size(), s));
  APS_CUDA(cudaMemcpyAsync(dtab.p, tab.data(), tab.size() * sizeof(HamBlock), cudaMemcpyHostToDevice, s));
  APS_CUDA(cudaStreamSynchronize(s));  // `tab` is pageable host memory
  const unsigned grid = (unsigned)tab.size();
  const uint4* x4 = (const uint4*)Xpad;
  switch (nb16 / 16) {
    case 1: k_knn_hamming_tab<1><<<grid, HQ, 0, s>>>(x4, dtab.p, idx, dist); break;
    case 2: k_knn_hamming_tab<2><<<grid, HQ, 0, s>>>(x4, dtab.p, idx, dist); break;
    case 3: k_knn_hamming_tab<3><<<grid, HQ, 0, s>>>(x4, dtab.p, idx, dist); break;
    case 4: k_knn_hamming_tab<4><<<grid, HQ, 0, s>>>(x4, dtab.p, idx, dist); break;
    default:
      aps_set_error(APS_ERR_DIM, "hamm2nn:cols", "binary descriptors wider than 64 bytes are not supported (%d)", nb16);
      return APS_ERR_DIM;
  }
  APS_LAUNCHED();
  return APS_OK;
}

int aps_k_pairs_k2_to_nn(cudaStream_t s, const aps_pair_tables& pt, int64_t E, int is_binary, int nb, const uint32_t* i2,
                         const float* dd, uint32_t* idx2, float* d1, float* d2) {
  if (E == 0) return APS_OK;
  k_pairs_k2_to_nn<<<(unsigned)aps_ceil_div(E, 256), 256, 0, s>>>(pt, E, is_binary, nb, i2, dd, idx2, d1, d2);
  APS_LAUNCHED();
  return APS_OK;
}

// stage 2 for a whole batch.  best: [sum tcnt] (pre-set to 0xff), keys / winners: [E], count: [npairs] (zeroed here)
int aps_k_pairs_filter_unique(cudaStream_t s, const aps_pair_tables& pt, int64_t E, int64_t Btotal,
                              const uint32_t* idx2, const float* d1, const float* d2, int is_binary, int nbits,
                              double match_threshold, double max_ratio, unsigned long long* best,
                              unsigned long long* keys, unsigned long long* winners, int32_t* count, uint32_t* matches,
                              double* metric) {
  APS_CUDA(cudaMemsetAsync(count, 0, (size_t)pt.npairs * sizeof(int32_t), s));
  if (E == 0 || pt.npairs == 0) return APS_OK;
  if (Btotal > 0) APS_CUDA(cudaMemsetAsync(best, 0xff, (size_t)Btotal * sizeof(unsigned long long), s));
  const unsigned grid = (unsigned)aps_ceil_div(E, 256);
  k_pairs_keys<<<grid, 256, 0, s>>>(pt, E, idx2, d1, d2, is_binary, nbits, match_threshold, max_ratio, best, keys);
  APS_LAUNCHED();
  k_pairs_collect<<<grid, 256, 0, s>>>(pt, E, idx2, best, keys, count, winners);
  APS_LAUNCHED();
  k_pairs_rank_emit<<<(unsigned)pt.npairs, 256, 0, s>>>(pt, winners, count, idx2, matches, metric);
  APS_LAUNCHED();
  return APS_OK;
}
