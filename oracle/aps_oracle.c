/*
 * aps_oracle.c -- CPU restatement of AutoPanoStitch's featureMatching/ hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package, csrc/, mex/) may
 * import, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * PP/ = "/root/reference/Procedural Program/".  Each function cites the reference lines it
 * restates.  Parity pinning (see DESIGN.md "Oracle pinning"):
 *   - orc_nearest2_hamming  : pinned bit-for-bit against the reference's own MEX sources
 *                             compiled verbatim into oracle/_ref (tests/test_oracle_pinning.py)
 *                             and against committed golden vectors made from them.
 *   - orc_knn_hamming       : pinned against cv2.BFMatcher(NORM_HAMMING).knnMatch (OpenCV 4.13;
 *                             the reference pins 4.12) golden vectors -- PP/mex/flann_knn.cpp:199-223.
 *   - orc_knn_l2            : exact search with the flann_knn output contract; pinned against
 *                             cv2.flann_Index(LINEAR).knnSearch golden vectors = FLANN's own L2<float>
 *                             functor and result order: indices AND float32 distance bits identical.
 *                             The reference's default float engine (FLANN KD-tree, randomised,
 *                             approximate) cannot pin an exact matcher.
 *   - MATLAB-only steps (normalisations, filter loop, ratio tests, unique, SSD 2-NN, top-m, packBits): no
 *     MATLAB / Octave here and the reference ships no tests => "parity unpinned" against MATLAB itself.
 *     Cross-checked bit for bit against an independent numpy restatement of the same .m lines
 *     (tests/matlab_restatement.py, tests/test_oracle_matlab_restatement.py: tie-heavy random inputs) and
 *     against golden vectors minted from that restatement alone (tests/golden/matlab_semantics_v1.npz),
 *     plus hand-made known-answer tests (tests/test_oracle_semantics.py).
 *
 * Floating-point discipline: compile with -ffp-contract=off (no FMA), every float operation is a
 * single IEEE-754 binary32 operation in the order written, so that the CUDA re-rank kernel
 * (__fadd_rn/__fmul_rn, same order) reproduces the bits.
 *
 * All matrices are ROW-major [N x D] here (the MEX boundary's column-major layout is an ABI
 * detail tested separately).  All returned indices are 1-based like the reference's.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EPS32 1.1920928955078125e-07f /* eps('single') = 2^-23 */

int orc_version(void) { return 1; }
/* bench.py's CPU arms use every host core even when the launcher (torch.distributed.run) exported OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * A1  featureMatchingGlobal.m:80-84     allDesc ./ sqrt(sum(allDesc.^2,2) + eps('single'))
 *     (eps INSIDE the sqrt).  Sum of squares: sequential over the D columns, float32.
 * ---------------------------------------------------------------------------------------- */
void orc_normalize_rows_global(float* X, int64_t N, int D) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < N; ++r) {
    float* x = X + r * (int64_t)D;
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
      float sq = x[d] * x[d];
      s = s + sq;
    }
    float n = sqrtf(s + ORC_EPS32);
    for (int d = 0; d < D; ++d) x[d] = x[d] / n;
  }
}

/* ------------------------------------------------------------------------------------------
 * A5  matchFeaturesScratch.m:217-234 (normalizeRowsL2): X ./ (sqrt(sum(X.^2,2)) + eps('single'))
 *     (eps OUTSIDE the sqrt -- differs from A1).
 * ---------------------------------------------------------------------------------------- */
void orc_normalize_rows_pairwise(float* X, int64_t N, int D) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < N; ++r) {
    float* x = X + r * (int64_t)D;
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
      float sq = x[d] * x[d];
      s = s + sq;
    }
    float n = sqrtf(s) + ORC_EPS32;
    for (int d = 0; d < D; ++d) x[d] = x[d] / n;
  }
}

/* matchFeaturesScratch.m:105  max(abs(A(:))) > 2 */
int orc_needs_normalization(const float* A, int64_t nA, const float* B, int64_t nB) {
  for (int64_t i = 0; i < nA; ++i)
    if (fabsf(A[i]) > 2.0f) return 1;
  for (int64_t i = 0; i < nB; ++i)
    if (fabsf(B[i]) > 2.0f) return 1;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Squared-L2 functor in the order of OpenCV 4.12 cvflann::L2<float>::operator() (modules/flann/
 * include/opencv2/flann/dist.h -- third-party, un-vendored; pinned by PP/mex/flann_knn.cpp:1,231
 * FLANN_DIST_L2).  Published algorithm: groups of four, result += d0*d0+d1*d1+d2*d2+d3*d3,
 * then a scalar tail.  This is the distance flann_knn_win returns for float descriptors
 * (squared, PP/mex/flann_knn.cpp:229-234, 249).
 * ---------------------------------------------------------------------------------------- */
static inline float l2sq_flann_order(const float* a, const float* b, int D) {
  float result = 0.0f;
  int d = 0;
  for (; d + 3 < D; d += 4) {
    float d0 = a[d] - b[d], d1 = a[d + 1] - b[d + 1], d2 = a[d + 2] - b[d + 2], d3 = a[d + 3] - b[d + 3];
    float g = d0 * d0 + d1 * d1;
    g = g + d2 * d2;
    g = g + d3 * d3;
    result = result + g;
  }
  for (; d < D; ++d) {
    float d0 = a[d] - b[d];
    result = result + d0 * d0;
  }
  return result;
}
float orc_l2sq(const float* a, const float* b, int D) { return l2sq_flann_order(a, b, D); }

/* insert (dist,idx) into ascending top-k; scan order is ascending idx, so strict '<' gives
 * "ties -> lower train index" (what cv::BFMatcher::knnMatch returns). */
static inline void topk_insert_f(float* bd, uint32_t* bi, int k, float dist, uint32_t idx1) {
  if (!(dist < bd[k - 1])) return;
  int p = k - 1;
  while (p > 0 && dist < bd[p - 1]) {
    bd[p] = bd[p - 1];
    bi[p] = bi[p - 1];
    --p;
  }
  bd[p] = dist;
  bi[p] = idx1;
}

/* ------------------------------------------------------------------------------------------
 * A2 (float)  flann_knn_win(train, query, k, 'flann', ...) output contract with an EXACT search:
 *   idx [Fq x k] uint32 1-based, dist [Fq x k] float squared L2, ascending, missing -> 0 / +inf
 *   (PP/mex/flann_knn.cpp:193-194, 216-219, 243-252).  Outputs here are ROW-major [Fq][k].
 *   Implementation note: train is transposed once so the j loop vectorises; per (q,j) the
 *   operation order is exactly l2sq_flann_order.
 * ---------------------------------------------------------------------------------------- */
void orc_knn_l2(const float* train, int64_t Ft, const float* query, int64_t Fq, int D, int k, uint32_t* idx,
                float* dist) {
  const int JB = 512;
  float* T = (float*)malloc((size_t)(Ft > 0 ? Ft : 1) * D * sizeof(float)); /* [D][Ft] */
  for (int64_t j = 0; j < Ft; ++j)
    for (int d = 0; d < D; ++d) T[(int64_t)d * Ft + j] = train[j * D + d];
#pragma omp parallel
  {
    float* acc = (float*)malloc(JB * sizeof(float));
#pragma omp for schedule(dynamic, 16)
    for (int64_t q = 0; q < Fq; ++q) {
      const float* a = query + q * D;
      float* bd = dist + q * k;
      uint32_t* bi = idx + q * k;
      for (int c = 0; c < k; ++c) {
        bd[c] = INFINITY;
        bi[c] = 0;
      }
      for (int64_t j0 = 0; j0 < Ft; j0 += JB) {
        int nj = (int)((Ft - j0 < JB) ? (Ft - j0) : JB);
        for (int j = 0; j < nj; ++j) acc[j] = 0.0f;
        int d = 0;
        for (; d + 3 < D; d += 4) {
          const float a0 = a[d], a1 = a[d + 1], a2 = a[d + 2], a3 = a[d + 3];
          const float *t0 = T + (int64_t)d * Ft + j0, *t1 = t0 + Ft, *t2 = t1 + Ft, *t3 = t2 + Ft;
#pragma omp simd
          for (int j = 0; j < nj; ++j) {
            float d0 = a0 - t0[j], d1 = a1 - t1[j], d2 = a2 - t2[j], d3 = a3 - t3[j];
            float g = d0 * d0 + d1 * d1;
            g = g + d2 * d2;
            g = g + d3 * d3;
            acc[j] = acc[j] + g;
          }
        }
        for (; d < D; ++d) {
          const float a0 = a[d];
          const float* t0 = T + (int64_t)d * Ft + j0;
#pragma omp simd
          for (int j = 0; j < nj; ++j) {
            float d0 = a0 - t0[j];
            acc[j] = acc[j] + d0 * d0;
          }
        }
        const float worst = bd[k - 1];
        (void)worst;
        for (int j = 0; j < nj; ++j)
          if (acc[j] < bd[k - 1]) topk_insert_f(bd, bi, k, acc[j], (uint32_t)(j0 + j + 1));
      }
    }
    free(acc);
  }
  free(T);
}

static inline int hamming_bytes(const uint8_t* a, const uint8_t* b, int nb) {
  int h = 0, i = 0;
  for (; i + 8 <= nb; i += 8) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);
    memcpy(&y, b + i, 8);
    h += __builtin_popcountll(x ^ y);
  }
  for (; i < nb; ++i) h += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return h;
}

/* ------------------------------------------------------------------------------------------
 * A2 (binary, 'bf')  cv::BFMatcher(NORM_HAMMING,false).knnMatch(query, train, k)
 *   PP/mex/flann_knn.cpp:199-223 : exact, ascending Hamming bit counts cast to float,
 *   ties -> lower train index, missing -> idx 0 / +inf.  ROW-major outputs [Fq][k].
 * ---------------------------------------------------------------------------------------- */
void orc_knn_hamming(const uint8_t* train, int64_t Ft, const uint8_t* query, int64_t Fq, int nb, int k,
                     uint32_t* idx, float* dist) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t q = 0; q < Fq; ++q) {
    float* bd = dist + q * k;
    uint32_t* bi = idx + q * k;
    for (int c = 0; c < k; ++c) {
      bd[c] = INFINITY;
      bi[c] = 0;
    }
    const uint8_t* a = query + q * nb;
    for (int64_t j = 0; j < Ft; ++j) {
      float h = (float)hamming_bytes(a, train + j * nb, nb);
      if (h < bd[k - 1]) topk_insert_f(bd, bi, k, h, (uint32_t)(j + 1));
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * A8  nearest2HammingExhaustiveMEX.cpp:42-79 / ...OMPMEX.cpp:47-82 restated (row-major input):
 *   best = FIRST index attaining the minimum; second = 2nd smallest value WITH multiplicity;
 *   N2==0 -> idx 0, NaN, NaN (:42-45);  N2==1 -> second = nb*8 (:71-74).
 *   The uint16 accumulator of the reference (:56) is kept (wraps for nb*8 > 65535).
 * ---------------------------------------------------------------------------------------- */
void orc_nearest2_hamming(const uint8_t* A, int64_t N1, const uint8_t* B, int64_t N2, int nb, uint32_t* idx2,
                          float* d1, float* d2) {
  if (N2 == 0) {
    for (int64_t i = 0; i < N1; ++i) {
      idx2[i] = 0;
      d1[i] = NAN;
      d2[i] = NAN;
    }
    return;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N1; ++i) {
    uint16_t best = 0xFFFF, second = 0xFFFF;
    int64_t ibest = -1, isecond = -1;
    for (int64_t j = 0; j < N2; ++j) {
      uint16_t h = (uint16_t)hamming_bytes(A + i * nb, B + j * nb, nb);
      if (h < best) {
        second = best;
        isecond = ibest;
        best = h;
        ibest = j;
      } else if (h <= second && j != ibest) {
        second = h;
        isecond = j;
      }
    }
    if (N2 == 1 || isecond < 0) second = (uint16_t)(nb * 8);
    idx2[i] = (uint32_t)(ibest + 1);
    d1[i] = (float)best;
    d2[i] = (float)second;
  }
}

/* ------------------------------------------------------------------------------------------
 * A6  nearest2SSDExhaustive  matchFeaturesScratch.m:322-366 (float32 inputs):
 *   a2 = sum(A.^2,2); b2 = sum(B.^2,2); G = A*B.'; D2 = a2 + b2.' - 2*G   (:351-354)
 *   [best,idx] = min(D2,[],2) (first on ties); mask; second = min(D2,[],2)  (:356-358)
 *   No clamp at zero.  The reference uses N2 without assigning it (:343) -- restated with the
 *   intended N2 = size(B,1); the row blocking (:343-349) does not change any value.
 *   Fixed summation order (the reference's is BLAS-dependent): a2,b2,G sequential over d,
 *   one rounding per operation, D2 = (a2 + b2) - (2*G).
 *   N2==0: MATLAB's validateattributes 'nonempty' (matchFeaturesScratch.m:281-282) throws; here
 *   idx 0 / +inf / +inf (documented deviation).  N2==1: second = +inf (min of all-inf row).
 * ---------------------------------------------------------------------------------------- */
void orc_nearest2_ssd(const float* A, int64_t N1, const float* B, int64_t N2, int D, uint32_t* idx2, float* d1,
                      float* d2) {
  float* b2 = (float*)malloc((size_t)(N2 > 0 ? N2 : 1) * sizeof(float));
  float* T = (float*)malloc((size_t)(N2 > 0 ? N2 : 1) * D * sizeof(float)); /* [D][N2] */
  for (int64_t j = 0; j < N2; ++j) {
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
      float v = B[j * D + d];
      T[(int64_t)d * N2 + j] = v;
      float sq = v * v;
      s = s + sq;
    }
    b2[j] = s;
  }
  const int JB = 512;
#pragma omp parallel
  {
    float* g = (float*)malloc(JB * sizeof(float));
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < N1; ++i) {
      const float* a = A + i * D;
      float a2 = 0.0f;
      for (int d = 0; d < D; ++d) {
        float sq = a[d] * a[d];
        a2 = a2 + sq;
      }
      float best = INFINITY, second = INFINITY;
      int64_t ibest = -1;
      for (int64_t j0 = 0; j0 < N2; j0 += JB) {
        int nj = (int)((N2 - j0 < JB) ? (N2 - j0) : JB);
        for (int j = 0; j < nj; ++j) g[j] = 0.0f;
        for (int d = 0; d < D; ++d) {
          const float ad = a[d];
          const float* t = T + (int64_t)d * N2 + j0;
#pragma omp simd
          for (int j = 0; j < nj; ++j) {
            float p = ad * t[j];
            g[j] = g[j] + p;
          }
        }
        for (int j = 0; j < nj; ++j) {
          float s = a2 + b2[j0 + j];
          float tg = 2.0f * g[j];
          float v = s - tg;
          if (v < best) {
            second = best;
            best = v;
            ibest = j0 + j;
          } else if (v < second) {
            second = v;
          }
        }
      }
      idx2[i] = (uint32_t)(ibest + 1);
      d1[i] = best;
      d2[i] = second;
    }
    free(g);
  }
  free(T);
  free(b2);
}

/* ------------------------------------------------------------------------------------------
 * A3  per-feature filter loop  featureMatchingGlobal.m:123-161  (+ bookkeeping :89-97).
 *   in : idx/dist ROW-major [F][k] (1-based idx, 0 = missing), counts[n] features per image
 *   out: target_img[F]  1-based matched image j, 0 = query rejected
 *        partner[F]     1-based local index of the match inside image j
 *   ratio test (:145): single(d1)/max(single(d2),eps('single')) > ratioThr  -> reject.
 *   MATLAB compares single with double after converting the double to single; ratio_mode 0
 *   restates that (threshold rounded to float), ratio_mode 1 compares in double.  Queries where
 *   the two disagree are the documented ratio ties (returned count in *n_ambiguous).
 *   Deviation: idx==0 (k > F) makes MATLAB error at :135; here such neighbours are dropped.
 * ---------------------------------------------------------------------------------------- */
void orc_global_filter(const uint32_t* idx, const float* dist, int64_t F, int k, const int64_t* counts, int n,
                       double ratioThr, int ratio_mode, int32_t* target_img, uint32_t* partner,
                       int64_t* n_ambiguous) {
  int32_t* imgIdx = (int32_t*)malloc((size_t)(F > 0 ? F : 1) * sizeof(int32_t));
  uint32_t* localIdx = (uint32_t*)malloc((size_t)(F > 0 ? F : 1) * sizeof(uint32_t));
  int64_t p = 0;
  for (int i = 0; i < n; ++i)
    for (int64_t l = 0; l < counts[i]; ++l) {
      imgIdx[p] = i + 1;
      localIdx[p] = (uint32_t)(l + 1);
      ++p;
    }
  int64_t amb = 0;
  for (int64_t q = 0; q < F; ++q) {
    target_img[q] = 0;
    partner[q] = 0;
    const int32_t qi = imgIdx[q];
    float sd[2];
    uint32_t si[2];
    int ns = 0;
    for (int c = 0; c < k && ns < 2; ++c) {
      uint32_t j = idx[q * k + c];
      if (j == 0) continue;                  /* missing neighbour (deviation, see above) */
      if (j == (uint32_t)(q + 1)) continue;  /* :130 self by INDEX */
      if (imgIdx[j - 1] == qi) continue;     /* :135 same image */
      sd[ns] = dist[q * k + c];
      si[ns] = j;
      ++ns;
    }
    if (ns < 2) {
      /* survivors beyond the first two never matter, but the count must reach 2 (:140) */
      continue;
    }
    float den = sd[1] > ORC_EPS32 ? sd[1] : ORC_EPS32; /* max(d2, eps) */
    float ratio = sd[0] / den;
    int rej_single = ratio > (float)ratioThr;
    int rej_double = (double)ratio > ratioThr;
    if (rej_single != rej_double) ++amb;
    if (ratio_mode == 0 ? rej_single : rej_double) continue;
    target_img[q] = imgIdx[si[0] - 1];
    partner[q] = localIdx[si[0] - 1];
  }
  if (n_ambiguous) *n_ambiguous = amb;
  free(imgIdx);
  free(localIdx);
}

/* ------------------------------------------------------------------------------------------
 * A3 (scatter)  featureMatchingGlobal.m:149-159 : matches{min,max}(end+1,:) in loop order.
 *   CSR output: pair_ptr[n*n+1] over MATLAB's column-major linear cell index (a-1)+(b-1)*n,
 *   rows[2*M] interleaved (col1,col2) = (local idx in lower image, local idx in higher image).
 *   Returns M.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_global_scatter(const int32_t* target_img, const uint32_t* partner, int64_t F, const int64_t* counts,
                           int n, int64_t* pair_ptr, uint32_t* rows) {
  int64_t* cnt = (int64_t*)calloc((size_t)n * n + 1, sizeof(int64_t));
  int64_t q = 0;
  for (int i = 0; i < n; ++i)
    for (int64_t l = 0; l < counts[i]; ++l, ++q)
      if (target_img[q] > 0) {
        int qi = i + 1, j = target_img[q];
        int a = qi < j ? qi : j, b = qi < j ? j : qi;
        cnt[(a - 1) + (int64_t)(b - 1) * n]++;
      }
  pair_ptr[0] = 0;
  for (int64_t c = 0; c < (int64_t)n * n; ++c) pair_ptr[c + 1] = pair_ptr[c] + cnt[c];
  int64_t M = pair_ptr[(int64_t)n * n];
  memset(cnt, 0, ((size_t)n * n + 1) * sizeof(int64_t));
  q = 0;
  for (int i = 0; i < n; ++i)
    for (int64_t l = 0; l < counts[i]; ++l, ++q)
      if (target_img[q] > 0) {
        int qi = i + 1, j = target_img[q];
        uint32_t li = (uint32_t)(l + 1), lj = partner[q];
        int a = qi < j ? qi : j, b = qi < j ? j : qi;
        int64_t cell = (a - 1) + (int64_t)(b - 1) * n;
        int64_t pos = pair_ptr[cell] + cnt[cell]++;
        if (qi < j) {
          rows[2 * pos] = li;
          rows[2 * pos + 1] = lj;
        } else {
          rows[2 * pos] = lj;
          rows[2 * pos + 1] = li;
        }
      }
  free(cnt);
  (void)F;
  return M;
}

/* ------------------------------------------------------------------------------------------
 * B.1 whole  featureMatchingGlobal.m:41-161 with an exact kNN (float: desc is copied, normalised
 * per A1, searched per A2; binary: useBF path).  desc = pooled ROW-major [F x D] (float) or
 * [F x nb] (uint8).  Optional outputs knn_idx/knn_dist [F][k] (may be NULL).
 * ---------------------------------------------------------------------------------------- */
int64_t orc_feature_matching_global(const void* desc, int is_binary, const int64_t* counts, int n, int D, int k,
                                    double ratioThr, int ratio_mode, int64_t* pair_ptr, uint32_t* rows,
                                    uint32_t* knn_idx, float* knn_dist, int64_t* n_ambiguous) {
  int64_t F = 0;
  for (int i = 0; i < n; ++i) F += counts[i];
  for (int64_t c = 0; c <= (int64_t)n * n; ++c) pair_ptr[c] = 0;
  if (n_ambiguous) *n_ambiguous = 0;
  if (F == 0) return 0; /* :49-52, :65-67 */
  uint32_t* idx = knn_idx ? knn_idx : (uint32_t*)malloc((size_t)F * k * sizeof(uint32_t));
  float* dist = knn_dist ? knn_dist : (float*)malloc((size_t)F * k * sizeof(float));
  if (is_binary) {
    orc_knn_hamming((const uint8_t*)desc, F, (const uint8_t*)desc, F, D, k, idx, dist);
  } else {
    float* X = (float*)malloc((size_t)F * D * sizeof(float));
    memcpy(X, desc, (size_t)F * D * sizeof(float));
    orc_normalize_rows_global(X, F, D);
    orc_knn_l2(X, F, X, F, D, k, idx, dist);
    free(X);
  }
  int32_t* tgt = (int32_t*)malloc((size_t)F * sizeof(int32_t));
  uint32_t* par = (uint32_t*)malloc((size_t)F * sizeof(uint32_t));
  orc_global_filter(idx, dist, F, k, counts, n, ratioThr, ratio_mode, tgt, par, n_ambiguous);
  int64_t M = orc_global_scatter(tgt, par, F, counts, n, pair_ptr, rows);
  free(tgt);
  free(par);
  if (!knn_idx) free(idx);
  if (!knn_dist) free(dist);
  return M;
}

/* ------------------------------------------------------------------------------------------
 * A7  ratio / threshold / unique  matchFeaturesScratch.m:169-215.
 *   float : dBest,dSecond are single-valued doubles (:344), r2 = MaxRatio^2 in double (:173-174)
 *   binary: dBest = (d1/nBits)*100 in SINGLE (d1 single, :120-121), compared with
 *           MaxRatio*dSecond evaluated in single (mixed single/double arithmetic is single).
 *   keep = ratioOK & dBest<=MatchThreshold & isfinite(dBest) & isfinite(dSecond)   (:177-178)
 *   Unique (:186-204): stable ascending sort by d, greedy accept if neither side used.
 *   out: matches[2*K] interleaved (query idx, train idx) 1-based, metric[K] (double), returns K.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double d;
  int64_t i1;
  uint32_t i2;
} orc_cand;

static void orc_merge_sort(orc_cand* a, orc_cand* tmp, int64_t n) { /* stable */
  if (n < 2) return;
  int64_t h = n / 2;
  orc_merge_sort(a, tmp, h);
  orc_merge_sort(a + h, tmp, n - h);
  int64_t i = 0, j = h, o = 0;
  while (i < h && j < n) tmp[o++] = (a[j].d < a[i].d) ? a[j++] : a[i++];
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(orc_cand));
}

int64_t orc_filter_unique(const uint32_t* idx2, const float* d1, const float* d2, int64_t N1, int64_t N2,
                          int is_binary, int nBits, double matchThreshold, double maxRatio, int unique,
                          uint32_t* matches, double* metric) {
  orc_cand* c = (orc_cand*)malloc((size_t)(N1 > 0 ? N1 : 1) * sizeof(orc_cand));
  int64_t m = 0;
  for (int64_t i = 0; i < N1; ++i) {
    int keep;
    double dB;
    if (is_binary) {
      float s = d2[i];
      if (!isfinite(s) || s == 0.0f) s = (float)nBits; /* matchFeaturesScratch.m:318 */
      float fB = (d1[i] / (float)nBits) * 100.0f;
      float fS = (s / (float)nBits) * 100.0f;
      float rhs = (float)maxRatio * fS;
      keep = (fB <= rhs) && (fB <= (float)matchThreshold) && isfinite(fB) && isfinite(fS);
      dB = (double)fB;
    } else {
      double b = (double)d1[i], s = (double)d2[i];
      double r2 = maxRatio * maxRatio;
      keep = (b <= r2 * s) && (b <= matchThreshold) && isfinite(b) && isfinite(s);
      dB = b;
    }
    if (keep) {
      c[m].d = dB;
      c[m].i1 = i + 1;
      c[m].i2 = idx2[i];
      ++m;
    }
  }
  int64_t K = 0;
  if (unique && m > 0) {
    orc_cand* tmp = (orc_cand*)malloc((size_t)m * sizeof(orc_cand));
    orc_merge_sort(c, tmp, m);
    free(tmp);
    uint8_t* used2 = (uint8_t*)calloc((size_t)N2 + 1, 1);
    for (int64_t t = 0; t < m; ++t) {
      if (!used2[c[t].i2]) { /* used1 can never be set twice: every query occurs once */
        used2[c[t].i2] = 1;
        matches[2 * K] = (uint32_t)c[t].i1;
        matches[2 * K + 1] = c[t].i2;
        metric[K] = c[t].d;
        ++K;
      }
    }
    free(used2);
  } else {
    for (int64_t t = 0; t < m; ++t) {
      matches[2 * K] = (uint32_t)c[t].i1;
      matches[2 * K + 1] = c[t].i2;
      metric[K] = c[t].d;
      ++K;
    }
  }
  free(c);
  return K;
}

/* ------------------------------------------------------------------------------------------
 * A5+A6/A8+A7  matchFeaturesScratch(F1,F2,'Method','Exhaustive',...)  :55-215 for one pair.
 *   float : A,B row-major [N x D] float32; normalised per :105-110 iff max|.|>2.
 *   binary: packed uint8 rows [N x nb], nBits = 8*nb (binaryFeatures, :252-257).
 *   Empty inputs -> 0 matches (binary :84-88; float: documented deviation, MATLAB throws).
 * ---------------------------------------------------------------------------------------- */
int64_t orc_match_features(const void* A, int64_t N1, const void* B, int64_t N2, int D, int is_binary,
                           double matchThreshold, double maxRatio, int unique, uint32_t* matches,
                           double* metric) {
  if (N1 == 0 || N2 == 0) return 0;
  uint32_t* idx2 = (uint32_t*)malloc((size_t)N1 * sizeof(uint32_t));
  float* d1 = (float*)malloc((size_t)N1 * sizeof(float));
  float* d2 = (float*)malloc((size_t)N1 * sizeof(float));
  int64_t K;
  if (is_binary) {
    orc_nearest2_hamming((const uint8_t*)A, N1, (const uint8_t*)B, N2, D, idx2, d1, d2);
    K = orc_filter_unique(idx2, d1, d2, N1, N2, 1, D * 8, matchThreshold, maxRatio, unique, matches, metric);
  } else {
    float* a = (float*)malloc((size_t)N1 * D * sizeof(float));
    float* b = (float*)malloc((size_t)N2 * D * sizeof(float));
    memcpy(a, A, (size_t)N1 * D * sizeof(float));
    memcpy(b, B, (size_t)N2 * D * sizeof(float));
    if (orc_needs_normalization(a, N1 * D, b, N2 * D)) {
      orc_normalize_rows_pairwise(a, N1, D);
      orc_normalize_rows_pairwise(b, N2, D);
    }
    orc_nearest2_ssd(a, N1, b, N2, D, idx2, d1, d2);
    K = orc_filter_unique(idx2, d1, d2, N1, N2, 0, 0, matchThreshold, maxRatio, unique, matches, metric);
    free(a);
    free(b);
  }
  free(idx2);
  free(d1);
  free(d2);
  return K;
}

/* ------------------------------------------------------------------------------------------
 * A10 (the two modes that are exact searches)  matchFeaturesScratch.m:142-155:
 *   'kdtree'       nearest2KDTree :411-440    knnsearch(createns(B,'kdtree',...), A, 'K', 2) -- an EXACT Euclidean search
 *   'subsetpdist2' nearest2SubsetPdist2 :370-409   pdist2(B(candB,:), A, 'euclidean', 'Smallest', 2), candB =
 *                  randperm(N2, min(subset, N2)): while N2 <= subset (12000) every row of B is a candidate
 *   both return EUCLIDEAN distances and the caller squares them: dBest = d1.^2, dSecond = d2.^2 (:146-147, :153-154).
 *   Restated: s = sum((a-b).^2) sequential float32, r = sqrtf(s), ranking by (r, index) [knnsearch / pdist2 are
 *   closed source: their summation order and tie order are unpinned; candB's permutation only reorders exact ties],
 *   outputs d1 = r1*r1, d2 = r2*r2 in float32.  N2 == 1: second = +inf (deviation: pdist2 path pads with
 *   d1 + eps(d1), :396-400; knnsearch errors).
 * ---------------------------------------------------------------------------------------- */
void orc_nearest2_euclid(const float* A, int64_t N1, const float* B, int64_t N2, int D, uint32_t* idx2, float* d1,
                         float* d2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N1; ++i) {
    const float* a = A + i * D;
    float r1 = INFINITY, r2 = INFINITY;
    uint32_t i1 = 0;
    for (int64_t j = 0; j < N2; ++j) {
      const float* b = B + j * D;
      float s = 0.0f;
      for (int d = 0; d < D; ++d) {
        const float e = a[d] - b[d];
        s = s + e * e;
      }
      const float r = sqrtf(s);
      if (r < r1) {
        r2 = r1;
        r1 = r;
        i1 = (uint32_t)(j + 1);
      } else if (r < r2) {
        r2 = r;
      }
    }
    idx2[i] = i1;
    d1[i] = i1 ? r1 * r1 : INFINITY;
    d2[i] = isinf(r2) ? INFINITY : r2 * r2;
  }
}

/* matchFeaturesScratch(F1,F2,'Method','Approximate','ApproxFloatNNMethod', m, ...) for float descriptors,
 * m = 1 'subsetpdist2' (N2 <= subset), 2 'kdtree'; binary descriptors and m = 0 take the exhaustive path. */
int64_t orc_match_features_method(const void* A, int64_t N1, const void* B, int64_t N2, int D, int is_binary, int method,
                                  double matchThreshold, double maxRatio, int unique, uint32_t* matches,
                                  double* metric) {
  if (is_binary || method == 0)
    return orc_match_features(A, N1, B, N2, D, is_binary, matchThreshold, maxRatio, unique, matches, metric);
  if (N1 == 0 || N2 == 0) return 0;
  uint32_t* idx2 = (uint32_t*)malloc((size_t)N1 * sizeof(uint32_t));
  float* d1 = (float*)malloc((size_t)N1 * sizeof(float));
  float* d2 = (float*)malloc((size_t)N1 * sizeof(float));
  float* a = (float*)malloc((size_t)N1 * D * sizeof(float));
  float* b = (float*)malloc((size_t)N2 * D * sizeof(float));
  memcpy(a, A, (size_t)N1 * D * sizeof(float));
  memcpy(b, B, (size_t)N2 * D * sizeof(float));
  if (orc_needs_normalization(a, N1 * D, b, N2 * D)) { /* :105-110 runs before the method switch */
    orc_normalize_rows_pairwise(a, N1, D);
    orc_normalize_rows_pairwise(b, N2, D);
  }
  orc_nearest2_euclid(a, N1, b, N2, D, idx2, d1, d2);
  const int64_t K = orc_filter_unique(idx2, d1, d2, N1, N2, 0, 0, matchThreshold, maxRatio, unique, matches, metric);
  free(a);
  free(b);
  free(idx2);
  free(d1);
  free(d2);
  return K;
}

/* 'subsetpdist2' with N2 > subset (matchFeaturesScratch.m:388-408): candB [subset] = 0-based rows of B in the order
 * randperm returned them (here: the table the GPU drew, aps_pplan_subset_table -- MATLAB's stream cannot be reproduced);
 * B2 = B(candB,:), Euclidean 2-NN over B2 (ties -> first row of B2), idx2 = candB(I). */
int64_t orc_match_features_subset(const float* A, int64_t N1, const float* B, int64_t N2, int D, const int32_t* candB,
                                  int64_t subset, double matchThreshold, double maxRatio, int unique, uint32_t* matches,
                                  double* metric) {
  if (N1 == 0 || N2 == 0 || subset == 0) return 0;
  uint32_t* idx2 = (uint32_t*)malloc((size_t)N1 * sizeof(uint32_t));
  float* d1 = (float*)malloc((size_t)N1 * sizeof(float));
  float* d2 = (float*)malloc((size_t)N1 * sizeof(float));
  float* a = (float*)malloc((size_t)N1 * D * sizeof(float));
  float* b = (float*)malloc((size_t)N2 * D * sizeof(float));
  float* b2 = (float*)malloc((size_t)subset * D * sizeof(float));
  memcpy(a, A, (size_t)N1 * D * sizeof(float));
  memcpy(b, B, (size_t)N2 * D * sizeof(float));
  if (orc_needs_normalization(a, N1 * D, b, N2 * D)) { /* :105-110: on the FULL inputs, before the method switch */
    orc_normalize_rows_pairwise(a, N1, D);
    orc_normalize_rows_pairwise(b, N2, D);
  }
  for (int64_t r = 0; r < subset; ++r) memcpy(b2 + r * D, b + (size_t)candB[r] * D, (size_t)D * sizeof(float));
  orc_nearest2_euclid(a, N1, b2, subset, D, idx2, d1, d2);
  /* uniqueness is decided per candidate row; positions in B2 and rows of B are in bijection, so filter first, map after */
  const int64_t K = orc_filter_unique(idx2, d1, d2, N1, subset, 0, 0, matchThreshold, maxRatio, unique, matches, metric);
  for (int64_t t = 0; t < K; ++t) matches[2 * t + 1] = (uint32_t)candB[matches[2 * t + 1] - 1] + 1u; /* :406 */
  free(a);
  free(b);
  free(b2);
  free(idx2);
  free(d1);
  free(d2);
  return K;
}

/* ------------------------------------------------------------------------------------------
 * A10 'pca2nn'  matchFeaturesScratch.m:130-141 -> nearest2ApproxFloatFast :442-534 + doBlock :536-573.
 *   D > 48: muB = mean(B,1); coeff = pca(B - muB, 'NumComponents', 48); B = (B-muB)*coeff; A = (A-muB)*coeff  (:476-483)
 *   rows re-normalised, x ./ (sqrt(sum(x.^2,2)) + eps) (:486-487); G = A*B.'; [sim1,id1] = max(G,[],2); mask; sim2 = max;
 *   d = single(2 - 2*sim) (:560-570).  The blocking / parfor (:493-531) does not change any value.
 *   MathWorks' pca (SVD of the centred data) is closed source: parity unpinned.  G and the row norms depend only on
 *   the SUBSPACE of the 48 leading components (not on the basis inside it, nor on component signs), so any exact
 *   eigen-solver gives the same result up to rounding; restated with ONE fixed arithmetic, which csrc/aps_pca.cu repeats
 *   operation by operation: mean float32 sequential; covariance float64 sequential over the rows ((x - mu) formed in
 *   float32); cyclic Jacobi, ORC_PCA_SWEEPS sweeps, rotations (p,q), p<q; eigenvalues descending, ties -> lower index;
 *   coeff = single(V(:, order)), min(48, N2-1) columns (pca returns at most N-1), the rest zero; projection
 *   y_c = sum_d fl(fl(x_d - mu_d) * coeff[d][c]) sequential float32; similarity = sequential float32 dot.
 * ---------------------------------------------------------------------------------------- */
#define ORC_PCA_SWEEPS 12
#define ORC_PCA_COMPONENTS 48
static void orc_pca_basis(const float* B, int64_t N2, int D, int P, float* mu, float* coeff /* [D][P] */) {
  double* A = (double*)calloc((size_t)D * D, sizeof(double));
  double* V = (double*)calloc((size_t)D * D, sizeof(double)); /* V[k][p] */
  for (int d = 0; d < D; ++d) {
    float s = 0.0f;
    for (int64_t r = 0; r < N2; ++r) s = s + B[r * D + d];
    mu[d] = N2 > 0 ? s / (float)N2 : 0.0f;
  }
  for (int a = 0; a < D; ++a)
    for (int b = a; b < D; ++b) {
      double s = 0.0;
      for (int64_t r = 0; r < N2; ++r) {
        const double xa = (double)(float)(B[r * D + a] - mu[a]), xb = (double)(float)(B[r * D + b] - mu[b]);
        s = s + xa * xb;
      }
      A[a * D + b] = s;
      A[b * D + a] = s;
    }
  for (int k = 0; k < D; ++k) V[k * D + k] = 1.0;
  for (int sweep = 0; sweep < ORC_PCA_SWEEPS; ++sweep)
    for (int p = 0; p < D - 1; ++p)
      for (int q = p + 1; q < D; ++q) {
        const double apq = A[p * D + q];
        if (apq == 0.0) continue;
        const double app = A[p * D + p], aqq = A[q * D + q];
        const double theta = (aqq - app) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double s = t * c;
        for (int k = 0; k < D; ++k) {
          const double akp = A[k * D + p], akq = A[k * D + q];
          A[k * D + p] = c * akp - s * akq;
          A[k * D + q] = s * akp + c * akq;
        }
        for (int k = 0; k < D; ++k) {
          const double apk = A[p * D + k], aqk = A[q * D + k];
          A[p * D + k] = c * apk - s * aqk;
          A[q * D + k] = s * apk + c * aqk;
          const double vkp = V[k * D + p], vkq = V[k * D + q];
          V[k * D + p] = c * vkp - s * vkq;
          V[k * D + q] = s * vkp + c * vkq;
        }
      }
  int* order = (int*)malloc((size_t)D * sizeof(int));
  for (int k = 0; k < D; ++k) {
    int r = 0;
    for (int i = 0; i < D; ++i) r += (A[i * D + i] > A[k * D + k]) || (A[i * D + i] == A[k * D + k] && i < k);
    order[r] = k;
  }
  const int Pj = (int)(N2 - 1 < P ? (N2 > 0 ? N2 - 1 : 0) : P);
  for (int d = 0; d < D; ++d)
    for (int c = 0; c < P; ++c) coeff[d * P + c] = c < Pj ? (float)V[d * D + order[c]] : 0.0f;
  free(order);
  free(A);
  free(V);
}
static void orc_pca_project(const float* X, int64_t N, int D, int P, const float* mu, const float* coeff, float* Y) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < N; ++r)
    for (int c = 0; c < P; ++c) {
      float y = 0.0f;
      for (int d = 0; d < D; ++d) {
        const float e = X[r * D + d] - mu[d];
        y = y + e * coeff[d * P + c];
      }
      Y[r * P + c] = y;
    }
}
/* 2-NN by cosine similarity of rows that are already projected and normalised */
void orc_nearest2_cosine(const float* A, int64_t N1, const float* B, int64_t N2, int P, uint32_t* idx2, float* d1, float* d2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N1; ++i) {
    float s1 = -INFINITY, s2 = -INFINITY;
    uint32_t i1 = 0;
    for (int64_t j = 0; j < N2; ++j) {
      float g = 0.0f;
      for (int c = 0; c < P; ++c) g = g + A[i * P + c] * B[j * P + c];
      if (g > s1) {
        s2 = s1;
        s1 = g;
        i1 = (uint32_t)(j + 1);
      } else if (g > s2) {
        s2 = g;
      }
    }
    idx2[i] = i1;
    d1[i] = i1 ? 2.0f - 2.0f * s1 : INFINITY;
    d2[i] = isinf(s2) ? INFINITY : 2.0f - 2.0f * s2;
  }
}
int64_t orc_match_features_pca(const float* Ain, int64_t N1, const float* Bin, int64_t N2, int D, double matchThreshold,
                               double maxRatio, int unique, uint32_t* matches, double* metric) {
  if (N1 == 0 || N2 == 0) return 0;
  float* a = (float*)malloc((size_t)N1 * D * sizeof(float));
  float* b = (float*)malloc((size_t)N2 * D * sizeof(float));
  memcpy(a, Ain, (size_t)N1 * D * sizeof(float));
  memcpy(b, Bin, (size_t)N2 * D * sizeof(float));
  if (orc_needs_normalization(a, N1 * D, b, N2 * D)) { /* :105-110, before the method switch */
    orc_normalize_rows_pairwise(a, N1, D);
    orc_normalize_rows_pairwise(b, N2, D);
  }
  const int P = D > ORC_PCA_COMPONENTS ? ORC_PCA_COMPONENTS : D;
  float* ap = a;
  float* bp = b;
  if (D > ORC_PCA_COMPONENTS) { /* :477 */
    float* mu = (float*)malloc((size_t)D * sizeof(float));
    float* coeff = (float*)malloc((size_t)D * P * sizeof(float));
    orc_pca_basis(b, N2, D, P, mu, coeff);
    ap = (float*)malloc((size_t)N1 * P * sizeof(float));
    bp = (float*)malloc((size_t)N2 * P * sizeof(float));
    orc_pca_project(a, N1, D, P, mu, coeff, ap);
    orc_pca_project(b, N2, D, P, mu, coeff, bp);
    free(mu);
    free(coeff);
  }
  orc_normalize_rows_pairwise(ap, N1, P); /* :486-487 */
  orc_normalize_rows_pairwise(bp, N2, P);
  uint32_t* idx2 = (uint32_t*)malloc((size_t)N1 * sizeof(uint32_t));
  float* d1 = (float*)malloc((size_t)N1 * sizeof(float));
  float* d2 = (float*)malloc((size_t)N1 * sizeof(float));
  orc_nearest2_cosine(ap, N1, bp, N2, P, idx2, d1, d2);
  const int64_t K = orc_filter_unique(idx2, d1, d2, N1, N2, 0, 0, matchThreshold, maxRatio, unique, matches, metric);
  if (ap != a) free(ap);
  if (bp != b) free(bp);
  free(a);
  free(b);
  free(idx2);
  free(d1);
  free(d2);
  return K;
}

/* ------------------------------------------------------------------------------------------
 * A4 / B.2 whole  featureMatchingPairwise.m:43-63 + getMatches :103-120 (useMATLABFeatureMatch=0,
 * Matchingmethod='Exhaustive'): every (i<j), query = image i, train = image j, Unique=true.
 *   desc pooled row-major; CSR output like orc_global_scatter (pair cell (i,j) i<j), rows =
 *   (idx in i, idx in j), metric[M] optional (may be NULL).  rows capacity: sum_i N_i*(n-1-i).
 * ---------------------------------------------------------------------------------------- */
int64_t orc_feature_matching_pairwise(const void* desc, int is_binary, const int64_t* counts, int n, int D,
                                      double matchThreshold, double maxRatio, int64_t* pair_ptr, uint32_t* rows,
                                      double* metric) {
  int64_t* off = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
  off[0] = 0;
  int64_t maxN = 0;
  for (int i = 0; i < n; ++i) {
    off[i + 1] = off[i] + counts[i];
    if (counts[i] > maxN) maxN = counts[i];
  }
  const size_t esz = is_binary ? 1 : sizeof(float);
  uint32_t* mm = (uint32_t*)malloc((size_t)(maxN > 0 ? maxN : 1) * 2 * sizeof(uint32_t));
  double* met = (double*)malloc((size_t)(maxN > 0 ? maxN : 1) * sizeof(double));
  int64_t M = 0;
  pair_ptr[0] = 0;
  /* CSR runs over the column-major linear cell index c = i + j*n; fill in that order */
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      int64_t cell = i + (int64_t)j * n;
      int64_t K = 0;
      if (i < j) {
        const char* A = (const char*)desc + (size_t)off[i] * D * esz;
        const char* B = (const char*)desc + (size_t)off[j] * D * esz;
        K = orc_match_features(A, counts[i], B, counts[j], D, is_binary, matchThreshold, maxRatio, 1, mm, met);
        memcpy(rows + 2 * M, mm, (size_t)K * 2 * sizeof(uint32_t));
        if (metric) memcpy(metric + M, met, (size_t)K * sizeof(double));
      }
      M += K;
      pair_ptr[cell + 1] = M;
    }
  free(mm);
  free(met);
  free(off);
  return M;
}

/* ------------------------------------------------------------------------------------------
 * A9  top-m partner selection  imageMatching.m:75-100.
 *   counts = putativeCount [n x n] COLUMN-major (cell (i,j) at i + j*n), cand out same layout.
 *   symCounts = C + C', zero diagonal (:82-83); stable descending sort per row (:86) -> ties to
 *   the lower column; first min(m,n-1) columns (:87); OR-symmetrise, strict upper (:94-96).
 *   Returns the number of candidate pairs; pairs_lin (optional) gets find(cand) 0-based linear
 *   indices in column-major order (:99).
 * ---------------------------------------------------------------------------------------- */
int64_t orc_select_partners(const int64_t* counts, int n, int m, uint8_t* cand, int64_t* pairs_lin) {
  int64_t* S = (int64_t*)malloc((size_t)n * n * sizeof(int64_t));
  uint8_t* P = (uint8_t*)calloc((size_t)n * n, 1);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S[i + (int64_t)j * n] = (i == j) ? 0 : counts[i + (int64_t)j * n] + counts[j + (int64_t)i * n];
  int take = m < n - 1 ? m : n - 1;
  if (take < 0) take = 0;
  int* order = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) order[j] = j;
    /* stable insertion sort, descending */
    for (int a = 1; a < n; ++a) {
      int v = order[a];
      int64_t sv = S[i + (int64_t)v * n];
      int b = a - 1;
      while (b >= 0 && S[i + (int64_t)order[b] * n] < sv) {
        order[b + 1] = order[b];
        --b;
      }
      order[b + 1] = v;
    }
    for (int t = 0; t < take; ++t) P[i + (int64_t)order[t] * n] = 1;
  }
  int64_t np = 0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      uint8_t v = (i < j) && (P[i + (int64_t)j * n] || P[j + (int64_t)i * n]);
      cand[i + (int64_t)j * n] = v;
      if (v) {
        if (pairs_lin) pairs_lin[np] = i + (int64_t)j * n;
        ++np;
      }
    }
  free(order);
  free(P);
  free(S);
  return np;
}

/* matchFeaturesScratch.m:617-646 packBits: MSB-first packing of 0/1 columns into bytes. */
void orc_pack_bits(const uint8_t* bits01, int64_t N, int Dbits, uint8_t* packed) {
  int nbytes = (Dbits + 7) / 8;
  memset(packed, 0, (size_t)N * nbytes);
  for (int64_t r = 0; r < N; ++r)
    for (int b = 0; b < Dbits; ++b)
      if (bits01[r * Dbits + b]) packed[r * nbytes + b / 8] |= (uint8_t)(1u << (7 - (b % 8)));
}
