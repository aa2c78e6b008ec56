// aps_pca.cu -- 'pca2nn': the PCA front end of the approximate float matcher.
//
// Replaces  PP/featureMatching/matchFeaturesScratch.m:476-483 (nearest2ApproxFloatFast, step 1):
//     muB = mean(B,1);  coeff = pca(B - muB, 'NumComponents', 48);  B = (B - muB)*coeff;  A = (A - muB)*coeff;
// (only when D > 48).  MathWorks' pca is closed source (SVD of the centred data); what matters downstream is the
// SUBSPACE of the leading components -- A'B'^T and the row norms do not depend on the basis chosen inside it, nor on the
// sign of a component (negating a column of coeff negates the projected component exactly).  Restated here with ONE fixed
// arithmetic, the same as the oracle's, so that both produce the same bits:
//   mean        float32, sequential over the rows
//   covariance  C = sum_r (x_r - mu)(x_r - mu)^T in float64, sequential over the rows (x - mu formed in float32)
//   eigenvectors cyclic-by-row Jacobi in float64, PCA_SWEEPS sweeps, rotations in the order (p, q), p < q; every
//               operation a single IEEE operation (no FMA); eigenvalues sorted descending (ties -> lower index)
//   coeff       float32( V(:, order(1:P)) ), P = min(48, N - 1) columns, the remaining of the 48 are zero
//   projection  y_c = sum_d fl( fl(x_d - mu_d) * coeff[d][c] ), float32, sequential over d
// One CTA per image for the basis (the matrix lives in shared memory, thread k owns row / column k of a rotation); the
// projections are table-driven (a segment = rows of one image projected with the basis of another).
#include "aps_common.cuh"

namespace {

constexpr int PCA_SWEEPS = 12;

__global__ void k_pca_mean(const float* __restrict__ X, const int64_t* __restrict__ img_off, int D, float* __restrict__ mu) {
  const int j = blockIdx.x;
  const int64_t r0 = img_off[j], r1 = img_off[j + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s = __fadd_rn(s, X[r * D + d]);
    mu[(int64_t)j * D + d] = r1 > r0 ? __fdiv_rn(s, (float)(r1 - r0)) : 0.f;   // mean(B,1), single
  }
}

// C[a][b] = sum_r (x_ra - mu_a)(x_rb - mu_b), float64, sequential over r; one thread per entry (a <= b), mirrored
__global__ void k_pca_cov(const float* __restrict__ X, const int64_t* __restrict__ img_off, int D,
                          const float* __restrict__ mu, double* __restrict__ C) {
  const int j = blockIdx.y;
  const int64_t r0 = img_off[j], r1 = img_off[j + 1];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D * D) return;
  const int a = e / D, b = e - a * D;
  if (a > b) return;
  const float ma = mu[(int64_t)j * D + a], mb = mu[(int64_t)j * D + b];
  double s = 0.0;
  for (int64_t r = r0; r < r1; ++r) {
    const double xa = (double)__fsub_rn(X[r * D + a], ma), xb = (double)__fsub_rn(X[r * D + b], mb);
    s = __dadd_rn(s, __dmul_rn(xa, xb));
  }
  double* Cj = C + (int64_t)j * D * D;
  Cj[a * D + b] = s;
  Cj[b * D + a] = s;
}

// cyclic Jacobi; A in shared memory [D][D+1], V^T in global memory (Vt[p][k] = V[k][p]: coalesced for thread k)
__global__ void k_pca_jacobi(double* __restrict__ C, double* __restrict__ Vt_all, const int64_t* __restrict__ img_off,
                             int D, int P, float* __restrict__ coeff) {
  extern __shared__ double sm[];
  const int ld = D + 1;
  double* A = sm;               // [D][ld]
  double* ev = A + D * ld;      // [D] eigenvalues
  int* order = (int*)(ev + D);  // [D]
  const int j = blockIdx.x, k = threadIdx.x;   // blockDim.x == D
  double* Cj = C + (int64_t)j * D * D;
  double* Vt = Vt_all + (int64_t)j * D * D;
  for (int i = 0; i < D; ++i) {
    A[i * ld + k] = Cj[i * D + k];
    Vt[i * D + k] = (i == k) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int sweep = 0; sweep < PCA_SWEEPS; ++sweep)
    for (int p = 0; p < D - 1; ++p)
      for (int q = p + 1; q < D; ++q) {
        const double apq = A[p * ld + q];
        if (apq == 0.0) continue;   // uniform over the CTA
        const double app = A[p * ld + p], aqq = A[q * ld + q];
        const double theta = __ddiv_rn(__dsub_rn(aqq, app), __dmul_rn(2.0, apq));
        const double at = fabs(theta);
        double t = __ddiv_rn(1.0, __dadd_rn(at, __dsqrt_rn(__dadd_rn(__dmul_rn(theta, theta), 1.0))));
        if (theta < 0.0) t = -t;
        const double c = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(t, t), 1.0)));
        const double s = __dmul_rn(t, c);
        __syncthreads();   // everybody has read app, aqq, apq
        // columns p, q of row k
        const double akp = A[k * ld + p], akq = A[k * ld + q];
        A[k * ld + p] = __dsub_rn(__dmul_rn(c, akp), __dmul_rn(s, akq));
        A[k * ld + q] = __dadd_rn(__dmul_rn(s, akp), __dmul_rn(c, akq));
        __syncthreads();
        // rows p, q of column k ; eigenvector columns p, q
        const double apk = A[p * ld + k], aqk = A[q * ld + k];
        A[p * ld + k] = __dsub_rn(__dmul_rn(c, apk), __dmul_rn(s, aqk));
        A[q * ld + k] = __dadd_rn(__dmul_rn(s, apk), __dmul_rn(c, aqk));
        const double vkp = Vt[p * D + k], vkq = Vt[q * D + k];
        Vt[p * D + k] = __dsub_rn(__dmul_rn(c, vkp), __dmul_rn(s, vkq));
        Vt[q * D + k] = __dadd_rn(__dmul_rn(s, vkp), __dmul_rn(c, vkq));
        __syncthreads();
      }
  ev[k] = A[k * ld + k];
  __syncthreads();
  // rank of eigenvalue k in descending order, ties -> lower index
  int r = 0;
  for (int i = 0; i < D; ++i) r += (ev[i] > ev[k]) || (ev[i] == ev[k] && i < k);
  order[r] = k;
  __syncthreads();
  const int64_t rows = img_off[j + 1] - img_off[j];
  const int Pj = (int)min((int64_t)P, rows > 0 ? rows - 1 : (int64_t)0);   // pca returns at most N - 1 components
  float* cj = coeff + (int64_t)j * D * P;
  for (int cidx = 0; cidx < P; ++cidx) cj[k * P + cidx] = cidx < Pj ? (float)Vt[order[cidx] * D + k] : 0.f;
}

struct ProjSeg {   // rows [src, src + cnt) of X projected with the basis of image `basis` into rows [dst, dst + cnt)
  int64_t src, dst;
  int32_t cnt, basis;
};

// block = PR_ROWS rows x P components (P <= 64): the (x - mu) rows and the basis staged in shared memory.  Thread =
// (component, row group): the component's basis column is read once per four dimensions and used for eight rows held in
// registers, the rows come as broadcast 128-bit loads -- 0.4 shared-memory loads per multiply-add instead of 2 (the
// first version was bound by the load/store pipe: 182 ms for the 44850 projections of C5).  Every y is still the
// sequential float32 sum over d of the oracle.
constexpr int PR_ROWS = 128;
__global__ void __launch_bounds__(256) k_pca_project(const float* __restrict__ X, int D, int P,  // (4 * P threads)
                                                     const ProjSeg* __restrict__ segs,
                                                     const int64_t* __restrict__ blk_off, int nseg,
                                                     const float* __restrict__ mu, const float* __restrict__ coeff,
                                                     float* __restrict__ out) {
  extern __shared__ __align__(16) float smf[];
  float* cf = smf;               // [D][P]
  float* xs = cf + D * P;        // [PR_ROWS][D]  (x - mu)
  // segment of this block: last s with blk_off[s] <= blockIdx.x
  int lo = 0, hi = nseg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (blk_off[mid] <= (int64_t)blockIdx.x) lo = mid; else hi = mid;
  }
  const ProjSeg sg = segs[lo];
  const int64_t row0 = ((int64_t)blockIdx.x - blk_off[lo]) * PR_ROWS;
  const float* cj = coeff + (int64_t)sg.basis * D * P;
  const float* mj = mu + (int64_t)sg.basis * D;
  const bool vec = ((D & 3) == 0) && (((D * P) & 3) == 0);
  for (int i = threadIdx.x; i < D * P; i += blockDim.x) cf[i] = cj[i];
  if (vec) {   // rows of X are 16-byte aligned then (cudaMalloc'ed [rows x D] matrix)
    const int d4 = D >> 2;
    for (int i = threadIdx.x; i < PR_ROWS * d4; i += blockDim.x) {
      const int r = i / d4, d = (i - r * d4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < sg.cnt) {
        const float4 x = *reinterpret_cast<const float4*>(X + (sg.src + row0 + r) * D + d);
        const float4 m = *reinterpret_cast<const float4*>(mj + d);
        v = make_float4(__fsub_rn(x.x, m.x), __fsub_rn(x.y, m.y), __fsub_rn(x.z, m.z), __fsub_rn(x.w, m.w));
      }
      *reinterpret_cast<float4*>(xs + r * D + d) = v;
    }
  } else {
    for (int i = threadIdx.x; i < PR_ROWS * D; i += blockDim.x) {
      const int r = i / D, d = i - r * D;
      xs[i] = (row0 + r < sg.cnt) ? __fsub_rn(X[(sg.src + row0 + r) * D + d], mj[d]) : 0.f;
    }
  }
  __syncthreads();
  // thread = (component, row group): with P = 48 components the block is launched with 4 * 48 = 192 threads, all busy
  const int cidx = threadIdx.x % P, rg = threadIdx.x / P;
  if (rg >= 4) return;
  for (int c0 = 0; c0 < PR_ROWS && row0 + c0 < sg.cnt; c0 += 32) {
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = 0.f;
    if (vec) {
      for (int d = 0; d < D; d += 4) {
        const float b0 = cf[d * P + cidx], b1 = cf[(d + 1) * P + cidx], b2 = cf[(d + 2) * P + cidx], b3 = cf[(d + 3) * P + cidx];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 x = *reinterpret_cast<const float4*>(xs + (c0 + rg + 4 * k) * D + d);
          y[k] = __fadd_rn(y[k], __fmul_rn(x.x, b0));
          y[k] = __fadd_rn(y[k], __fmul_rn(x.y, b1));
          y[k] = __fadd_rn(y[k], __fmul_rn(x.z, b2));
          y[k] = __fadd_rn(y[k], __fmul_rn(x.w, b3));
        }
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const float bb = cf[d * P + cidx];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = __fadd_rn(y[k], __fmul_rn(xs[(c0 + rg + 4 * k) * D + d], bb));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t r = row0 + c0 + rg + 4 * k;
      if (r < sg.cnt) out[(sg.dst + r) * P + cidx] = y[k];
    }
  }
}

}  // namespace

int aps_pca_components() { return 48; }   // 'ApproxNumComponents', matchFeaturesScratch.m:133

// mu [n x D], coeff [n x D x P] (float), scratch: 2 * n * D * D doubles
int aps_k_pca_basis(cudaStream_t s, const float* X, const int64_t* d_img_off, int n, int D, int P, float* mu, float* coeff,
                    double* scratch) {
  if (n == 0) return APS_OK;
  if (D > 128 || D < 2 || P > 64) {
    aps_set_error(APS_ERR_DIM, "", "pca2nn supports descriptor lengths 2..128 (got %d)", D);
    return APS_ERR_DIM;
  }
  double* C = scratch;
  double* Vt = scratch + (size_t)n * D * D;
  k_pca_mean<<<n, 128, 0, s>>>(X, d_img_off, D, mu);
  APS_LAUNCHED();
  dim3 g((unsigned)aps_ceil_div((int64_t)D * D, 256), (unsigned)n);
  k_pca_cov<<<g, 256, 0, s>>>(X, d_img_off, D, mu, C);
  APS_LAUNCHED();
  const size_t smem = ((size_t)D * (D + 1) + D) * sizeof(double) + (size_t)D * sizeof(int);
  APS_CUDA(cudaFuncSetAttribute(k_pca_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pca_jacobi<<<n, D, smem, s>>>(C, Vt, d_img_off, D, P, coeff);
  APS_LAUNCHED();
  return APS_OK;
}

// segs (host): n_seg projection segments; out [sum cnt x P]
int aps_k_pca_project(cudaStream_t s, const float* X, int D, int P, const std::vector<aps_proj_seg>& segs, const float* mu,
                      const float* coeff, float* out) {
  if (segs.empty()) return APS_OK;
  std::vector<ProjSeg> h(segs.size());
  std::vector<int64_t> boff(segs.size() + 1, 0);
  for (size_t i = 0; i < segs.size(); ++i) {
    h[i].src = segs[i].src; h[i].dst = segs[i].dst; h[i].cnt = segs[i].cnt; h[i].basis = segs[i].basis;
    boff[i + 1] = boff[i] + (segs[i].cnt + PR_ROWS - 1) / PR_ROWS;
  }
  if (boff.back() == 0) return APS_OK;
  DevBuf<ProjSeg> d_seg;
  DevBuf<int64_t> d_boff;
  APS_TRY(d_seg.alloc(h.size(), s));
  APS_TRY(d_boff.alloc(boff.size(), s));
  APS_CUDA(cudaMemcpyAsync(d_seg.p, h.data(), h.size() * sizeof(ProjSeg), cudaMemcpyHostToDevice, s));
  APS_CUDA(cudaMemcpyAsync(d_boff.p, boff.data(), boff.size() * 8, cudaMemcpyHostToDevice, s));
  APS_CUDA(cudaStreamSynchronize(s));   // pageable host tables
  const size_t smem = ((size_t)D * P + (size_t)PR_ROWS * D) * sizeof(float);
  APS_CUDA(cudaFuncSetAttribute(k_pca_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pca_project<<<(unsigned)boff.back(), 4 * P, smem, s>>>(X, D, P, d_seg.p, d_boff.p, (int)segs.size(), mu, coeff, out);
  APS_LAUNCHED();   // d_seg / d_boff are released in stream order
  return APS_OK;
}
