/* aps_oracle_ransac.c -- TEST INFRASTRUCTURE ONLY (same rules as aps_oracle.c: only tests/, smoke() and
 * bench.py's CPU legs may load it; the product never does).
 *
 * CPU restatement (plain C, double precision, -ffp-contract=off, sequential sums) of the consumer that
 * follows the feature-matching path -- SURVEY.md section 8(f) rank 1:
 *   PP/imageMatching/imageMatching.m:121-156            candidate-pair loop, acceptance ni > 8 + 0.3 nf, inv(model)
 *   PP/imageMatching/imageMatching.m:224-246            refineMatch: matchedPts, estimateTransformationRANSAC(P2, P1)
 *   PP/imageMatching/estimateTransformationRANSAC.m     (projective model only = PP/inputs.m:73)
 *       :94-183  main loop + refit,  :188-225 estimateHomography,  :444-516 findInliers,
 *       :518-530 checkModel,  :532-572 isDegenerate,  :574-596 normalizePoints
 *
 * PARITY UNPINNED against MATLAB: randperm / svd / rcond / mldivide are closed-source or LAPACK calls and
 * the reference ships no tests.  What IS pinned: tests/test_oracle_ransac.py checks this file against an
 * independent numpy restatement that uses LAPACK (numpy.linalg.svd / solve / cond), the library MATLAB
 * itself calls.  Documented choices:
 *   - random minimal samples are an INPUT (table of draws); draw d is consumed by loop iteration d, valid or
 *     skipped.  The loop also stops when the table is exhausted (the reference would keep drawing up to
 *     10*maxTrials skipped samples; identical unless more than n_draws - maxTrials samples are invalid);
 *   - V(:,end) of svd(A) is computed as the eigenvector of the smallest eigenvalue of A'A (cyclic Jacobi);
 *   - singular values of the centred n x 2 inlier matrix (isDegenerate) come from its 2 x 2 Gram matrix;
 *   - sums over a pair's correspondences (inlier errors, centroid, scatter matrix in findInliers / isDegenerate)
 *     use 32 interleaved partial accumulators (point r -> accumulator r mod 32) combined by a fixed butterfly
 *     (offsets 16, 8, 4, 2, 1); MATLAB's own sum order is unspecified (multithreaded for long vectors);
 *   - rcond(H) is the exact 1-norm reciprocal condition number (MATLAB: LAPACK's estimate of it);
 *   - H \ x uses a 3 x 3 LU factorisation with partial pivoting, T2 \ X back-substitution (T2 is triangular).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS 2.220446049250313e-16

/* ---- cyclic Jacobi eigen-decomposition of a symmetric 9 x 9 matrix (upper triangle is used) -------------
 * Sweep order (p,q) = (0,1),(0,2),...,(7,8); rotation angle from theta = (aqq-app)/(2 apq),
 * t = sgn(theta)/(|theta|+sqrt(theta^2+1)); small-element rule of the classical algorithm
 * (after 3 sweeps an element that cannot change either diagonal entry is set to zero);
 * stops when the off-diagonal absolute sum is exactly zero or after 50 sweeps. */
static void jacobi9(double a[9][9], double v[9][9], double d[9]) {
  double b[9], z[9];
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    b[i] = d[i] = a[i][i];
    z[i] = 0.0;
  }
  for (int sweep = 0; sweep < 50; ++sweep) {
    double sm = 0.0;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) sm += fabs(a[p][q]);
    if (sm == 0.0) return;
    const double tresh = (sweep < 3) ? 0.2 * sm / 81.0 : 0.0;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        const double g = 100.0 * fabs(a[p][q]);
        if (sweep > 3 && fabs(d[p]) + g == fabs(d[p]) && fabs(d[q]) + g == fabs(d[q])) {
          a[p][q] = 0.0;
        } else if (fabs(a[p][q]) > tresh) {
          const double h = d[q] - d[p];
          double t;
          if (fabs(h) + g == fabs(h)) {
            t = a[p][q] / h;
          } else {
            const double theta = 0.5 * h / a[p][q];
            t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          const double c = 1.0 / sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c), hh = t * a[p][q];
          z[p] -= hh; z[q] += hh; d[p] -= hh; d[q] += hh;
          a[p][q] = 0.0;
#define ORC_ROT(M, i, j, k, l) { const double g_ = M[i][j], h_ = M[k][l]; M[i][j] = g_ - s * (h_ + g_ * tau); M[k][l] = h_ + s * (g_ - h_ * tau); }
          for (int j = 0; j < p; ++j) ORC_ROT(a, j, p, j, q)
          for (int j = p + 1; j < q; ++j) ORC_ROT(a, p, j, j, q)
          for (int j = q + 1; j < 9; ++j) ORC_ROT(a, p, j, q, j)
          for (int j = 0; j < 9; ++j) ORC_ROT(v, j, p, j, q)
#undef ORC_ROT
        }
      }
    for (int i = 0; i < 9; ++i) {
      b[i] += z[i];
      d[i] = b[i];
      z[i] = 0.0;
    }
  }
}

/* normalizePoints (estimateTransformationRANSAC.m:574-596) over the points sel[0..n) (sel == NULL: 0..n-1).
 * T = [s 0 -s*cx; 0 s -s*cy; 0 0 1]; returns s, -s*cx, -s*cy */
static void normalize_T(const double* p, const uint32_t* sel, int64_t n, double T[3]) {
  double sx = 0.0, sy = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = sel ? sel[i] : i;
    sx += p[2 * r];
    sy += p[2 * r + 1];
  }
  const double cx = sx / (double)n, cy = sy / (double)n;
  double sd = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = sel ? sel[i] : i;
    const double dx = p[2 * r] - cx, dy = p[2 * r + 1] - cy;
    sd += sqrt(dx * dx + dy * dy);
  }
  const double scale = 1.0 / (sd / (double)n);
  T[0] = scale;
  T[1] = -scale * cx;
  T[2] = -scale * cy;
}

/* estimateHomography (:188-225): normalised DLT on the selected correspondences; H row-major 3 x 3 */
static void estimate_homography(const double* p1, const double* p2, const uint32_t* sel, int64_t n, double H[9]) {
  double T1[3], T2[3];
  normalize_T(p1, sel, n, T1);
  normalize_T(p2, sel, n, T2);
  double M[9][9];
  memset(M, 0, sizeof M);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = sel ? sel[i] : i;
    const double x = T1[0] * p1[2 * r] + T1[1], y = T1[0] * p1[2 * r + 1] + T1[2];
    const double u = T2[0] * p2[2 * r] + T2[1], v = T2[0] * p2[2 * r + 1] + T2[2];
    const double r1[9] = {-x, -y, -1.0, 0.0, 0.0, 0.0, x * u, y * u, u};
    const double r2[9] = {0.0, 0.0, 0.0, -x, -y, -1.0, x * v, y * v, v};
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b) M[a][b] += r1[a] * r1[b] + r2[a] * r2[b];
  }
  double V[9][9], d[9];
  jacobi9(M, V, d);
  int im = 0;
  for (int i = 1; i < 9; ++i)
    if (d[i] < d[im]) im = i;
  double Hn[9];
  for (int i = 0; i < 9; ++i) Hn[i] = V[i][im] / V[8][im]; /* H_norm / H_norm(3,3) */
  /* X = T2 \ Hn  (T2 upper triangular: rows 1,2 = (Hn_row - t*Hn_row3)/s) */
  double X[9];
  for (int j = 0; j < 3; ++j) {
    X[6 + j] = Hn[6 + j];
    X[j] = (Hn[j] - T2[1] * Hn[6 + j]) / T2[0];
    X[3 + j] = (Hn[3 + j] - T2[2] * Hn[6 + j]) / T2[0];
  }
  /* H = X * T1 */
  for (int i = 0; i < 3; ++i) {
    H[3 * i + 0] = X[3 * i + 0] * T1[0];
    H[3 * i + 1] = X[3 * i + 1] * T1[0];
    H[3 * i + 2] = (X[3 * i + 0] * T1[1] + X[3 * i + 1] * T1[2]) + X[3 * i + 2];
  }
}

static double det3(const double H[9]) {
  return (H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6])) + H[2] * (H[3] * H[7] - H[4] * H[6]);
}
static void adj3(const double H[9], double A[9]) {
  A[0] = H[4] * H[8] - H[5] * H[7]; A[1] = H[2] * H[7] - H[1] * H[8]; A[2] = H[1] * H[5] - H[2] * H[4];
  A[3] = H[5] * H[6] - H[3] * H[8]; A[4] = H[0] * H[8] - H[2] * H[6]; A[5] = H[2] * H[3] - H[0] * H[5];
  A[6] = H[3] * H[7] - H[4] * H[6]; A[7] = H[1] * H[6] - H[0] * H[7]; A[8] = H[0] * H[4] - H[1] * H[3];
}
static double norm1_3(const double H[9]) {
  double m = 0.0;
  for (int j = 0; j < 3; ++j) {
    const double c = (fabs(H[j]) + fabs(H[3 + j])) + fabs(H[6 + j]);
    if (c > m) m = c;
  }
  return m;
}
/* checkModel (:518-530) */
static int check_model(const double H[9]) {
  for (int i = 0; i < 9; ++i)
    if (!isfinite(H[i])) return 0;
  const double det = det3(H);
  if (!(fabs(det) > ORC_EPS)) return 0;
  double A[9];
  adj3(H, A);
  for (int i = 0; i < 9; ++i) A[i] = A[i] / det;
  const double rc = 1.0 / (norm1_3(H) * norm1_3(A));
  return rc > ORC_EPS;
}

typedef struct { double L10, L20, L21, U[6]; int piv[3]; } lu3_t; /* U = u00 u01 u02 u11 u12 u22 */
static void lu3(const double H[9], lu3_t* f) {
  double a[3][3] = {{H[0], H[1], H[2]}, {H[3], H[4], H[5]}, {H[6], H[7], H[8]}};
  int pv[3] = {0, 1, 2};
  int m = 0;
  if (fabs(a[1][0]) > fabs(a[m][0])) m = 1;
  if (fabs(a[2][0]) > fabs(a[m][0])) m = 2;
  if (m != 0) { for (int j = 0; j < 3; ++j) { double t = a[0][j]; a[0][j] = a[m][j]; a[m][j] = t; } int t = pv[0]; pv[0] = pv[m]; pv[m] = t; }
  a[1][0] = a[1][0] / a[0][0];
  a[2][0] = a[2][0] / a[0][0];
  a[1][1] -= a[1][0] * a[0][1]; a[1][2] -= a[1][0] * a[0][2];
  a[2][1] -= a[2][0] * a[0][1]; a[2][2] -= a[2][0] * a[0][2];
  if (fabs(a[2][1]) > fabs(a[1][1])) { for (int j = 0; j < 3; ++j) { double t = a[1][j]; a[1][j] = a[2][j]; a[2][j] = t; } int t = pv[1]; pv[1] = pv[2]; pv[2] = t; }
  a[2][1] = a[2][1] / a[1][1];
  a[2][2] -= a[2][1] * a[1][2];
  f->L10 = a[1][0]; f->L20 = a[2][0]; f->L21 = a[2][1];
  f->U[0] = a[0][0]; f->U[1] = a[0][1]; f->U[2] = a[0][2]; f->U[3] = a[1][1]; f->U[4] = a[1][2]; f->U[5] = a[2][2];
  f->piv[0] = pv[0]; f->piv[1] = pv[1]; f->piv[2] = pv[2];
}
static void lu3_solve(const lu3_t* f, const double b[3], double x[3]) {
  const double y0 = b[f->piv[0]];
  const double y1 = b[f->piv[1]] - f->L10 * y0;
  const double y2 = (b[f->piv[2]] - f->L20 * y0) - f->L21 * y1;
  x[2] = y2 / f->U[5];
  x[1] = (y1 - f->U[4] * x[2]) / f->U[3];
  x[0] = ((y0 - f->U[1] * x[1]) - f->U[2] * x[2]) / f->U[0];
}

/* symmetric transfer error of correspondence r under H (:466-471) */
static double point_error(const double H[9], const lu3_t* f, const double* p1, const double* p2, int64_t r) {
  const double x = p1[2 * r], y = p1[2 * r + 1], u = p2[2 * r], v = p2[2 * r + 1];
  const double t2 = (H[6] * x + H[7] * y) + H[8];
  const double tx = ((H[0] * x + H[1] * y) + H[2]) / t2, ty = ((H[3] * x + H[4] * y) + H[5]) / t2;
  const double b[3] = {u, v, 1.0};
  double w[3];
  lu3_solve(f, b, w);
  const double ix = w[0] / w[2], iy = w[1] / w[2];
  const double d1 = (u - tx) * (u - tx) + (v - ty) * (v - ty);
  const double d2 = (x - ix) * (x - ix) + (y - iy) * (y - iy);
  double e = sqrt(d1 + d2);
  if (!isfinite(e)) e = INFINITY;
  if (!isfinite(t2 / t2)) e = INFINITY; /* errors(abs(transformed(:,3)) < eps) = inf : column 3 is t2/t2 */
  return e;
}

/* total of 32 interleaved partial sums, butterfly order */
static double lane_total(double a[32]) {
  double t[32];
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; ++l) t[l] = a[l] + a[l ^ o];
    memcpy(a, t, sizeof t);
  }
  return a[0];
}

/* findInliers (:444-516) incl. isDegenerate (:532-572) on the inliers' source points.
 * returns the inlier count (0 when degenerate) and their error sum */
static int64_t find_inliers(const double H[9], const double* p1, const double* p2, int64_t n, double thr,
                            uint8_t* mask, double* err_sum) {
  lu3_t f;
  lu3(H, &f);
  int64_t cnt = 0;
  double esl[32] = {0}, sxl[32] = {0}, syl[32] = {0};
  for (int64_t r = 0; r < n; ++r) {
    const double e = point_error(H, &f, p1, p2, r);
    const int in = e < thr;
    mask[r] = (uint8_t)in;
    if (in) {
      ++cnt;
      esl[r & 31] += e;
      sxl[r & 31] += p1[2 * r];
      syl[r & 31] += p1[2 * r + 1];
    }
  }
  double es = lane_total(esl);
  const double sx = lane_total(sxl), sy = lane_total(syl);
  if (cnt >= 4) {
    const double mx = sx / (double)cnt, my = sy / (double)cnt;
    double sxxl[32] = {0}, sxyl[32] = {0}, syyl[32] = {0};
    for (int64_t r = 0; r < n; ++r)
      if (mask[r]) {
        const double dx = p1[2 * r] - mx, dy = p1[2 * r + 1] - my;
        sxxl[r & 31] += dx * dx;
        sxyl[r & 31] += dx * dy;
        syyl[r & 31] += dy * dy;
      }
    const double sxx = lane_total(sxxl), sxy = lane_total(sxyl), syy = lane_total(syyl);
    /* singular values^2 of the centred matrix = eigenvalues of [sxx sxy; sxy syy] */
    const double hd = 0.5 * (sxx - syy);
    const double l1 = 0.5 * (sxx + syy) + sqrt(hd * hd + sxy * sxy);
    const double l2 = (sxx * syy - sxy * sxy) / l1;
    const double ratio = sqrt(fmax(l2, 0.0)) / sqrt(l1);
    if (ratio < 1e-3) { /* NaN (all inliers identical) compares false, as in MATLAB */
      memset(mask, 0, (size_t)n);
      cnt = 0;
      es = 0.0;
    }
  }
  *err_sum = es;
  return cnt;
}

/* estimateTransformationRANSAC (:94-183), projective.  p1 = matchedPoints1, p2 = matchedPoints2 (row-major n x 2).
 * samples: n_draws x 4 zero-based row indices.  model: 9 doubles row-major (NaN when not found).
 * Returns isFound; *draws_used = loop iterations executed. */
int orc_ransac_homography(const double* p1, const double* p2, int64_t n, double max_distance, double confidence,
                          int max_trials_in, const uint32_t* samples, int64_t n_draws, double* model,
                          uint8_t* inliers, int32_t* n_inliers, int32_t* draws_used) {
  for (int i = 0; i < 9; ++i) model[i] = NAN;
  memset(inliers, 0, (size_t)(n > 0 ? n : 0));
  *n_inliers = 0;
  *draws_used = 0;
  if (n < 4) return 0;
  uint8_t* cur = (uint8_t*)malloc((size_t)n);
  uint8_t* best = (uint8_t*)calloc((size_t)n, 1);
  double bestH[9], bestErr = INFINITY;
  int64_t bestCnt = 0;
  int haveBest = 0;
  double maxTrials = (double)max_trials_in;
  const int64_t maxSkip = (int64_t)max_trials_in * 10;
  int64_t trial = 1, skip = 0, d = 0;
  while ((double)trial <= maxTrials && skip < maxSkip && d < n_draws) {
    const uint32_t* s = samples + 4 * d;
    ++d;
    double H[9];
    if (s[0] >= n || s[1] >= n || s[2] >= n || s[3] >= n) { /* table entry outside the pair: a skipped sample */
      ++skip;
      continue;
    }
    estimate_homography(p1, p2, s, 4, H);
    if (!check_model(H)) {
      ++skip;
      continue;
    }
    double es;
    const int64_t cnt = find_inliers(H, p1, p2, n, max_distance, cur, &es);
    if (cnt >= 4) {
      const double meanErr = es / (double)cnt;
      if (cnt > bestCnt || (cnt == bestCnt && meanErr < bestErr)) {
        memcpy(best, cur, (size_t)n);
        memcpy(bestH, H, sizeof bestH);
        bestCnt = cnt;
        bestErr = meanErr;
        haveBest = 1;
        const double ratio = (double)cnt / (double)n;
        if (ratio > 0.0) {
          const double r2 = ratio * ratio;
          const double t = ceil(log(1.0 - confidence / 100.0) / log(1.0 - r2 * r2));
          if (t < maxTrials) maxTrials = t; /* min(maxTrials, NaN) keeps maxTrials */
        }
      }
    }
    ++trial;
  }
  *draws_used = (int32_t)d;
  int found = 0;
  if (bestCnt >= 4) {
    /* refit on all inliers (:150-176) */
    uint32_t* sel = (uint32_t*)malloc((size_t)bestCnt * sizeof(uint32_t));
    int64_t m = 0;
    for (int64_t r = 0; r < n; ++r)
      if (best[r]) sel[m++] = (uint32_t)r;
    double H[9];
    estimate_homography(p1, p2, sel, m, H);
    free(sel);
    int use_best = 1;
    if (check_model(H)) {
      double es;
      const int64_t cnt = find_inliers(H, p1, p2, n, max_distance, cur, &es);
      if (cnt >= 4) {
        memcpy(model, H, 9 * sizeof(double));
        memcpy(inliers, cur, (size_t)n);
        *n_inliers = (int32_t)cnt;
        use_best = 0;
      }
    }
    if (use_best) {
      memcpy(model, bestH, 9 * sizeof(double));
      memcpy(inliers, best, (size_t)n);
      *n_inliers = (int32_t)bestCnt;
    }
    found = 1;
  } else if (haveBest) {
    memcpy(model, bestH, 9 * sizeof(double));
  }
  free(cur);
  free(best);
  return found;
}

/* inv(model) for tforms{jj,ii} (imageMatching.m:151): adjugate / determinant */
void orc_inv3(const double* H, double* out) {
  double A[9];
  adj3(H, A);
  const double det = det3(H);
  for (int i = 0; i < 9; ++i) out[i] = A[i] / det;
}

/* imageMatching.m:121-156 over a batch of candidate pairs.  Pair p owns correspondences pt_ptr[p]..pt_ptr[p+1]
 * (p1 = keypoints of image jj, p2 = keypoints of image ii: refineMatch passes (matchedPts_2, matchedPts_1)).
 * accepted[p] = ni > 8 + 0.3 nf (pairs with nf < 4 are skipped = not accepted, model NaN). */
void orc_image_matching_batch(int64_t n_pairs, const int64_t* pt_ptr, const double* p1, const double* p2,
                              double max_distance, double confidence, int max_trials, const uint32_t* samples,
                              int64_t n_draws, double* models, double* models_inv, uint8_t* inliers,
                              int32_t* n_inliers, uint8_t* accepted, int32_t* draws_used) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int64_t o = pt_ptr[p], nf = pt_ptr[p + 1] - o;
    const int found = orc_ransac_homography(p1 + 2 * o, p2 + 2 * o, nf, max_distance, confidence, max_trials,
                                            samples + p * n_draws * 4, n_draws, models + 9 * p, inliers + o,
                                            n_inliers + p, draws_used + p);
    (void)found;
    accepted[p] = (nf >= 4) && ((double)n_inliers[p] > 8.0 + 0.3 * (double)nf);
    if (accepted[p]) {
      orc_inv3(models + 9 * p, models_inv + 9 * p);
    } else {
      for (int i = 0; i < 9; ++i) models_inv[9 * p + i] = NAN;
    }
  }
}
