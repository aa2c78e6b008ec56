// aps_matchFeatures_mex.cpp -- gateway of the per-pair matcher behind matlab/matchFeaturesScratch.m.
//   [matches, matchMetric] = aps_matchFeatures_mex(A, B, kind, matchThreshold, maxRatio, unique [, method])
//     method (float only, aps_method of include/apsmatch.h): 0 exhaustive [default], 1 'subsetpdist2', 2 'kdtree'
//     (Euclidean searches of matchFeaturesScratch.m:142-155; Unique = true only)
//     kind 0 : float descriptors  [N x D]     single / double   (nearest2SSDExhaustive + filters,
//                                                                 PP/featureMatching/matchFeaturesScratch.m:105-126,169-215,322-366)
//     kind 1 : packed binary      [N x nb]    uint8             (binaryFeatures.Features; nearest2HammingExhaustiveMEX, :295-319)
//     kind 2 : unpacked 0/1 bits  [N x Dbits] logical / uint8   (packBits :617-646 runs on the device)
//   matches  [K x 2] uint32, 1-based rows of A / B   (:205,208: [double, uint32] concatenates to uint32)
//   matchMetric [K x 1]: SSD as double (d1 lives in inf(N1,1), :344) or percent Hamming as single (:120)
// The decision "normalise iff max|A| > 2 or max|B| > 2" (:105-110), the ratio / threshold tests and the greedy
// Unique pass all run inside aps_match_features; the .m only parses name-value pairs and picks `kind`.
#include "aps_mex_common.h"

static void empty_result(int nlhs, mxArray* plhs[]) {  // :84-88
  plhs[0] = mxCreateNumericMatrix(0, 2, mxUINT32_CLASS, mxREAL);
  if (nlhs > 1) plhs[1] = mxCreateNumericMatrix(0, 1, mxSINGLE_CLASS, mxREAL);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 6 && nrhs != 7)
    mexErrMsgIdAndTxt("apsmatch:args", "usage: [matches, metric] = aps_matchFeatures_mex(A, B, kind, matchThreshold, maxRatio, unique [, method])");
  const int method = nrhs > 6 ? (int)mxGetScalar(prhs[6]) : APS_METHOD_EXHAUSTIVE;
  if (nlhs > 2) mexErrMsgIdAndTxt("apsmatch:args", "at most two outputs");
  const mxArray *A = prhs[0], *B = prhs[1];
  const int kind = (int)mxGetScalar(prhs[2]);
  const double thr = mxGetScalar(prhs[3]), ratio = mxGetScalar(prhs[4]);
  const int unique = mxIsLogicalScalarTrue(prhs[5]) || (!mxIsLogical(prhs[5]) && mxGetScalar(prhs[5]) != 0.0);
  if (kind < 0 || kind > 2) mexErrMsgIdAndTxt("apsmatch:args", "kind must be 0 (float), 1 (packed bytes) or 2 (unpacked bits)");
  if (!(ratio > 0.0 && ratio <= 1.0) || !(thr >= 0.0))  // inputParser validators, :60-62
    mexErrMsgIdAndTxt("apsmatch:args", "MaxRatio must be in (0,1] and MatchThreshold >= 0");
  if (mxGetNumberOfDimensions(A) != 2 || mxGetNumberOfDimensions(B) != 2 || mxIsComplex(A) || mxIsComplex(B))
    mexErrMsgIdAndTxt("apsmatch:type", "descriptors must be real 2-D matrices");
  const int64_t N1 = (int64_t)mxGetM(A), N2 = (int64_t)mxGetM(B);
  const int D = (int)mxGetN(A);
  const bool bytes_ok = (mxIsUint8(A) || mxIsLogical(A)) && (mxIsUint8(B) || mxIsLogical(B));
  if (kind != 0) {
    if (!bytes_ok) mexErrMsgIdAndTxt("apsmatch:type", "binary descriptors must be uint8 or logical");
    if (mxIsEmpty(A) || mxIsEmpty(B)) { empty_result(nlhs, plhs); return; }
    if ((int)mxGetN(B) != D) mexErrMsgIdAndTxt("hamm2nn:cols", "A and B must have same number of columns (bytes).");
  } else {
    if (!((mxIsSingle(A) || mxIsDouble(A)) && (mxIsSingle(B) || mxIsDouble(B))))
      mexErrMsgIdAndTxt("apsmatch:type", "float descriptors must be single or double");
    if (mxIsEmpty(A) || mxIsEmpty(B))  // validateattributes(...,'nonempty'), :281-282
      mexErrMsgIdAndTxt("MATLAB:expectedNonempty", "Expected input to be nonempty.");
    if ((int)mxGetN(B) != D) mexErrMsgIdAndTxt("apsmatch:dim", "Descriptor dimensions must match for non-binary.");
  }
  // double inputs are computed in single precision: :107-108 does the same cast whenever it normalises; for small-magnitude
  // double inputs the reference would stay in double (the pipeline's descriptors are single, getFeaturePoints.m)
  std::vector<float> ca, cb;
  const void *pa = mxGetData(A), *pb = mxGetData(B);
  if (kind == 0 && mxIsDouble(A)) { const double* s = mxGetPr(A); ca.assign(s, s + mxGetNumberOfElements(A)); pa = ca.data(); }
  if (kind == 0 && mxIsDouble(B)) { const double* s = mxGetPr(B); cb.assign(s, s + mxGetNumberOfElements(B)); pb = cb.data(); }
  std::vector<uint32_t> rows((size_t)(N1 > 0 ? N1 : 1) * 2);
  std::vector<double> met((size_t)(N1 > 0 ? N1 : 1));
  int64_t K = 0;
  int rc;
  if (kind == 0 && method != APS_METHOD_EXHAUSTIVE) {   // two-image staged plan with the Euclidean metric
    if (!unique) mexErrMsgIdAndTxt("apsmatch:method", "approximate float matching is built for Unique = true only");
    const void* desc[2] = {pa, pb};
    const int64_t counts[2] = {N1, N2};
    aps_pplan* plan = nullptr;
    aps_matchlist* ml = nullptr;
    rc = aps_pplan_create(aps_mex_ctx(), counts, 2, D, APS_F32, &plan);
    if (rc == APS_OK) rc = aps_pplan_set_method(plan, method, 12000, 0);
    if (rc == APS_OK) rc = aps_pplan_upload(plan, desc, APS_COL_MAJOR);
    if (rc == APS_OK) rc = aps_pplan_prepare(plan);
    if (rc == APS_OK) rc = aps_pplan_match(plan, thr, ratio, 0, 1, &ml);
    aps_pplan_destroy(plan);
    if (rc == APS_OK) {
      K = aps_matchlist_total(ml);
      const uint32_t* r = aps_matchlist_rows(ml);
      const double* m = aps_matchlist_metric(ml);
      for (int64_t i = 0; i < K; ++i) { rows[2 * i] = r[2 * i]; rows[2 * i + 1] = r[2 * i + 1]; met[i] = m[i]; }
      aps_matchlist_free(ml);
    }
  } else if (kind == 2)
    rc = aps_match_features_bits(aps_mex_ctx(), (const uint8_t*)pa, N1, (const uint8_t*)pb, N2, D, APS_COL_MAJOR, thr, ratio,
                                 unique, rows.data(), met.data(), &K);
  else
    rc = aps_match_features(aps_mex_ctx(), pa, N1, pb, N2, D, kind == 0 ? APS_F32 : APS_U8, APS_COL_MAJOR, thr, ratio, unique,
                            rows.data(), met.data(), &K);
  if (rc != APS_OK) aps_mex_fail("apsmatch:args");
  plhs[0] = mxCreateNumericMatrix((mwSize)K, 2, mxUINT32_CLASS, mxREAL);
  uint32_t* o = (uint32_t*)mxGetData(plhs[0]);
  for (int64_t r = 0; r < K; ++r) {  // library rows are interleaved (col1, col2); MATLAB wants column-major [K x 2]
    o[r] = rows[2 * r];
    o[r + K] = rows[2 * r + 1];
  }
  if (nlhs > 1) {
    if (kind == 0) {
      plhs[1] = mxCreateDoubleMatrix((mwSize)K, 1, mxREAL);
      for (int64_t r = 0; r < K; ++r) mxGetPr(plhs[1])[r] = met[r];
    } else {
      plhs[1] = mxCreateNumericMatrix((mwSize)K, 1, mxSINGLE_CLASS, mxREAL);
      for (int64_t r = 0; r < K; ++r) ((float*)mxGetData(plhs[1]))[r] = (float)met[r];
    }
  }
}
