// aps_featureMatching_mex.cpp -- batched gateway: the whole featureMatching/ stage in one device round trip.
//   matches = aps_featureMatching_mex('global',   allDescriptors, numImg, k, ratioThr, useBF)
//   matches = aps_featureMatching_mex('pairwise', allDescriptors, numImg, matchThreshold, maxRatio [, method])
//   [cand, IuptriIdx] = aps_featureMatching_mex('partners', matchesAll | countMatrix, m)
// Called by the drop-in matlab/featureMatchingGlobal.m / featureMatchingPairwise.m (same signatures as
// PP/featureMatching/featureMatchingGlobal.m:1 and featureMatchingPairwise.m:1) and by the two-line patch of
// PP/imageMatching/imageMatching.m:75-100 shown in INTEGRATION.md.
#include "aps_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 3 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("apsmatch:args", "first argument must be 'global', 'pairwise' or 'partners'");
  char* m = mxArrayToString(prhs[0]);
  const std::string mode = m ? m : "";
  if (m) mxFree(m);
  if (mode == "partners") {
    const mxArray* in = prhs[1];
    const int n = (int)mxGetM(in);
    if ((int)mxGetN(in) != n) mexErrMsgIdAndTxt("imageMatching:InvalidMatchesAllSize", "matchesAll must be an n-by-n cell array.");
    std::vector<int64_t> counts((size_t)n * n, 0);
    for (size_t c = 0; c < counts.size(); ++c) {
      if (mxIsCell(in)) {
        const mxArray* e = mxGetCell(in, c);
        counts[c] = e ? (int64_t)mxGetM(e) : 0;  // putativeCount = cellfun(@(x) size(x,1), matchesAll), :76
      } else {
        counts[c] = (int64_t)mxGetPr(in)[c];
      }
    }
    plhs[0] = mxCreateLogicalMatrix((mwSize)n, (mwSize)n);
    std::vector<int64_t> lin(counts.size() + 1);
    int64_t np = 0;
    if (aps_select_partners(aps_mex_ctx(), counts.data(), n, (int)mxGetScalar(prhs[2]), (uint8_t*)mxGetData(plhs[0]),
                            lin.data(), &np) != APS_OK)
      aps_mex_fail("apsmatch:args");
    if (nlhs > 1) {
      plhs[1] = mxCreateDoubleMatrix((mwSize)np, 1, mxREAL);
      for (int64_t i = 0; i < np; ++i) mxGetPr(plhs[1])[i] = (double)(lin[i] + 1);  // find(): 1-based linear indices
    }
    return;
  }
  if (!mxIsCell(prhs[1])) mexErrMsgIdAndTxt("apsmatch:args", "allDescriptors must be a cell array");
  const int n = (int)mxGetScalar(prhs[2]);
  std::vector<const void*> ptrs;
  std::vector<int64_t> counts;
  std::vector<std::vector<float>> converted;
  int dtype, D;
  aps_mex_collect(prhs[1], n, ptrs, counts, dtype, D, converted);
  if (dtype < 0) {  // all empty -> cell(numImg)   (featureMatchingGlobal.m:49-52)
    plhs[0] = mxCreateCellMatrix((mwSize)n, (mwSize)n);
    return;
  }
  aps_matchlist* ml = nullptr;
  int rc;
  if (mode == "global") {
    if (nrhs < 5) mexErrMsgIdAndTxt("apsmatch:args", "global: need k and ratio threshold");
    rc = aps_feature_matching_global(aps_mex_ctx(), ptrs.data(), counts.data(), n, D, dtype, APS_COL_MAJOR,
                                     (int)mxGetScalar(prhs[3]), mxGetScalar(prhs[4]),
                                     nrhs > 5 ? (int)mxGetScalar(prhs[5]) : 0, &ml);
  } else if (mode == "pairwise") {
    if (nrhs < 5) mexErrMsgIdAndTxt("apsmatch:args", "pairwise: need MatchThreshold and MaxRatio");
    const int method = nrhs > 5 ? (int)mxGetScalar(prhs[5]) : APS_METHOD_EXHAUSTIVE;   // aps_method
    if (method == APS_METHOD_EXHAUSTIVE || dtype == APS_U8) {
      rc = aps_feature_matching_pairwise(aps_mex_ctx(), ptrs.data(), counts.data(), n, D, dtype, APS_COL_MAJOR,
                                         mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), &ml);
    } else {   // 'subsetpdist2' / 'kdtree' (matchFeaturesScratch.m:142-155): staged plan with the Euclidean metric
      aps_pplan* plan = nullptr;
      rc = aps_pplan_create(aps_mex_ctx(), counts.data(), n, D, dtype, &plan);
      if (rc == APS_OK) rc = aps_pplan_set_method(plan, method, 12000, 0);
      if (rc == APS_OK) rc = aps_pplan_upload(plan, ptrs.data(), APS_COL_MAJOR);
      if (rc == APS_OK) rc = aps_pplan_prepare(plan);
      if (rc == APS_OK) rc = aps_pplan_match(plan, mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), 0, 1, &ml);
      aps_pplan_destroy(plan);
    }
  } else {
    mexErrMsgIdAndTxt("apsmatch:args", "unknown mode '%s'", mode.c_str());
    return;
  }
  if (rc != APS_OK) aps_mex_fail("apsmatch:args");
  plhs[0] = aps_mex_cells(ml, n, mode == "pairwise");
  aps_matchlist_free(ml);
}
