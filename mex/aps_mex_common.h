// aps_mex_common.h -- shared glue of the MEX gateways: one lazily created GPU context per MATLAB
// process (or parfor worker), released at mexAtExit, and status -> mexErrMsgIdAndTxt mapping that keeps
// the reference's error identifiers (PP/mex/flann_knn.cpp:94-95,126-177,202; nearest2Hamming...MEX.cpp:17-28).
// The gateways only marshal mxArrays into the C ABI of include/apsmatch.h: column-major pointers go
// through unchanged (APS_COL_MAJOR), so the host-side transposes of flann_knn.cpp:99-116 disappear.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../include/apsmatch.h"
#include "mex.h"

static aps_ctx* g_aps_ctx = nullptr;
static void aps_mex_cleanup(void) {
  if (g_aps_ctx) aps_ctx_destroy(g_aps_ctx);
  g_aps_ctx = nullptr;
}
static inline void aps_mex_fail(const char* fallback_id) {
  const char* id = aps_error_id();
  mexErrMsgIdAndTxt((id && id[0]) ? id : fallback_id, "%s", aps_last_error());
}
static inline aps_ctx* aps_mex_ctx(void) {
  if (!g_aps_ctx) {
    if (aps_ctx_create(0, &g_aps_ctx) != APS_OK) aps_mex_fail("apsmatch:nogpu");  // no CPU fallback
    mexAtExit(aps_mex_cleanup);
  }
  return g_aps_ctx;
}
// allDescriptors{i}: numeric matrix, or a binaryFeatures object (its .Features, featureMatchingGlobal.m:56-75)
static inline const mxArray* aps_mex_features(const mxArray* cell_elem) {
  if (!cell_elem) return nullptr;
  if (mxIsClass(cell_elem, "binaryFeatures")) return mxGetProperty(cell_elem, 0, "Features");
  return cell_elem;
}
// Collects the cell array into pointer / count vectors; returns dtype (APS_F32 / APS_U8) and D.
static inline void aps_mex_collect(const mxArray* cellArr, int n, std::vector<const void*>& ptrs,
                                   std::vector<int64_t>& counts, int& dtype, int& D,
                                   std::vector<std::vector<float>>& converted) {
  ptrs.assign(n, nullptr);
  counts.assign(n, 0);
  dtype = -1;
  D = 0;
  for (int i = 0; i < n; ++i) {
    const mxArray* f = (mwSize)i < mxGetNumberOfElements(cellArr) ? aps_mex_features(mxGetCell(cellArr, i)) : nullptr;
    if (!f || mxIsEmpty(f)) continue;
    if (mxGetNumberOfDimensions(f) != 2 || mxIsComplex(f))
      mexErrMsgIdAndTxt("flann_knn:type", "descriptors must be real 2D matrices");
    int dt;
    if (mxIsUint8(f)) dt = APS_U8;
    else if (mxIsSingle(f) || mxIsDouble(f)) dt = APS_F32;
    else mexErrMsgIdAndTxt("flann_knn:type", "Descriptors must be single (float) or uint8 (binary)");
    if (dtype < 0) { dtype = dt; D = (int)mxGetN(f); }
    if (dt != dtype) mexErrMsgIdAndTxt("flann_knn:type", "all descriptor matrices must have the same class");
    if ((int)mxGetN(f) != D) mexErrMsgIdAndTxt("flann_knn:dim", "all descriptor matrices must have the same width");
    counts[i] = (int64_t)mxGetM(f);
    if (mxIsDouble(f)) {  // single(allDesc), featureMatchingGlobal.m:81
      const double* s = mxGetPr(f);
      converted.emplace_back(s, s + mxGetNumberOfElements(f));
      ptrs[i] = converted.back().data();
    } else {
      ptrs[i] = mxGetData(f);
    }
  }
}
// CSR match list -> n x n cell of [M x 2] double (featureMatchingGlobal.m:155-159; untouched cells stay [])
static inline mxArray* aps_mex_cells(const aps_matchlist* ml, int n, bool fill_upper_with_0x2) {
  mxArray* out = mxCreateCellMatrix((mwSize)n, (mwSize)n);
  const int64_t* pp = aps_matchlist_pair_ptr(ml);
  const uint32_t* rows = aps_matchlist_rows(ml);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) {
      const int64_t c = i + (int64_t)j * n, a = pp[c], b = pp[c + 1];
      if (b == a && !fill_upper_with_0x2) continue;
      mxArray* m = mxCreateDoubleMatrix((mwSize)(b - a), 2, mxREAL);
      double* d = mxGetPr(m);
      for (int64_t r = a; r < b; ++r) {
        d[r - a] = (double)rows[2 * r];
        d[(r - a) + (b - a)] = (double)rows[2 * r + 1];
      }
      mxSetCell(out, (mwIndex)c, m);
    }
  return out;
}
