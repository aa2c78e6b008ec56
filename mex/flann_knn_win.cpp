// flann_knn_win.cpp -- drop-in MEX gateway with the reference's name and signatures
//   (A) [idx, dist] = flann_knn_win(train, k [, method, trees, checks])
//   (B) [idx, dist] = flann_knn_win(train, query, k [, method, trees, checks])
// replacing PP/mex/flann_knn.cpp:118-253 (OpenCV FLANN / BFMatcher) by aps_flann_knn (B200, exact search).
// Build with MATLAB:  mex -O flann_knn_win.cpp -I../include -L<pkg dir> -lapsmatch
#include "aps_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs;
  if (nrhs < 2)
    mexErrMsgIdAndTxt("flann_knn:args", "Usage: [idx, dist] = flann_knn(train, k [, method, trees, checks])\n"
                                        "   or: [idx, dist] = flann_knn(train, query, k [, method, trees, checks])");
  const mxArray* train = prhs[0];
  if (mxIsComplex(train) || mxGetNumberOfDimensions(train) != 2) mexErrMsgIdAndTxt("flann_knn:type", "train must be real 2D");
  const bool tf = mxIsSingle(train), tb = mxIsUint8(train);
  if (!tf && !tb) mexErrMsgIdAndTxt("flann_knn:type", "Descriptors must be single (float) or uint8 (binary)");
  const mxArray* query = train;
  int kArg = 1;
  if (nrhs >= 3 && (mxIsSingle(prhs[1]) || mxIsUint8(prhs[1])) && ((tf && mxIsSingle(prhs[1])) || (tb && mxIsUint8(prhs[1])))) {
    if (mxIsComplex(prhs[1]) || mxGetNumberOfDimensions(prhs[1]) != 2) mexErrMsgIdAndTxt("flann_knn:type", "query must be real 2D");
    query = prhs[1];
    kArg = 2;
  }
  if (!mxIsDouble(prhs[kArg]) || mxGetNumberOfElements(prhs[kArg]) != 1) mexErrMsgIdAndTxt("flann_knn:type", "k must be a scalar double");
  const int k = (int)mxGetScalar(prhs[kArg]);
  if (k <= 0) mexErrMsgIdAndTxt("flann_knn:k", "k must be > 0");
  std::string method = "flann";
  if (nrhs >= kArg + 2) {
    char* s = mxArrayToString(prhs[kArg + 1]);
    method = s ? s : "";
    if (s) mxFree(s);
  }
  const int trees = (nrhs >= kArg + 3) ? (int)mxGetScalar(prhs[kArg + 2]) : 4;
  const int checks = (nrhs >= kArg + 4) ? (int)mxGetScalar(prhs[kArg + 3]) : 32;
  const mwSize Ft = mxGetM(train), D = mxGetN(train), Fq = mxGetM(query);
  if (mxGetN(query) != D) mexErrMsgIdAndTxt("flann_knn:dim", "query must have same descriptor dimension as train");
  plhs[0] = mxCreateNumericMatrix(Fq, (mwSize)k, mxUINT32_CLASS, mxREAL);
  plhs[1] = mxCreateNumericMatrix(Fq, (mwSize)k, mxSINGLE_CLASS, mxREAL);
  const int rc = aps_flann_knn(aps_mex_ctx(), mxGetData(train), (int64_t)Ft, mxGetData(query), (int64_t)Fq, (int)D,
                               tf ? APS_F32 : APS_U8, APS_COL_MAJOR, k, method.c_str(), trees, checks,
                               (uint32_t*)mxGetData(plhs[0]), (float*)mxGetData(plhs[1]));
  if (rc != APS_OK) aps_mex_fail("flann_knn:args");
}
