// nearest2HammingExhaustiveMEX.cpp -- drop-in MEX gateway: [idx2, d1, d2] = nearest2HammingExhaustiveMEX(A, B)
// replacing PP/mex/nearest2HammingExhaustiveMEX.cpp:16-80 by aps_nearest2_hamming (B200).
#include "aps_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs;
  if (nrhs != 2) mexErrMsgIdAndTxt("hamm2nn:nrhs", "Need Abytes,Bbytes");
  const mxArray *A = prhs[0], *B = prhs[1];
  if (!mxIsUint8(A) || !mxIsUint8(B)) mexErrMsgIdAndTxt("hamm2nn:type", "Inputs must be uint8.");
  if (mxGetNumberOfDimensions(A) != 2 || mxGetNumberOfDimensions(B) != 2) mexErrMsgIdAndTxt("hamm2nn:dim", "2D only.");
  const mwSize N1 = mxGetM(A), nb = mxGetN(A), N2 = mxGetM(B);
  if (mxGetN(B) != nb) mexErrMsgIdAndTxt("hamm2nn:cols", "Byte width mismatch.");
  plhs[0] = mxCreateNumericMatrix(N1, 1, mxUINT32_CLASS, mxREAL);
  plhs[1] = mxCreateNumericMatrix(N1, 1, mxSINGLE_CLASS, mxREAL);
  plhs[2] = mxCreateNumericMatrix(N1, 1, mxSINGLE_CLASS, mxREAL);
  const int rc = aps_nearest2_hamming(aps_mex_ctx(), (const uint8_t*)mxGetData(A), (int64_t)N1, (const uint8_t*)mxGetData(B),
                                      (int64_t)N2, (int)nb, APS_COL_MAJOR, (uint32_t*)mxGetData(plhs[0]),
                                      (float*)mxGetData(plhs[1]), (float*)mxGetData(plhs[2]));
  if (rc != APS_OK) aps_mex_fail("hamm2nn:nrhs");
}
