/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 * Plain-C entry point around the reference's own MEX gateway, which is compiled VERBATIM from
 * /root/reference/Procedural Program/mex/nearest2HammingExhaustive{,OMP}MEX.cpp (never copied
 * into this repo) against oracle/mexshim/mex.h.  One shared object per reference file, because
 * each defines mexFunction.  Inputs/outputs use the MEX boundary's COLUMN-major layout.
 */
#include "mex.h"

extern "C" int ref_nearest2_hamming(const uint8_t* A_cm, int64_t N1, const uint8_t* B_cm, int64_t N2, int nb,
                                    uint32_t* idx2, float* d1, float* d2, char* err, int errlen) {
  mxArray* A = mxCreateNumericMatrix((mwSize)N1, (mwSize)nb, mxUINT8_CLASS, mxREAL);
  mxArray* B = mxCreateNumericMatrix((mwSize)N2, (mwSize)nb, mxUINT8_CLASS, mxREAL);
  if (N1 * nb) std::memcpy(mxGetData(A), A_cm, (size_t)N1 * nb);
  if (N2 * nb) std::memcpy(mxGetData(B), B_cm, (size_t)N2 * nb);
  mxArray* out[3] = {nullptr, nullptr, nullptr};
  const mxArray* in[2] = {A, B};
  int rc = 0;
  try {
    mexFunction(3, out, 2, in);
    if (N1) {
      std::memcpy(idx2, mxGetData(out[0]), (size_t)N1 * 4);
      std::memcpy(d1, mxGetData(out[1]), (size_t)N1 * 4);
      std::memcpy(d2, mxGetData(out[2]), (size_t)N1 * 4);
    }
  } catch (const mexShimError& e) {
    std::snprintf(err, (size_t)errlen, "%s: %s", e.id.c_str(), e.what());
    rc = 1;
  }
  for (mxArray* o : out) mxDestroyArray(o);
  mxDestroyArray(A);
  mxDestroyArray(B);
  return rc;
}
