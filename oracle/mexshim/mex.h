/*
 * mex.h -- minimal stand-in for MATLAB's MEX C API (TEST INFRASTRUCTURE ONLY).
 *
 * MATLAB is not installed in this image, so (a) the reference's own MEX sources under
 * /root/reference/Procedural Program/mex (its .cpp files) and (b) this repo's gateways in mex/ are
 * compiled against this header instead of MathWorks' <mex.h>.  It implements just enough of the
 * published API (column-major mxArray with class id, dims, data; cells; structs; char rows) for
 * those files to compile AND run under a plain C++ driver.  Written from the public MEX API
 * documentation; it is not MathWorks code.  With a real MATLAB the same sources build against the
 * real <mex.h> unchanged.
 */
#ifndef APS_MEXSHIM_MEX_H
#define APS_MEXSHIM_MEX_H

#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef size_t mwSize;
typedef size_t mwIndex;
typedef uint16_t mxChar;

typedef enum {
  mxUNKNOWN_CLASS = 0,
  mxCELL_CLASS,
  mxSTRUCT_CLASS,
  mxLOGICAL_CLASS,
  mxCHAR_CLASS,
  mxVOID_CLASS,
  mxDOUBLE_CLASS,
  mxSINGLE_CLASS,
  mxINT8_CLASS,
  mxUINT8_CLASS,
  mxINT16_CLASS,
  mxUINT16_CLASS,
  mxINT32_CLASS,
  mxUINT32_CLASS,
  mxINT64_CLASS,
  mxUINT64_CLASS
} mxClassID;

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

struct mxArray_tag {
  mxClassID cls = mxUNKNOWN_CLASS;
  std::vector<mwSize> dims;
  std::vector<unsigned char> data;                               /* numeric / char / logical payload */
  std::vector<mxArray_tag*> cells;                               /* cell elements or struct field values */
  std::vector<std::string> fields;                               /* struct field names (1x1 structs only) */
  std::string class_name;                                        /* for mxIsClass on opaque objects */
  std::vector<std::pair<std::string, mxArray_tag*>> properties;  /* mxGetProperty on objects */
  bool is_complex = false;
};
typedef struct mxArray_tag mxArray;

struct mexShimError : public std::runtime_error {
  std::string id;
  mexShimError(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};

static inline size_t mxshim_elem_size(mxClassID c) {
  switch (c) {
    case mxDOUBLE_CLASS: case mxINT64_CLASS: case mxUINT64_CLASS: return 8;
    case mxSINGLE_CLASS: case mxINT32_CLASS: case mxUINT32_CLASS: return 4;
    case mxINT16_CLASS: case mxUINT16_CLASS: case mxCHAR_CLASS: return 2;
    case mxINT8_CLASS: case mxUINT8_CLASS: case mxLOGICAL_CLASS: return 1;
    default: return 0;
  }
}

static inline mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity) {
  mxArray* a = new mxArray();
  a->cls = cls;
  a->dims = {m, n};
  a->data.assign(m * n * mxshim_elem_size(cls), 0);
  return a;
}
static inline mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
  return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c);
}
static inline mxArray* mxCreateDoubleScalar(double v) {
  mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
  std::memcpy(a->data.data(), &v, 8);
  return a;
}
static inline mxArray* mxCreateLogicalMatrix(mwSize m, mwSize n) {
  return mxCreateNumericMatrix(m, n, mxLOGICAL_CLASS, mxREAL);
}
static inline mxArray* mxCreateCellMatrix(mwSize m, mwSize n) {
  mxArray* a = new mxArray();
  a->cls = mxCELL_CLASS;
  a->dims = {m, n};
  a->cells.assign(m * n, nullptr);
  return a;
}
static inline mxArray* mxCreateString(const char* s) {
  mxArray* a = new mxArray();
  a->cls = mxCHAR_CLASS;
  size_t n = std::strlen(s);
  a->dims = {(mwSize)1, (mwSize)n};
  a->data.resize(2 * n);
  for (size_t i = 0; i < n; ++i) {
    mxChar c = (mxChar)(unsigned char)s[i];
    std::memcpy(a->data.data() + 2 * i, &c, 2);
  }
  return a;
}
static inline mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names) {
  mxArray* a = new mxArray();
  a->cls = mxSTRUCT_CLASS;
  a->dims = {m, n};
  for (int i = 0; i < nfields; ++i) a->fields.push_back(names[i]);
  a->cells.assign((size_t)nfields * m * n, nullptr);
  return a;
}
static inline void mxDestroyArray(mxArray* a) {
  if (!a) return;
  for (mxArray* c : a->cells) mxDestroyArray(c);
  for (auto& p : a->properties) mxDestroyArray(p.second);
  delete a;
}

static inline mwSize mxGetM(const mxArray* a) { return a->dims.empty() ? 0 : a->dims[0]; }
static inline mwSize mxGetN(const mxArray* a) {
  mwSize n = 1;
  for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i];
  return a->dims.size() < 2 ? 0 : n;
}
static inline mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
static inline const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
static inline size_t mxGetNumberOfElements(const mxArray* a) {
  size_t n = 1;
  for (mwSize d : a->dims) n *= d;
  return a->dims.empty() ? 0 : n;
}
static inline bool mxIsEmpty(const mxArray* a) { return mxGetNumberOfElements(a) == 0; }
static inline void* mxGetData(const mxArray* a) { return (void*)a->data.data(); }
static inline double* mxGetPr(const mxArray* a) { return (double*)a->data.data(); }
static inline mxClassID mxGetClassID(const mxArray* a) { return a->cls; }
static inline bool mxIsComplex(const mxArray* a) { return a->is_complex; }
static inline bool mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
static inline bool mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
static inline bool mxIsUint8(const mxArray* a) { return a->cls == mxUINT8_CLASS; }
static inline bool mxIsUint32(const mxArray* a) { return a->cls == mxUINT32_CLASS; }
static inline bool mxIsLogical(const mxArray* a) { return a->cls == mxLOGICAL_CLASS; }
static inline bool mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
static inline bool mxIsCell(const mxArray* a) { return a->cls == mxCELL_CLASS; }
static inline bool mxIsStruct(const mxArray* a) { return a->cls == mxSTRUCT_CLASS; }
static inline bool mxIsNumeric(const mxArray* a) { return a->cls >= mxDOUBLE_CLASS; }
static inline bool mxIsClass(const mxArray* a, const char* name) {
  if (!a->class_name.empty()) return a->class_name == name;
  switch (a->cls) {
    case mxDOUBLE_CLASS: return !std::strcmp(name, "double");
    case mxSINGLE_CLASS: return !std::strcmp(name, "single");
    case mxUINT8_CLASS: return !std::strcmp(name, "uint8");
    case mxUINT32_CLASS: return !std::strcmp(name, "uint32");
    case mxCELL_CLASS: return !std::strcmp(name, "cell");
    case mxSTRUCT_CLASS: return !std::strcmp(name, "struct");
    case mxCHAR_CLASS: return !std::strcmp(name, "char");
    case mxLOGICAL_CLASS: return !std::strcmp(name, "logical");
    default: return false;
  }
}
static inline double mxGetScalar(const mxArray* a) {
  const unsigned char* p = a->data.data();
  switch (a->cls) {
    case mxDOUBLE_CLASS: { double v; std::memcpy(&v, p, 8); return v; }
    case mxSINGLE_CLASS: { float v; std::memcpy(&v, p, 4); return v; }
    case mxUINT8_CLASS: case mxLOGICAL_CLASS: return (double)p[0];
    case mxINT32_CLASS: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case mxUINT32_CLASS: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    default: return 0.0;
  }
}
static inline bool mxIsLogicalScalar(const mxArray* a) { return a->cls == mxLOGICAL_CLASS && mxGetNumberOfElements(a) == 1; }
static inline bool mxIsLogicalScalarTrue(const mxArray* a) { return mxIsLogicalScalar(a) && a->data[0] != 0; }
static inline mxArray* mxGetCell(const mxArray* a, mwIndex i) { return i < a->cells.size() ? a->cells[i] : nullptr; }
static inline void mxSetCell(mxArray* a, mwIndex i, mxArray* v) {
  if (i < a->cells.size()) a->cells[i] = v;
}
static inline int mxGetFieldNumber(const mxArray* a, const char* name) {
  for (size_t i = 0; i < a->fields.size(); ++i)
    if (a->fields[i] == name) return (int)i;
  return -1;
}
static inline mxArray* mxGetField(const mxArray* a, mwIndex idx, const char* name) {
  int f = mxGetFieldNumber(a, name);
  if (f < 0) return nullptr;
  return a->cells[idx * a->fields.size() + (size_t)f];
}
static inline void mxSetField(mxArray* a, mwIndex idx, const char* name, mxArray* v) {
  int f = mxGetFieldNumber(a, name);
  if (f >= 0) a->cells[idx * a->fields.size() + (size_t)f] = v;
}
static inline mxArray* mxGetProperty(const mxArray* a, mwIndex, const char* name) {
  for (auto& p : a->properties)
    if (p.first == name) return p.second;
  return nullptr;
}
static inline char* mxArrayToString(const mxArray* a) {
  if (a->cls != mxCHAR_CLASS) return nullptr;
  size_t n = mxGetNumberOfElements(a);
  char* s = (char*)std::malloc(n + 1);
  for (size_t i = 0; i < n; ++i) {
    mxChar c;
    std::memcpy(&c, a->data.data() + 2 * i, 2);
    s[i] = (char)c;
  }
  s[n] = 0;
  return s;
}
static inline void mxFree(void* p) { std::free(p); }
static inline void* mxMalloc(size_t n) { return std::malloc(n); }
static inline void* mxCalloc(size_t n, size_t s) { return std::calloc(n, s); }
static inline double mxGetNaN(void) { return std::numeric_limits<double>::quiet_NaN(); }
static inline double mxGetInf(void) { return std::numeric_limits<double>::infinity(); }
static inline double mxGetEps(void) { return std::numeric_limits<double>::epsilon(); }

static inline void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw mexShimError(id ? id : "", buf);
}
static inline void mexErrMsgTxt(const char* msg) { throw mexShimError("", msg); }
static inline void mexWarnMsgIdAndTxt(const char*, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vfprintf(stderr, fmt, ap);
  va_end(ap);
}
static inline int mexPrintf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int r = std::vprintf(fmt, ap);
  va_end(ap);
  return r;
}
static inline int mexAtExit(void (*fn)(void)) { return std::atexit(fn); }
static inline void mexLock(void) {}
static inline void mexUnlock(void) {}

#ifdef __cplusplus
extern "C"
#endif
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

#endif /* APS_MEXSHIM_MEX_H */
