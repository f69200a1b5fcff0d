// Minimal stand-in for MATLAB's mex.h / matrix.h so that the MEX gateway
// (tinympc-matlab_b200/matlab/bindings.cpp) compiles and runs in an image without MATLAB.
// TEST INFRASTRUCTURE: column-major arrays, only the API subset the gateway uses; errors throw.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

typedef size_t mwSize;
enum mxClassID { mxDOUBLE_CLASS, mxSINGLE_CLASS, mxINT32_CLASS, mxCHAR_CLASS };
enum mxComplexity { mxREAL, mxCOMPLEX };

struct mxArray {
    mxClassID cls = mxDOUBLE_CLASS;
    std::vector<mwSize> dims;
    std::vector<unsigned char> bytes;
    std::string str;
    size_t numel() const { size_t n = 1; for (mwSize d : dims) n *= d; return dims.empty() ? 0 : n; }
};

inline size_t mx_elem_size(mxClassID c) { return c == mxDOUBLE_CLASS ? 8 : (c == mxCHAR_CLASS ? 1 : 4); }
inline mxArray* mxCreateNumericArray(mwSize nd, const mwSize* d, mxClassID c, mxComplexity) {
    mxArray* a = new mxArray();
    a->cls = c; a->dims.assign(d, d + nd);
    a->bytes.assign(a->numel() * mx_elem_size(c), 0);
    return a;
}
inline mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity cx) { const mwSize d[2] = {m, n}; return mxCreateNumericArray(2, d, mxDOUBLE_CLASS, cx); }
inline mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); std::memcpy(a->bytes.data(), &v, 8); return a; }
inline mxArray* mxCreateString(const char* s) { mxArray* a = new mxArray(); a->cls = mxCHAR_CLASS; a->str = s; a->dims = {1, std::strlen(s)}; return a; }
inline void mxDestroyArray(mxArray* a) { delete a; }
inline bool mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
inline bool mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
inline bool mxIsInt32(const mxArray* a) { return a->cls == mxINT32_CLASS; }
inline bool mxIsComplex(const mxArray*) { return false; }
inline size_t mxGetM(const mxArray* a) { return a->dims.empty() ? 0 : a->dims[0]; }
inline size_t mxGetN(const mxArray* a) { if (a->dims.size() < 2) return a->dims.empty() ? 0 : 1; size_t n = 1; for (size_t k = 1; k < a->dims.size(); ++k) n *= a->dims[k]; return n; }
inline size_t mxGetNumberOfElements(const mxArray* a) { return a->numel(); }
inline size_t mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
inline const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
inline double* mxGetPr(const mxArray* a) { return reinterpret_cast<double*>(const_cast<unsigned char*>(a->bytes.data())); }
inline void* mxGetData(const mxArray* a) { return const_cast<unsigned char*>(a->bytes.data()); }
inline double mxGetScalar(const mxArray* a) {
    if (a->cls == mxDOUBLE_CLASS) return mxGetPr(a)[0];
    if (a->cls == mxSINGLE_CLASS) return reinterpret_cast<const float*>(a->bytes.data())[0];
    return reinterpret_cast<const int*>(a->bytes.data())[0];
}
inline char* mxArrayToString(const mxArray* a) { char* s = static_cast<char*>(std::malloc(a->str.size() + 1)); std::strcpy(s, a->str.c_str()); return s; }
inline void mxFree(void* p) { std::free(p); }
inline int mexPrintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); int n = std::vprintf(fmt, ap); va_end(ap); return n; }
struct MexError : std::runtime_error { std::string id; MexError(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {} };
[[noreturn]] inline void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw MexError(id, buf);
}
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
