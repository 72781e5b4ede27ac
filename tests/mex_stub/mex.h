/* Minimal stand-in for MATLAB's mex.h (R2018a separate-complex API subset used by lbmpc_mex.c).
 * TEST INFRASTRUCTURE: lets the gateway be compiled and driven from C without MATLAB
 * (tests/test_mex_gateway.py).  Semantics follow the documented MEX API: column-major data,
 * mxGetM = rows, mxGetN = product of the remaining dimensions. */
#ifndef MEX_STUB_H
#define MEX_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef size_t mwSize;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15, mxCHAR_CLASS = 4, mxSTRUCT_CLASS = 2 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef struct mxArray_tag mxArray;
#ifdef __cplusplus
extern "C" {
#endif
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names);
mxArray *mxCreateString(const char *s);
void mxDestroyArray(mxArray *a);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
double mxGetScalar(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
int mxIsEmpty(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsComplex(const mxArray *a);
int mxIsStruct(const mxArray *a);
int mxIsUint64(const mxArray *a);
mxArray *mxGetField(const mxArray *a, mwSize idx, const char *name);
void mxSetField(mxArray *a, mwSize idx, const char *name, mxArray *v);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
/* stub-only helpers for the C driver */
const char *mexstub_last_error(void);
int mexstub_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]); /* 0 ok, 1 = mexErrMsgIdAndTxt raised */
void mexstub_run_atexit(void);
#ifdef __cplusplus
}
#endif
#endif
