/* Implementation of the mex.h stand-in (see mex.h).  TEST INFRASTRUCTURE. */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define MAXF 32
struct mxArray_tag {
    mxClassID cls;
    size_t m, n, elsize;
    void *data;
    int nfields;
    char names[MAXF][24];
    mxArray *vals[MAXF];
};
static char g_msg[512];
static jmp_buf g_jmp;
static int g_armed;
static void (*g_atexit)(void);
static size_t elsize(mxClassID c) { return c == mxDOUBLE_CLASS || c == mxUINT64_CLASS ? 8 : c == mxINT32_CLASS ? 4 : c == mxCHAR_CLASS ? 1 : 0; }
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) {
    mxArray *a = (mxArray *)calloc(1, sizeof *a);
    (void)c;
    a->cls = cls; a->m = m; a->n = n; a->elsize = elsize(cls);
    a->data = calloc((m * n) > 0 ? m * n : 1, a->elsize ? a->elsize : 1);
    return a;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c); }
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c) {
    size_t n = 1, i;
    for (i = 1; i < ndim; ++i) n *= dims[i];
    return mxCreateNumericMatrix(ndim ? dims[0] : 0, ndim ? n : 0, cls, c);
}
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names) {
    mxArray *a = (mxArray *)calloc(1, sizeof *a);
    int i;
    a->cls = mxSTRUCT_CLASS; a->m = m; a->n = n; a->nfields = nfields;
    for (i = 0; i < nfields && i < MAXF; ++i) strncpy(a->names[i], names[i], 23);
    return a;
}
mxArray *mxCreateString(const char *s) {
    mxArray *a = mxCreateNumericMatrix(1, strlen(s), mxCHAR_CLASS, mxREAL);
    memcpy(a->data, s, strlen(s));
    return a;
}
void mxDestroyArray(mxArray *a) {
    int i;
    if (!a) return;
    for (i = 0; i < a->nfields; ++i) mxDestroyArray(a->vals[i]);
    free(a->data); free(a);
}
double *mxGetPr(const mxArray *a) { return (double *)a->data; }
void *mxGetData(const mxArray *a) { return a->data; }
double mxGetScalar(const mxArray *a) {
    if (a->cls == mxDOUBLE_CLASS) return *(double *)a->data;
    if (a->cls == mxINT32_CLASS) return *(int32_t *)a->data;
    if (a->cls == mxUINT64_CLASS) return (double)*(uint64_t *)a->data;
    return 0;
}
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
int mxIsEmpty(const mxArray *a) { return a->m * a->n == 0; }
int mxIsDouble(const mxArray *a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsComplex(const mxArray *a) { (void)a; return 0; }
int mxIsStruct(const mxArray *a) { return a->cls == mxSTRUCT_CLASS; }
int mxIsUint64(const mxArray *a) { return a->cls == mxUINT64_CLASS; }
mxArray *mxGetField(const mxArray *a, mwSize idx, const char *name) {
    int i;
    (void)idx;
    for (i = 0; i < a->nfields; ++i)
        if (strcmp(a->names[i], name) == 0) return a->vals[i];
    return NULL;
}
void mxSetField(mxArray *a, mwSize idx, const char *name, mxArray *v) {
    int i;
    (void)idx;
    for (i = 0; i < a->nfields; ++i)
        if (strcmp(a->names[i], name) == 0) { a->vals[i] = v; return; }
}
int mxGetString(const mxArray *a, char *buf, mwSize buflen) {
    size_t n = a->m * a->n;
    if (a->cls != mxCHAR_CLASS || n + 1 > buflen) return 1;
    memcpy(buf, a->data, n); buf[n] = 0;
    return 0;
}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    va_list ap;
    int k = snprintf(g_msg, sizeof g_msg, "%s: ", id);
    va_start(ap, fmt);
    vsnprintf(g_msg + k, sizeof g_msg - k, fmt, ap);
    va_end(ap);
    if (g_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "%s\n", g_msg);
    abort();
}
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
const char *mexstub_last_error(void) { return g_msg; }
int mexstub_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    g_msg[0] = 0;
    g_armed = 1;
    if (setjmp(g_jmp)) { g_armed = 0; return 1; }
    mexFunction(nlhs, plhs, nrhs, prhs);
    g_armed = 0;
    return 0;
}
void mexstub_run_atexit(void) { if (g_atexit) g_atexit(); }

/* ---- helpers for the ctypes driver: build arrays / structs from Python ---- */
mxArray *mexstub_double(const double *src, size_t m, size_t n) {
    mxArray *a = mxCreateDoubleMatrix(m, n, mxREAL);
    if (src && (m * n) > 0) memcpy(a->data, src, 8 * m * n);
    return a;
}
mxArray *mexstub_struct(int nfields, const char **names) { return mxCreateStructMatrix(1, 1, nfields, names); }
void mexstub_copy_out(const mxArray *a, void *dst) { memcpy(dst, a->data, a->m * a->n * a->elsize); }
