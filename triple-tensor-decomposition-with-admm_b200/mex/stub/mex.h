/* Minimal stand-in for MathWorks' mex.h / matrix.h, declaring only what the gateways in this
 * directory use, so they can be compiled in an image without MATLAB or Octave.
 * NOT a MEX implementation; the only definitions of these functions live in the test mock tests/mex_mock/mexmock.c. */
#ifndef TRITD_STUB_MEX_H
#define TRITD_STUB_MEX_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
int mxIsDouble(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsSparse(const mxArray*);
int mxIsStruct(const mxArray*);
int mxIsEmpty(const mxArray*);
int mxIsLogical(const mxArray*);
typedef unsigned char mxLogical;
mxLogical* mxGetLogicals(const mxArray*);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxSetField(mxArray*, size_t, const char*, mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
double* mxGetPr(const mxArray*);
double mxGetScalar(const mxArray*);
mxArray* mxGetField(const mxArray*, size_t, const char*);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
void mxDestroyArray(mxArray*);
void* mxMalloc(size_t);
void mxFree(void*);
int mexCallMATLAB(int, mxArray**, int, mxArray**, const char*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexPrintf(const char*, ...);
void mexLock(void);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
