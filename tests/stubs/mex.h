/* test stub (tests/test_callers_compile.py): the part of the Matlab MEX C API the reference's wrappers use, for a syntax-only
   compile of /root/reference/matlabWrapper/{write,read}LFMstack.cpp and readLFMheader.cpp against this repository's include/ */
#ifndef STUB_MEX_H
#define STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS, mxDOUBLE_CLASS, mxSINGLE_CLASS,
               mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS, mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;
typedef uint16_t mxChar;
#ifdef __cplusplus
extern "C" {
#endif
double* mxGetPr(const mxArray*);
void* mxGetData(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsChar(const mxArray*);
void mexErrMsgTxt(const char*);
void mexWarnMsgTxt(const char*);
int mexPrintf(const char*, ...);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetM(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
mxClassID mxGetClassID(const mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateString(const char*);
void mxSetFieldByNumber(mxArray*, mwIndex, int, mxArray*);
char* mxArrayToString(const mxArray*);
void mxFree(void*);
void* mxMalloc(size_t);
void* mxCalloc(size_t, size_t);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
