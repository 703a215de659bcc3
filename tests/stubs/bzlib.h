/* test stub (tests/test_callers_compile.py): the reference's test program includes bzlib.h but calls nothing from it */
#ifndef STUB_BZLIB_H
#define STUB_BZLIB_H
#endif
