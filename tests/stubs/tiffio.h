/* test stub (tests/test_callers_compile.py): the libtiff declarations the reference's test program uses, for a syntax-only
   compile of /root/reference/test/mainTest_lfmIO.cxx against this repository's include/ */
#ifndef STUB_TIFFIO_H
#define STUB_TIFFIO_H
#include <stdint.h>
typedef struct tiff TIFF;
typedef uint32_t ttag_t;
#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
TIFF* TIFFOpen(const char*, const char*);
void TIFFClose(TIFF*);
int TIFFGetField(TIFF*, ttag_t, ...);
uint16_t TIFFNumberOfDirectories(TIFF*);
int TIFFReadDirectory(TIFF*);
int TIFFReadScanline(TIFF*, void*, uint32_t, uint16_t sample = 0);
#endif
