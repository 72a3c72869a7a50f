/* TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for libjpeg's <jpeglib.h>: just the declarations the reference's *headers*
 * (lib/include/ultrahdr/jpeg{en,de}coderhelper.h) need in order to parse.  The reference's own
 * jpegencoderhelper.cpp / jpegdecoderhelper.cpp are NOT compiled (libjpeg-turbo is not in this
 * image as a dev package); jpeg_helpers_shim.cpp provides those two classes on top of
 * oracle/jpeg_oracle.c instead. */
#ifndef UHDR_ORACLE_JPEGLIB_STUB_H
#define UHDR_ORACLE_JPEGLIB_STUB_H
#include <stddef.h>
typedef unsigned char JOCTET;
typedef unsigned char JSAMPLE;
typedef JSAMPLE* JSAMPROW;
typedef JSAMPROW* JSAMPARRAY;
typedef unsigned int JDIMENSION;
typedef int boolean;
#define DCTSIZE 8
#define JPEG_LIB_VERSION 62 /* libjpeg-turbo default API level */
struct jpeg_compress_struct;
struct jpeg_decompress_struct;
typedef struct jpeg_compress_struct* j_compress_ptr;
typedef struct jpeg_decompress_struct* j_decompress_ptr;
struct jpeg_destination_mgr {
  JOCTET* next_output_byte;
  size_t free_in_buffer;
  void (*init_destination)(j_compress_ptr);
  boolean (*empty_output_buffer)(j_compress_ptr);
  void (*term_destination)(j_compress_ptr);
};
#endif
