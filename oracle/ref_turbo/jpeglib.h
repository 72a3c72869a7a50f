/* TEST INFRASTRUCTURE ONLY.
 *
 * Hand-written declaration of the public libjpeg API, version 62, as libjpeg-turbo 3.1.x exports it
 * (jpeglib.h / jmorecfg.h / jconfig.h: 8-bit samples, `boolean` = int, JPEG_LIB_VERSION 62, the
 * libjpeg-turbo colour-space extensions).  libjpeg-turbo's development headers are not in this image,
 * its shared library is (Pillow bundles 3.1.4.1 as pillow.libs/libjpeg-*.so.62.4.0 with the full
 * libjpeg-62 ABI exported).  This header lets the reference's OWN lib/src/jpegencoderhelper.cpp and
 * lib/src/jpegdecoderhelper.cpp compile unmodified and link against that library, so the checker and
 * the CPU baseline run the reference's JPEG path on the real, SIMD-accelerated libjpeg-turbo.
 *
 * Safety net: jpeg_CreateCompress / jpeg_CreateDecompress take sizeof(struct) and refuse to run when
 * it differs from the library's own (JERR_BAD_STRUCT_SIZE); tests/test_oracle_turbo.py also checks
 * that streams written through this header are byte-identical to Pillow's and to oracle/jpeg_oracle.c.
 */
#ifndef UHDR_ORACLE_JPEGLIB62_H
#define UHDR_ORACLE_JPEGLIB62_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPEG_LIB_VERSION 62
#define LIBJPEG_TURBO_VERSION_NUMBER 3001000
#define C_ARITH_CODING_SUPPORTED 1
#define D_ARITH_CODING_SUPPORTED 1
#define BITS_IN_JSAMPLE 8

/* ---- jmorecfg.h ---- */
#define MAX_COMPONENTS 10
typedef unsigned char JSAMPLE;
typedef short J12SAMPLE;
typedef unsigned short J16SAMPLE;
typedef short JCOEF;
typedef unsigned char JOCTET;
typedef unsigned char UINT8;
typedef unsigned short UINT16;
typedef short INT16;
typedef long INT32;
typedef unsigned int JDIMENSION;
#define JPEG_MAX_DIMENSION 65500L
#ifndef FALSE
#define FALSE 0
#endif
#ifndef TRUE
#define TRUE 1
#endif
typedef int boolean;
#define JMETHOD(type, methodname, arglist) type(*methodname) arglist
#define EXTERN(type) extern type

/* ---- jpeglib.h ---- */
#define DCTSIZE 8
#define DCTSIZE2 64
#define NUM_QUANT_TBLS 4
#define NUM_HUFF_TBLS 4
#define NUM_ARITH_TBLS 16
#define MAX_COMPS_IN_SCAN 4
#define MAX_SAMP_FACTOR 4
#define C_MAX_BLOCKS_IN_MCU 10
#define D_MAX_BLOCKS_IN_MCU 10

typedef JSAMPLE* JSAMPROW;
typedef JSAMPROW* JSAMPARRAY;
typedef JSAMPARRAY* JSAMPIMAGE;
typedef JCOEF JBLOCK[DCTSIZE2];
typedef JBLOCK* JBLOCKROW;
typedef JBLOCKROW* JBLOCKARRAY;
typedef JBLOCKARRAY* JBLOCKIMAGE;
typedef JCOEF* JCOEFPTR;

typedef struct {
  UINT16 quantval[DCTSIZE2];
  boolean sent_table;
} JQUANT_TBL;

typedef struct {
  UINT8 bits[17];
  UINT8 huffval[256];
  boolean sent_table;
} JHUFF_TBL;

typedef struct {
  int component_id;
  int component_index;
  int h_samp_factor;
  int v_samp_factor;
  int quant_tbl_no;
  int dc_tbl_no;
  int ac_tbl_no;
  JDIMENSION width_in_blocks;
  JDIMENSION height_in_blocks;
  int DCT_scaled_size; /* API < 70 */
  JDIMENSION downsampled_width;
  JDIMENSION downsampled_height;
  boolean component_needed;
  int MCU_width;
  int MCU_height;
  int MCU_blocks;
  int MCU_sample_width;
  int last_col_width;
  int last_row_height;
  JQUANT_TBL* quant_table;
  void* dct_table;
} jpeg_component_info;

typedef struct {
  int comps_in_scan;
  int component_index[MAX_COMPS_IN_SCAN];
  int Ss, Se;
  int Ah, Al;
} jpeg_scan_info;

typedef struct jpeg_marker_struct* jpeg_saved_marker_ptr;
struct jpeg_marker_struct {
  jpeg_saved_marker_ptr next;
  UINT8 marker;
  unsigned int original_length;
  unsigned int data_length;
  JOCTET* data;
};

#define JCS_EXTENSIONS 1
#define JCS_ALPHA_EXTENSIONS 1
typedef enum {
  JCS_UNKNOWN,
  JCS_GRAYSCALE,
  JCS_RGB,
  JCS_YCbCr,
  JCS_CMYK,
  JCS_YCCK,
  JCS_EXT_RGB,
  JCS_EXT_RGBX,
  JCS_EXT_BGR,
  JCS_EXT_BGRX,
  JCS_EXT_XBGR,
  JCS_EXT_XRGB,
  JCS_EXT_RGBA,
  JCS_EXT_BGRA,
  JCS_EXT_ABGR,
  JCS_EXT_ARGB,
  JCS_RGB565
} J_COLOR_SPACE;

typedef enum { JDCT_ISLOW, JDCT_IFAST, JDCT_FLOAT } J_DCT_METHOD;
#define JDCT_DEFAULT JDCT_ISLOW
#define JDCT_FASTEST JDCT_IFAST
typedef enum { JDITHER_NONE, JDITHER_ORDERED, JDITHER_FS } J_DITHER_MODE;

#define jpeg_common_fields            \
  struct jpeg_error_mgr* err;         \
  struct jpeg_memory_mgr* mem;        \
  struct jpeg_progress_mgr* progress; \
  void* client_data;                  \
  boolean is_decompressor;            \
  int global_state

struct jpeg_common_struct {
  jpeg_common_fields;
};
typedef struct jpeg_common_struct* j_common_ptr;
typedef struct jpeg_compress_struct* j_compress_ptr;
typedef struct jpeg_decompress_struct* j_decompress_ptr;

struct jpeg_compress_struct {
  jpeg_common_fields;
  struct jpeg_destination_mgr* dest;
  JDIMENSION image_width;
  JDIMENSION image_height;
  int input_components;
  J_COLOR_SPACE in_color_space;
  double input_gamma;
  int data_precision;
  int num_components;
  J_COLOR_SPACE jpeg_color_space;
  jpeg_component_info* comp_info;
  JQUANT_TBL* quant_tbl_ptrs[NUM_QUANT_TBLS];
  JHUFF_TBL* dc_huff_tbl_ptrs[NUM_HUFF_TBLS];
  JHUFF_TBL* ac_huff_tbl_ptrs[NUM_HUFF_TBLS];
  UINT8 arith_dc_L[NUM_ARITH_TBLS];
  UINT8 arith_dc_U[NUM_ARITH_TBLS];
  UINT8 arith_ac_K[NUM_ARITH_TBLS];
  int num_scans;
  const jpeg_scan_info* scan_info;
  boolean raw_data_in;
  boolean arith_code;
  boolean optimize_coding;
  boolean CCIR601_sampling;
  int smoothing_factor;
  J_DCT_METHOD dct_method;
  unsigned int restart_interval;
  int restart_in_rows;
  boolean write_JFIF_header;
  UINT8 JFIF_major_version;
  UINT8 JFIF_minor_version;
  UINT8 density_unit;
  UINT16 X_density;
  UINT16 Y_density;
  boolean write_Adobe_marker;
  JDIMENSION next_scanline;
  boolean progressive_mode;
  int max_h_samp_factor;
  int max_v_samp_factor;
  JDIMENSION total_iMCU_rows;
  int comps_in_scan;
  jpeg_component_info* cur_comp_info[MAX_COMPS_IN_SCAN];
  JDIMENSION MCUs_per_row;
  JDIMENSION MCU_rows_in_scan;
  int blocks_in_MCU;
  int MCU_membership[C_MAX_BLOCKS_IN_MCU];
  int Ss, Se, Ah, Al;
  struct jpeg_comp_master* master;
  struct jpeg_c_main_controller* main;
  struct jpeg_c_prep_controller* prep;
  struct jpeg_c_coef_controller* coef;
  struct jpeg_marker_writer* marker;
  struct jpeg_color_converter* cconvert;
  struct jpeg_downsampler* downsample;
  struct jpeg_forward_dct* fdct;
  struct jpeg_entropy_encoder* entropy;
  jpeg_scan_info* script_space;
  int script_space_size;
};

struct jpeg_decompress_struct {
  jpeg_common_fields;
  struct jpeg_source_mgr* src;
  JDIMENSION image_width;
  JDIMENSION image_height;
  int num_components;
  J_COLOR_SPACE jpeg_color_space;
  J_COLOR_SPACE out_color_space;
  unsigned int scale_num, scale_denom;
  double output_gamma;
  boolean buffered_image;
  boolean raw_data_out;
  J_DCT_METHOD dct_method;
  boolean do_fancy_upsampling;
  boolean do_block_smoothing;
  boolean quantize_colors;
  J_DITHER_MODE dither_mode;
  boolean two_pass_quantize;
  int desired_number_of_colors;
  boolean enable_1pass_quant;
  boolean enable_external_quant;
  boolean enable_2pass_quant;
  JDIMENSION output_width;
  JDIMENSION output_height;
  int out_color_components;
  int output_components;
  int rec_outbuf_height;
  int actual_number_of_colors;
  JSAMPARRAY colormap;
  JDIMENSION output_scanline;
  int input_scan_number;
  JDIMENSION input_iMCU_row;
  int output_scan_number;
  JDIMENSION output_iMCU_row;
  int (*coef_bits)[DCTSIZE2];
  JQUANT_TBL* quant_tbl_ptrs[NUM_QUANT_TBLS];
  JHUFF_TBL* dc_huff_tbl_ptrs[NUM_HUFF_TBLS];
  JHUFF_TBL* ac_huff_tbl_ptrs[NUM_HUFF_TBLS];
  int data_precision;
  jpeg_component_info* comp_info;
  boolean progressive_mode;
  boolean arith_code;
  UINT8 arith_dc_L[NUM_ARITH_TBLS];
  UINT8 arith_dc_U[NUM_ARITH_TBLS];
  UINT8 arith_ac_K[NUM_ARITH_TBLS];
  unsigned int restart_interval;
  boolean saw_JFIF_marker;
  UINT8 JFIF_major_version;
  UINT8 JFIF_minor_version;
  UINT8 density_unit;
  UINT16 X_density;
  UINT16 Y_density;
  boolean saw_Adobe_marker;
  UINT8 Adobe_transform;
  boolean CCIR601_sampling;
  jpeg_saved_marker_ptr marker_list;
  int max_h_samp_factor;
  int max_v_samp_factor;
  int min_DCT_scaled_size; /* API < 70 */
  JDIMENSION total_iMCU_rows;
  JSAMPLE* sample_range_limit;
  int comps_in_scan;
  jpeg_component_info* cur_comp_info[MAX_COMPS_IN_SCAN];
  JDIMENSION MCUs_per_row;
  JDIMENSION MCU_rows_in_scan;
  int blocks_in_MCU;
  int MCU_membership[D_MAX_BLOCKS_IN_MCU];
  int Ss, Se, Ah, Al;
  int unread_marker;
  struct jpeg_decomp_master* master;
  struct jpeg_d_main_controller* main;
  struct jpeg_d_coef_controller* coef;
  struct jpeg_d_post_controller* post;
  struct jpeg_input_controller* inputctl;
  struct jpeg_marker_reader* marker;
  struct jpeg_entropy_decoder* entropy;
  struct jpeg_inverse_dct* idct;
  struct jpeg_upsampler* upsample;
  struct jpeg_color_deconverter* cconvert;
  struct jpeg_color_quantizer* cquantize;
};

#define JMSG_LENGTH_MAX 200
#define JMSG_STR_PARM_MAX 80
struct jpeg_error_mgr {
  void (*error_exit)(j_common_ptr cinfo);
  void (*emit_message)(j_common_ptr cinfo, int msg_level);
  void (*output_message)(j_common_ptr cinfo);
  void (*format_message)(j_common_ptr cinfo, char* buffer);
  void (*reset_error_mgr)(j_common_ptr cinfo);
  int msg_code;
  union {
    int i[8];
    char s[JMSG_STR_PARM_MAX];
  } msg_parm;
  int trace_level;
  long num_warnings;
  const char* const* jpeg_message_table;
  int last_jpeg_message;
  const char* const* addon_message_table;
  int first_addon_message;
  int last_addon_message;
};

struct jpeg_progress_mgr {
  void (*progress_monitor)(j_common_ptr cinfo);
  long pass_counter;
  long pass_limit;
  int completed_passes;
  int total_passes;
};

struct jpeg_destination_mgr {
  JOCTET* next_output_byte;
  size_t free_in_buffer;
  void (*init_destination)(j_compress_ptr cinfo);
  boolean (*empty_output_buffer)(j_compress_ptr cinfo);
  void (*term_destination)(j_compress_ptr cinfo);
};

struct jpeg_source_mgr {
  const JOCTET* next_input_byte;
  size_t bytes_in_buffer;
  void (*init_source)(j_decompress_ptr cinfo);
  boolean (*fill_input_buffer)(j_decompress_ptr cinfo);
  void (*skip_input_data)(j_decompress_ptr cinfo, long num_bytes);
  boolean (*resync_to_restart)(j_decompress_ptr cinfo, int desired);
  void (*term_source)(j_decompress_ptr cinfo);
};

EXTERN(struct jpeg_error_mgr*) jpeg_std_error(struct jpeg_error_mgr* err);
#define jpeg_create_compress(cinfo) \
  jpeg_CreateCompress((cinfo), JPEG_LIB_VERSION, (size_t)sizeof(struct jpeg_compress_struct))
#define jpeg_create_decompress(cinfo) \
  jpeg_CreateDecompress((cinfo), JPEG_LIB_VERSION, (size_t)sizeof(struct jpeg_decompress_struct))
EXTERN(void) jpeg_CreateCompress(j_compress_ptr cinfo, int version, size_t structsize);
EXTERN(void) jpeg_CreateDecompress(j_decompress_ptr cinfo, int version, size_t structsize);
EXTERN(void) jpeg_destroy_compress(j_compress_ptr cinfo);
EXTERN(void) jpeg_destroy_decompress(j_decompress_ptr cinfo);
EXTERN(void) jpeg_set_defaults(j_compress_ptr cinfo);
EXTERN(void) jpeg_set_colorspace(j_compress_ptr cinfo, J_COLOR_SPACE colorspace);
EXTERN(void) jpeg_set_quality(j_compress_ptr cinfo, int quality, boolean force_baseline);
EXTERN(void) jpeg_suppress_tables(j_compress_ptr cinfo, boolean suppress);
EXTERN(void) jpeg_start_compress(j_compress_ptr cinfo, boolean write_all_tables);
EXTERN(JDIMENSION) jpeg_write_scanlines(j_compress_ptr cinfo, JSAMPARRAY scanlines, JDIMENSION num_lines);
EXTERN(void) jpeg_finish_compress(j_compress_ptr cinfo);
EXTERN(JDIMENSION) jpeg_write_raw_data(j_compress_ptr cinfo, JSAMPIMAGE data, JDIMENSION num_lines);
EXTERN(void) jpeg_write_marker(j_compress_ptr cinfo, int marker, const JOCTET* dataptr, unsigned int datalen);
EXTERN(int) jpeg_read_header(j_decompress_ptr cinfo, boolean require_image);
#define JPEG_SUSPENDED 0
#define JPEG_HEADER_OK 1
#define JPEG_HEADER_TABLES_ONLY 2
EXTERN(boolean) jpeg_start_decompress(j_decompress_ptr cinfo);
EXTERN(JDIMENSION) jpeg_read_scanlines(j_decompress_ptr cinfo, JSAMPARRAY scanlines, JDIMENSION max_lines);
EXTERN(boolean) jpeg_finish_decompress(j_decompress_ptr cinfo);
EXTERN(JDIMENSION) jpeg_read_raw_data(j_decompress_ptr cinfo, JSAMPIMAGE data, JDIMENSION max_lines);
EXTERN(void) jpeg_save_markers(j_decompress_ptr cinfo, int marker_code, unsigned int length_limit);
EXTERN(void) jpeg_abort_compress(j_compress_ptr cinfo);
EXTERN(void) jpeg_abort_decompress(j_decompress_ptr cinfo);
EXTERN(boolean) jpeg_resync_to_restart(j_decompress_ptr cinfo, int desired);

#define JPEG_RST0 0xD0
#define JPEG_EOI 0xD9
#define JPEG_APP0 0xE0
#define JPEG_COM 0xFE

#ifdef __cplusplus
}
#endif
#endif
