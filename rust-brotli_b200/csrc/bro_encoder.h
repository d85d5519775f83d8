/* bro_encoder.h -- internal C interface of the device encoder (the public C ABI is include/brotli_b200.h). */
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct B200Encoder B200Encoder;
enum { B200_OPT_UNIT = 1, B200_OPT_MB_UNITS = 2, B200_OPT_LCAP = 3, B200_OPT_RLE_OPT = 4, B200_OPT_SPLIT = 5,
       B200_OPT_CTX_MODEL = 6, B200_OPT_TIMING = 7, B200_OPT_LANES = 8, B200_OPT_DICT = 9, B200_OPT_SHALLOW_MATCH = 10, B200_OPT_PAIR_PARSE = 11, B200_OPT_HQ_SPLIT = 12, B200_OPT_HQ_UNIT = 13, B200_OPT_HQ_THREAD_UNITS = 14, B200_OPT_ONDEMAND = 15, B200_OPT_HQ_LEVELS = 16 };
/* stage timing slots of b200_encoder_last_timings */
enum { B200_ST_SORT = 0, B200_ST_MATCH = 1, B200_ST_PARSE = 2, B200_ST_FINALIZE = 3, B200_ST_SPLIT = 4, B200_ST_HEADER = 5,
       B200_ST_EMIT = 6, B200_NUM_STAGES = 7 };
int b200_device_count(void);
int b200_effective_quality(int requested_quality);
B200Encoder* b200_encoder_create(int device);
void b200_encoder_destroy(B200Encoder* e);
int b200_encoder_set_option(B200Encoder* e, int option, uint32_t value);
size_t b200_max_compressed_size(size_t n);
int b200_encoder_compress(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint8_t* out, size_t out_cap,
                          size_t* out_size, int device_io);
int b200_encoder_compress_range(B200Encoder* e, int quality, int lgwin, uint64_t size_hint, const uint8_t* in, size_t n,
                                size_t range_start, size_t range_len, int first, int last, int byte_align, uint8_t* out,
                                size_t out_cap, size_t* out_size, int device_io);
int b200_encoder_last_timings(B200Encoder* e, float* ms /* [B200_NUM_STAGES] */, uint32_t* launches);
int b200_stage_match(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint32_t* best_out);
#ifdef __cplusplus
}
#endif
