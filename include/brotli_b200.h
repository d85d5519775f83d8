/* brotli_b200.h -- C ABI of the B200-native brotli compression path.
 *
 * This library is a drop-in for the COMPRESSION entry points that the reference (dropbox/rust-brotli 8.0.4)
 * exports from its cdylib; each declaration cites the reference interface it replaces.  Decompression, the
 * BroCatli concatenator and the CLI are out of scope.  All pointers are plain host
 * pointers unless a function says otherwise; no CUDA or torch types appear in any signature.
 *
 * Failure behaviour mirrors the reference: functions return BROTLI_FALSE / NULL / 0 on any error (including
 * "no usable CUDA device" -- there is no CPU fallback), and never abort the process
 * (src/ffi/compressor.rs:253-256, :419-422).
 */
#ifndef BROTLI_B200_H_
#define BROTLI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BROTLI_BOOL int
#define BROTLI_TRUE 1
#define BROTLI_FALSE 0

/* c/brotli/encode.h ; src/enc/encode.rs BrotliEncoderMode */
typedef enum BrotliEncoderMode { BROTLI_MODE_GENERIC = 0, BROTLI_MODE_TEXT = 1, BROTLI_MODE_FONT = 2 } BrotliEncoderMode;

/* src/enc/encode.rs:1380-1385 */
typedef enum BrotliEncoderOperation {
  BROTLI_OPERATION_PROCESS = 0,
  BROTLI_OPERATION_FLUSH = 1,
  BROTLI_OPERATION_FINISH = 2,
  BROTLI_OPERATION_EMIT_METADATA = 3
} BrotliEncoderOperation;

/* src/enc/parameters.rs:1-32 (same numeric values as c/brotli/encode.h) */
typedef enum BrotliEncoderParameter {
  BROTLI_PARAM_MODE = 0,
  BROTLI_PARAM_QUALITY = 1,
  BROTLI_PARAM_LGWIN = 2,
  BROTLI_PARAM_LGBLOCK = 3,
  BROTLI_PARAM_DISABLE_LITERAL_CONTEXT_MODELING = 4,
  BROTLI_PARAM_SIZE_HINT = 5,
  BROTLI_PARAM_LARGE_WINDOW = 6,
  BROTLI_PARAM_Q9_5 = 150,
  BROTLI_METABLOCK_CALLBACK = 151,
  BROTLI_PARAM_STRIDE_DETECTION_QUALITY = 152,
  BROTLI_PARAM_HIGH_ENTROPY_DETECTION_QUALITY = 153,
  BROTLI_PARAM_LITERAL_BYTE_SCORE = 154,
  BROTLI_PARAM_CDF_ADAPTATION_DETECTION = 155,
  BROTLI_PARAM_PRIOR_BITMASK_DETECTION = 156,
  BROTLI_PARAM_SPEED = 157,
  BROTLI_PARAM_SPEED_MAX = 158,
  BROTLI_PARAM_CM_SPEED = 159,
  BROTLI_PARAM_CM_SPEED_MAX = 160,
  BROTLI_PARAM_SPEED_LOW = 161,
  BROTLI_PARAM_SPEED_LOW_MAX = 162,
  BROTLI_PARAM_CM_SPEED_LOW = 164,
  BROTLI_PARAM_CM_SPEED_LOW_MAX = 165,
  BROTLI_PARAM_AVOID_DISTANCE_PREFIX_SEARCH = 166,
  BROTLI_PARAM_CATABLE = 167,
  BROTLI_PARAM_APPENDABLE = 168,
  BROTLI_PARAM_MAGIC_NUMBER = 169,
  BROTLI_PARAM_NO_DICTIONARY = 170,
  BROTLI_PARAM_FAVOR_EFFICIENCY = 171,
  BROTLI_PARAM_BYTE_ALIGN = 172,
  BROTLI_PARAM_BARE_STREAM = 173
} BrotliEncoderParameter;

typedef void* (*brotli_alloc_func)(void* opaque, size_t size);
typedef void (*brotli_free_func)(void* opaque, void* address);

typedef struct BrotliEncoderStateStruct BrotliEncoderState;
typedef struct BrotliEncoderWorkPoolStruct BrotliEncoderWorkPool;

/* ---- single stream: src/ffi/compressor.rs ---- */
/* :72  BrotliEncoderCreateInstance.  Host-side bookkeeping uses alloc_func when given; device memory is owned by the state. */
BrotliEncoderState* BrotliEncoderCreateInstance(brotli_alloc_func alloc_func, brotli_free_func free_func, void* opaque);
/* :115 BrotliEncoderSetParameter (refused after the first byte was consumed, encode.rs:289-295).  Also refused
 * (BROTLI_FALSE, state unchanged): a value this path cannot honour -- LARGE_WINDOW != 0, LGBLOCK outside 0 / 16..24.
 * Framing parameters act as in the reference (encode.rs:264-283, :559-568, :1928-1940, :2258-2333): CATABLE (first two bytes as an
 * uncompressed metablock, no static dictionary, implies APPENDABLE), APPENDABLE, MAGIC_NUMBER (metadata metablock e1 97 8x,
 * VERSION, size hint), BYTE_ALIGN (padding metablock in front of the final empty one), BARE_STREAM (no final metablock; with
 * CATABLE no window bits either).  Streams made with CATABLE go through the reference's BroCatli (src/concat/mod.rs).
 * QUALITY: 5..9 run the hash-chain family, 10 and 11 the optimal-parse family; values below 5 run as 5 (the q0..q4
 * hashers are not built) -- b200_effective_quality() reports the quality that will really be used. */
BROTLI_BOOL BrotliEncoderSetParameter(BrotliEncoderState* state, BrotliEncoderParameter p, uint32_t value);
/* :128 */
void BrotliEncoderDestroyInstance(BrotliEncoderState* state);
/* :141 / encode.rs:1273 */
size_t BrotliEncoderMaxCompressedSize(size_t input_size);
/* :194 BrotliEncoderCompress -- one-shot; falls back to an uncompressed stream when the result would not fit
 * (encode.rs:1528-1536) */
BROTLI_BOOL BrotliEncoderCompress(int quality, int lgwin, BrotliEncoderMode mode, size_t input_size, const uint8_t* input_buffer,
                                  size_t* encoded_size, uint8_t* encoded_buffer);
/* :280 BrotliEncoderCompressStream */
BROTLI_BOOL BrotliEncoderCompressStream(BrotliEncoderState* state, BrotliEncoderOperation op, size_t* available_in,
                                        const uint8_t** next_in, size_t* available_out, uint8_t** next_out, size_t* total_out);
/* :260 BrotliEncoderCompressStreaming -- CompressStream with the buffer pointers passed by value and no total_out */
BROTLI_BOOL BrotliEncoderCompressStreaming(BrotliEncoderState* state, BrotliEncoderOperation op, size_t* available_in,
                                           const uint8_t* input_buf, size_t* available_out, uint8_t* output_buf);
/* :162 BrotliEncoderSetCustomDictionary -- the last min(size, 2^lgwin - 16) bytes of dict become window content in front
 * of the stream; the static dictionary is switched off (encode.rs:1205-1260).  Must precede the first input byte. */
void BrotliEncoderSetCustomDictionary(BrotliEncoderState* state, size_t size, const uint8_t* dict);
/* :359-419 host memory through the instance's allocator (malloc / free when none was given) */
uint8_t* BrotliEncoderMallocU8(BrotliEncoderState* state, size_t size);
void BrotliEncoderFreeU8(BrotliEncoderState* state, uint8_t* data, size_t size);
size_t* BrotliEncoderMallocUsize(BrotliEncoderState* state, size_t size);
void BrotliEncoderFreeUsize(BrotliEncoderState* state, size_t* data, size_t size);
/* :150-192 */
BROTLI_BOOL BrotliEncoderIsFinished(BrotliEncoderState* state);
BROTLI_BOOL BrotliEncoderHasMoreOutput(BrotliEncoderState* state);
const uint8_t* BrotliEncoderTakeOutput(BrotliEncoderState* state, size_t* size);
uint32_t BrotliEncoderVersion(void);

/* ---- multi-shard: src/ffi/multicompress/mod.rs ---- */
/* :49 */
size_t BrotliEncoderMaxCompressedSizeMulti(size_t input_size, size_t num_threads);
/* :93  shards = desired_num_threads (<= 16, fixed_queue.rs:1); shard i covers [i*len/n, (i+1)*len/n)
 * (threading/mod.rs:333) and sees the previous 2^lgwin bytes as its window; shards are placed round-robin on the
 * visible GPUs. */
int32_t BrotliEncoderCompressMulti(size_t num_params, const BrotliEncoderParameter* param_keys, const uint32_t* param_values,
                                   size_t input_size, const uint8_t* input, size_t* encoded_size, uint8_t* encoded,
                                   size_t desired_num_threads, brotli_alloc_func alloc_func, brotli_free_func free_func,
                                   void** alloc_opaque_per_thread);
/* :240, :294, :312 */
BrotliEncoderWorkPool* BrotliEncoderCreateWorkPool(size_t num_workers, brotli_alloc_func alloc_func, brotli_free_func free_func,
                                                   void** alloc_opaque_per_thread);
void BrotliEncoderDestroyWorkPool(BrotliEncoderWorkPool* work_pool);
int32_t BrotliEncoderCompressWorkPool(BrotliEncoderWorkPool* work_pool, size_t num_params, const BrotliEncoderParameter* param_keys,
                                      const uint32_t* param_values, size_t input_size, const uint8_t* input, size_t* encoded_size,
                                      uint8_t* encoded, size_t desired_num_threads, brotli_alloc_func alloc_func,
                                      brotli_free_func free_func, void** alloc_opaque_per_thread);

/* ---- device-resident entry points (B200 additions; pointers are CUDA device pointers where noted) ---- */
typedef struct B200Encoder B200Encoder;
int b200_device_count(void);
int b200_effective_quality(int requested_quality);
B200Encoder* b200_encoder_create(int device);
void b200_encoder_destroy(B200Encoder* e);
int b200_encoder_device(const B200Encoder* e); /* CUDA ordinal the encoder lives on */
/* bit d set: device d compressed at least one shard of the last BrotliEncoderCompressMulti / CompressWorkPool call (diagnostic) */
uint32_t b200_last_multi_device_mask(void);
/* development switches (csrc/bro_encoder.h: A/B of kernel variants, stage timing, lanes); the defaults are the product configuration */
int b200_encoder_set_option(B200Encoder* e, int option, uint32_t value);
size_t b200_max_compressed_size(size_t n);
/* device_io: 0 = in / out are host pointers; 1 = both are device pointers on the encoder's GPU; 2 = host input, device output
 * (e.g. shard outputs that travel on to a peer GPU over NVLink); 3 = device input, host output */
int b200_encoder_compress(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint8_t* out, size_t out_cap,
                          size_t* out_size, int device_io);
int b200_encoder_compress_range(B200Encoder* e, int quality, int lgwin, uint64_t size_hint, const uint8_t* in, size_t n,
                                size_t range_start, size_t range_len, int first, int last, int byte_align, uint8_t* out,
                                size_t out_cap, size_t* out_size, int device_io);
int b200_encoder_last_timings(B200Encoder* e, float* ms, uint32_t* launches);
/* stage hook used by the parity tests: per-position best bucket match (distance << 8 | capped length) */
int b200_stage_match(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint32_t* best_out);

#ifdef __cplusplus
}
#endif
#endif /* BROTLI_B200_H_ */
