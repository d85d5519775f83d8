/* C client of include/brotli_b200.h, compiled by tests/test_c_client.py with `gcc -Iinclude ... -lbrotli_b200`.
 * Restates what the reference's own C client does (c/multiexample.c:51-116: work pool with custom allocators,
 * BrotliEncoderCompressWorkPool, BrotliEncoderCompressMulti) plus the single-stream calls of c/brotli/encode.h, so the
 * header is exercised by a compiler and the symbols by a linker, not only through ctypes.
 * usage: client <input file> <output prefix>; writes <prefix>.pool, .multi, .oneshot, .stream, .streaming
 * exit code: 0 ok, 77 no CUDA device (every entry point failed the way the header says), 1 anything else. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "brotli_b200.h"

static size_t g_allocs = 0, g_frees = 0;
static void* counting_malloc(void* opaque, size_t size) { (void)opaque; ++g_allocs; return malloc(size); }
static void counting_free(void* opaque, void* p) { (void)opaque; if (p) ++g_frees; free(p); }

static int save(const char* prefix, const char* ext, const uint8_t* p, size_t n) {
  char name[4096];
  snprintf(name, sizeof(name), "%s.%s", prefix, ext);
  FILE* f = fopen(name, "wb");
  if (!f) return 0;
  fwrite(p, 1, n, f);
  fclose(f);
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END);
  size_t len = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* data = (uint8_t*)malloc(len ? len : 1);
  if (fread(data, 1, len, f) != len) return 1;
  fclose(f);

  BrotliEncoderParameter keys[3] = {BROTLI_PARAM_QUALITY, BROTLI_PARAM_LGWIN, BROTLI_PARAM_SIZE_HINT};
  uint32_t values[3] = {5, 22, (uint32_t)len};
  const size_t num_threads = 4;
  void* opaque_per_thread[16] = {0};

  /* ---- work pool (c/multiexample.c:51-97) ---- */
  BrotliEncoderWorkPool* pool = BrotliEncoderCreateWorkPool(num_threads - 1, counting_malloc, counting_free, opaque_per_thread);
  if (!pool) {
    /* no device: the state-less and the instance entry points must fail too, never produce bytes on the CPU */
    size_t cap = BrotliEncoderMaxCompressedSize(len), got = cap;
    uint8_t* out = (uint8_t*)malloc(cap);
    if (BrotliEncoderCompress(5, 22, BROTLI_MODE_GENERIC, len, data, &got, out) != BROTLI_FALSE && len != 0) return 1;
    if (BrotliEncoderCreateInstance(NULL, NULL, NULL) != NULL) return 1;
    return 77;
  }
  size_t cap = BrotliEncoderMaxCompressedSizeMulti(len, num_threads);
  uint8_t* out = (uint8_t*)malloc(cap);
  size_t out_len = cap;
  if (!BrotliEncoderCompressWorkPool(pool, 3, keys, values, len, data, &out_len, out, num_threads, counting_malloc, counting_free,
                                     opaque_per_thread))
    return 1;
  BrotliEncoderDestroyWorkPool(pool);
  if (!save(argv[2], "pool", out, out_len)) return 1;

  /* ---- immediate (c/multiexample.c:99-146) ---- */
  out_len = cap;
  if (!BrotliEncoderCompressMulti(3, keys, values, len, data, &out_len, out, num_threads, NULL, NULL, NULL)) return 1;
  if (!save(argv[2], "multi", out, out_len)) return 1;

  /* ---- one-shot ---- */
  out_len = BrotliEncoderMaxCompressedSize(len);
  if (out_len > cap) return 1;
  if (!BrotliEncoderCompress(5, 22, BROTLI_MODE_GENERIC, len, data, &out_len, out)) return 1;
  if (!save(argv[2], "oneshot", out, out_len)) return 1;

  /* ---- stream with custom allocators, TakeOutput draining ---- */
  {
    BrotliEncoderState* s = BrotliEncoderCreateInstance(counting_malloc, counting_free, NULL);
    if (!s) return 1;
    if (!BrotliEncoderSetParameter(s, BROTLI_PARAM_QUALITY, 5) || !BrotliEncoderSetParameter(s, BROTLI_PARAM_LGWIN, 22)) return 1;
    if (BrotliEncoderSetParameter(s, BROTLI_PARAM_LARGE_WINDOW, 1)) return 1; /* refused, not ignored */
    uint8_t* scratch = BrotliEncoderMallocU8(s, 64);
    if (!scratch) return 1;
    BrotliEncoderFreeU8(s, scratch, 64);
    size_t* scratch2 = BrotliEncoderMallocUsize(s, 8);
    if (!scratch2) return 1;
    BrotliEncoderFreeUsize(s, scratch2, 8);
    size_t total = 0, avail_in = len, avail_out = 0, produced = 0;
    const uint8_t* next_in = data;
    uint8_t* next_out = NULL;
    if (!BrotliEncoderCompressStream(s, BROTLI_OPERATION_FINISH, &avail_in, &next_in, &avail_out, &next_out, &total)) return 1;
    if (avail_in != 0 || BrotliEncoderIsFinished(s)) return 1; /* output still pending */
    while (BrotliEncoderHasMoreOutput(s)) {
      size_t n = 1000; /* in pieces */
      const uint8_t* p = BrotliEncoderTakeOutput(s, &n);
      if (!p || n == 0 || produced + n > cap) return 1;
      memcpy(out + produced, p, n);
      produced += n;
    }
    if (!BrotliEncoderIsFinished(s)) return 1;
    if (BrotliEncoderSetParameter(s, BROTLI_PARAM_QUALITY, 9)) return 1; /* refused after the stream started */
    BrotliEncoderDestroyInstance(s);
    if (!save(argv[2], "stream", out, produced)) return 1;
  }
  /* ---- BrotliEncoderCompressStreaming: pointers by value, small output buffer, cumulative total_out ---- */
  {
    BrotliEncoderState* s = BrotliEncoderCreateInstance(NULL, NULL, NULL);
    if (!s) return 1;
    BrotliEncoderSetParameter(s, BROTLI_PARAM_QUALITY, 5);
    size_t produced = 0, avail_in = len;
    uint8_t buf[4096];
    const uint8_t* in = data;
    for (;;) {
      size_t avail_out = sizeof(buf);
      size_t before_in = avail_in;
      if (!BrotliEncoderCompressStreaming(s, BROTLI_OPERATION_FINISH, &avail_in, in, &avail_out, buf)) return 1;
      in += before_in - avail_in;
      size_t n = sizeof(buf) - avail_out;
      if (produced + n > cap) return 1;
      memcpy(out + produced, buf, n);
      produced += n;
      if (BrotliEncoderIsFinished(s)) break;
      if (n == 0 && !BrotliEncoderHasMoreOutput(s)) return 1;
    }
    BrotliEncoderDestroyInstance(s);
    if (!save(argv[2], "streaming", out, produced)) return 1;
  }
  if (g_allocs == 0 || g_allocs != g_frees) return 1; /* custom allocators were used and balanced */
  printf("ok %zu\n", len);
  return 0;
}
