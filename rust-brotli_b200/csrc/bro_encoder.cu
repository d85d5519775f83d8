// bro_encoder.cu -- host orchestration of the B200 brotli compression path and its C ABI.
//
// One Encoder object = one GPU + one CUDA stream + a reusable device workspace.  A stream is compressed as a
// sequence of independent ranges ("chunks", <= 128 MiB) whose match search sees a left halo of the previous
// 2^lgwin bytes; every chunk runs  sort -> match -> parse -> finalise -> context -> symbols -> split -> header ->
// bit lengths -> layout -> emit  entirely on the device and appends its metablocks at the running bit position.
// No stage has a CPU fallback: if CUDA is unavailable every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "bro_kernels.cuh"
#include "bro_encoder.h"

using namespace bro;

#define CUDA_OK(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) {                                                                         \
      fprintf(stderr, "[brotli_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return false;                                                                                  \
    }                                                                                                \
  } while (0)

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + (bytes >> 4) + 4096;
    if (cudaMalloc(&p, want) != cudaSuccess) {
      fprintf(stderr, "[brotli_b200] cudaMalloc(%zu) failed\n", want);
      return false;
    }
    cap = want;
    return true;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr uint32_t kPad = 512;                 // zero bytes after the input
constexpr uint32_t kChunk = 128u << 20;        // bytes per pipeline pass
constexpr uint32_t kBatchMax = 1u << 25;       // positions per sort batch (25-bit packed positions)

}  // namespace

struct B200Encoder {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ok = false;
  // configuration knobs (tests flip these)
  uint32_t unit = 4096, mb_units = 1024, lcap = 64;
  int use_rle_opt = 1, split = 1, ctx_model = 1;
  // buffers
  DevBuf d_data, d_lut, d_sortA, d_sortB, d_hist, d_digit, d_best, d_raw, d_unit, d_cmds, d_cmd_bits, d_lit_syms,
      d_cmd_syms, d_dist_syms, d_mb, d_split_u8, d_split_u32, d_split_counts, d_hist_lit, d_hist_cmd, d_hist_dist,
      d_split_codes, d_codes_u8, d_codes_u16, d_hdr, d_huff_ws, d_ctxmap_ws, d_out, d_total, d_tree_ws, d_tree_bits, d_tree_nbits, d_cmd_tile, d_long_tab, d_seg_bits;
  uint8_t* h_pinned = nullptr;
  size_t h_pinned_cap = 0;
  uint64_t data_base = 0;  // absolute stream position of d_data[0]
  // timing of the last compress call: a mark = (event, stage that starts there); -1 ends the last stage
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> mark_stage;
  float stage_ms[B200_NUM_STAGES];
  uint32_t launches = 0;
  bool timing = false;

  bool init(int dev) {
    device = dev;
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (!d_lut.ensure(65536 * 4)) return false;
    std::vector<uint32_t> lut(65536);
    lut[0] = 0;
    for (uint32_t i = 1; i < 65536; ++i) lut[i] = (uint32_t)llround(std::log2((double)i) * 65536.0);
    CUDA_OK(cudaMemcpy(d_lut.p, lut.data(), 65536 * 4, cudaMemcpyHostToDevice));
    if (!d_total.ensure(64)) return false;
    CUDA_OK(cudaDeviceSetLimit(cudaLimitStackSize, 4096));
    ok = true;
    return true;
  }
  void destroy() {
    cudaSetDevice(device);
    DevBuf* all[] = {&d_data, &d_lut, &d_sortA, &d_sortB, &d_hist, &d_digit, &d_best, &d_raw, &d_unit, &d_cmds, &d_cmd_bits,
                     &d_lit_syms, &d_cmd_syms, &d_dist_syms, &d_mb, &d_split_u8, &d_split_u32, &d_split_counts, &d_hist_lit,
                     &d_hist_cmd, &d_hist_dist, &d_split_codes, &d_codes_u8, &d_codes_u16, &d_hdr, &d_huff_ws, &d_ctxmap_ws,
                     &d_out, &d_total, &d_tree_ws, &d_tree_bits, &d_tree_nbits, &d_cmd_tile, &d_long_tab, &d_seg_bits};
    for (auto* b : all) b->release();
    if (h_pinned) cudaFreeHost(h_pinned);
    for (auto& e : ev_pool) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }

  void fill_params(EncParams* P, int quality, int lgwin, uint64_t size_hint) const {
    memset(P, 0, sizeof(*P));
    if (quality < 5) quality = 5;  // the device path implements the hash-chain family q5..q9
    if (quality > 9) quality = 9;
    if (lgwin < 10) lgwin = 10;
    if (lgwin > 24) lgwin = 24;
    P->quality = quality;
    P->lgwin = lgwin;
    uint32_t hint = size_hint > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)size_hint;
    P->size_hint = hint;
    // ChooseHasher, encode.rs:834-893 (H40-42 are not implemented there and fall back to H6 with default params)
    if (quality == 9) { P->hash_type = 9; P->key_bits = 15; P->hash_len = 4; P->depth = 256; P->n_last = 16; }
    else if (lgwin <= 16) { P->hash_type = 6; P->key_bits = 15; P->hash_len = 5; P->depth = 256; P->n_last = 16; }
    else if (hint > (1u << 22) && lgwin >= 19) {
      P->hash_type = 6; P->key_bits = 15; P->hash_len = 5; P->depth = 1 << (quality - 1);
      P->n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
    } else {
      P->hash_type = 5; P->key_bits = (quality < 7 && hint <= (1u << 20)) ? 14 : 15; P->hash_len = 4;
      P->depth = 1 << (quality - 1);
      P->n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
    }
    P->lcap = lcap;
    P->unit = unit;
    P->mb_units = mb_units;
    P->max_backward = (1u << lgwin) - 16;
    P->use_rle_opt = use_rle_opt;
    P->split = split;
    P->ctx_model = ctx_model;
  }

  // device buffers for one chunk of `c` bytes
  bool ensure_chunk(uint32_t c, const EncParams& P, Workspace* W) {
    const uint32_t mb_span = P.unit * P.mb_units;
    const uint32_t NU = (c + P.unit - 1) / P.unit;
    const uint32_t NM = (NU + P.mb_units - 1) / P.mb_units;
    const uint32_t cu = P.unit / 2 + 1;
    const uint32_t cmd_cap = mb_span / 2 + 2;
    W->num_units = NU;
    W->num_mb = NM;
    W->cmd_cap = cmd_cap;
    W->lit_blk_cap = mb_span / 512 + 2;
    W->cmd_blk_cap = cmd_cap / 1024 + 2;
    W->dist_blk_cap = cmd_cap / 512 + 2;
    W->max_lit_trees = P.split ? 256 : 13;
    W->max_cmd_types = P.split ? 256 : 1;
    W->max_dist_types = P.split ? 256 : 1;
    W->hdr_cap = P.split ? (384u << 10) : (16u << 10);
    if (!d_best.ensure(((size_t)c + 64) * 4)) return false;
    if (!d_raw.ensure((size_t)NU * cu * sizeof(RawCmd))) return false;
    if (!d_unit.ensure((size_t)NU * 7 * 4)) return false;
    if (!d_cmds.ensure((size_t)NM * cmd_cap * sizeof(GCmd))) return false;
    if (!d_cmd_bits.ensure((size_t)NM * cmd_cap * 4)) return false;
    W->tile_cap = cmd_cap / 256 + 2;
    if (!d_cmd_tile.ensure((size_t)NM * W->tile_cap * 4)) return false;
    W->long_cap = mb_span / LONG_INS + 1;
    if (!d_long_tab.ensure((size_t)NM * W->long_cap * 8) || !d_seg_bits.ensure((size_t)NM * W->long_cap * 4)) return false;
    if (!d_lit_syms.ensure(((size_t)c + 64) * 2)) return false;
    if (!d_cmd_syms.ensure((size_t)NM * cmd_cap * 2)) return false;
    if (!d_dist_syms.ensure((size_t)NM * cmd_cap * 2)) return false;
    if (!d_mb.ensure((size_t)NM * sizeof(MBDesc))) return false;
    const size_t blk_total = (size_t)W->lit_blk_cap + W->cmd_blk_cap + W->dist_blk_cap;
    if (!d_split_u8.ensure((size_t)NM * blk_total)) return false;
    if (!d_split_u32.ensure((size_t)NM * blk_total * 2 * 4)) return false;
    if (!d_split_counts.ensure((size_t)NM * 6 * 4)) return false;
    if (!d_hist_lit.ensure((size_t)NM * (W->max_lit_trees + 13) * 256 * 4)) return false;
    if (!d_hist_cmd.ensure((size_t)NM * (W->max_cmd_types + 1) * 704 * 4)) return false;
    if (!d_hist_dist.ensure((size_t)NM * (W->max_dist_types + 1) * 64 * 4)) return false;
    if (!d_split_codes.ensure((size_t)NM * 3 * sizeof(SplitCode))) return false;
    const size_t code_syms = (size_t)W->max_lit_trees * 256 + (size_t)W->max_cmd_types * 704 + (size_t)W->max_dist_types * 64;
    if (!d_codes_u8.ensure((size_t)NM * code_syms)) return false;
    if (!d_codes_u16.ensure((size_t)NM * code_syms * 2)) return false;
    if (!d_hdr.ensure((size_t)NM * W->hdr_cap)) return false;
    if (!d_huff_ws.ensure((size_t)NM * sizeof(HuffStoreWs))) return false;
    if (!d_ctxmap_ws.ensure((size_t)NM * 256 * 64 * 4)) return false;
    const size_t tree_cap = (size_t)W->max_lit_trees + W->max_cmd_types + W->max_dist_types;
    if (!d_tree_bits.ensure((size_t)NM * tree_cap * TREE_SLOT_BYTES)) return false;
    if (!d_tree_nbits.ensure((size_t)NM * tree_cap * 4)) return false;
    // sort scratch
    const uint32_t nb = std::min<uint64_t>((uint64_t)c + (1ull << P.lgwin), kBatchMax);
    const uint32_t tiles = (nb + SORT_TILE - 1) / SORT_TILE;
    if (!d_sortA.ensure((size_t)nb * 4 + 64)) return false;
    if (!d_sortB.ensure((size_t)nb * 4 + 64)) return false;
    if (!d_hist.ensure((size_t)256 * tiles * 4)) return false;
    if (!d_digit.ensure(512 * 4)) return false;
    // wire pointers
    W->lut = d_lut.as<uint32_t>();
    W->best = d_best.as<uint32_t>();
    W->raw = d_raw.as<RawCmd>();
    uint32_t* up = d_unit.as<uint32_t>();
    W->unit_ncmd = up; W->unit_tail = up + NU; W->unit_ncopy = up + 2 * (size_t)NU;
    W->unit_cmd_off = up + 3 * (size_t)NU; W->unit_lit_off = up + 4 * (size_t)NU; W->unit_ndist = up + 5 * (size_t)NU; W->unit_dist_off = up + 6 * (size_t)NU;
    W->cmds = d_cmds.as<GCmd>();
    W->cmd_bits = d_cmd_bits.as<uint32_t>();
    W->cmd_tile = d_cmd_tile.as<uint32_t>();
    W->long_tab = d_long_tab.as<uint2>();
    W->seg_bits = d_seg_bits.as<uint32_t>();
    W->lit_syms = d_lit_syms.as<uint16_t>();
    W->cmd_syms = d_cmd_syms.as<uint16_t>();
    W->dist_syms = d_dist_syms.as<uint16_t>();
    W->mb = d_mb.as<MBDesc>();
    uint8_t* t8 = d_split_u8.as<uint8_t>();
    W->lit_types = t8; W->cmd_types = t8 + (size_t)NM * W->lit_blk_cap; W->dist_types = W->cmd_types + (size_t)NM * W->cmd_blk_cap;
    uint32_t* t32 = d_split_u32.as<uint32_t>();
    W->lit_lengths = t32; t32 += (size_t)NM * W->lit_blk_cap;
    W->lit_starts = t32; t32 += (size_t)NM * W->lit_blk_cap;
    W->cmd_lengths = t32; t32 += (size_t)NM * W->cmd_blk_cap;
    W->cmd_starts = t32; t32 += (size_t)NM * W->cmd_blk_cap;
    W->dist_lengths = t32; t32 += (size_t)NM * W->dist_blk_cap;
    W->dist_starts = t32;
    W->split_counts = d_split_counts.as<uint32_t>();
    W->lit_hist = d_hist_lit.as<uint32_t>(); W->cmd_hist = d_hist_cmd.as<uint32_t>(); W->dist_hist = d_hist_dist.as<uint32_t>();
    W->split_codes = d_split_codes.as<SplitCode>();
    uint8_t* c8 = d_codes_u8.as<uint8_t>();
    W->lit_depth = c8; W->cmd_depth = c8 + (size_t)NM * W->max_lit_trees * 256;
    W->dist_depth = W->cmd_depth + (size_t)NM * W->max_cmd_types * 704;
    uint16_t* c16 = d_codes_u16.as<uint16_t>();
    W->lit_code = c16; W->cmd_code = c16 + (size_t)NM * W->max_lit_trees * 256;
    W->dist_code = W->cmd_code + (size_t)NM * W->max_cmd_types * 704;
    W->hdr = d_hdr.as<uint8_t>();
    W->huff_ws = d_huff_ws.as<HuffStoreWs>();
    W->ctxmap_ws = d_ctxmap_ws.as<uint32_t>();
    W->tree_ws = d_tree_ws.as<HuffStoreWs>();
    W->tree_bits = d_tree_bits.as<uint8_t>();
    W->tree_nbits = d_tree_nbits.as<uint32_t>();
    W->total_bits = d_total.as<uint64_t>();
    return true;
  }

  void mark(int stage) {
    if (!timing) return;
    size_t i = mark_stage.size();
    if (i >= ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    cudaEventRecord(ev_pool[i], stream);
    mark_stage.push_back(stage);
  }
  void collect_timings() {
    for (int i = 0; i < B200_NUM_STAGES; ++i) stage_ms[i] = 0;
    for (size_t i = 0; i + 1 < mark_stage.size(); ++i) {
      if (mark_stage[i] < 0) continue;
      float ms = 0;
      cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]);
      stage_ms[mark_stage[i]] += ms;
    }
    mark_stage.clear();
  }

  // Compresses data[range_start, range_start + range_len) of the stream resident at d_data (absolute positions) and
  // appends its metablocks to W.out at *W.total_bits.  first/last control the stream header / trailer.
  bool run_chunk(const EncParams& Pstream, uint32_t range_start, uint32_t range_len, uint32_t* d_outw,
                 uint64_t out_cap_bytes, bool first, bool last, bool byte_align_end) {
    Workspace W;
    memset(&W, 0, sizeof(W));
    EncParams P = Pstream;
    P.n = range_len;
    P.abs_base = range_start;
    if (!ensure_chunk(range_len, P, &W)) return false;
    W.P = P;
    W.data = d_data.as<uint8_t>() + (range_start - data_base);
    W.out = d_outw;
    W.out_cap_bytes = out_cap_bytes;
    const uint8_t* d_all = d_data.as<uint8_t>() - data_base;  // indexable by absolute position >= data_base
    // metablock descriptors
    {
      std::vector<MBDesc> mbs(W.num_mb);
      for (uint32_t m = 0; m < W.num_mb; ++m) {
        MBDesc& d = mbs[m];
        memset(&d, 0, sizeof(d));
        d.u0 = m * P.mb_units;
        d.u1 = std::min(W.num_units, d.u0 + P.mb_units);
        d.start = d.u0 * P.unit;
        d.len = std::min<uint64_t>(range_len, (uint64_t)d.u1 * P.unit) - d.start;
      }
      CUDA_OK(cudaMemcpyAsync(W.mb, mbs.data(), mbs.size() * sizeof(MBDesc), cudaMemcpyHostToDevice, stream));
      CUDA_OK(cudaStreamSynchronize(stream));  // mbs is a stack-lifetime vector
    }
    // ---- sort + match, batch by batch ----
    const uint32_t window = 1u << P.lgwin;
    const uint32_t payload_max = kBatchMax - window - 4096;
    for (uint64_t b0 = range_start; b0 < (uint64_t)range_start + range_len; b0 += payload_max) {
      const uint32_t b1 = (uint32_t)std::min<uint64_t>((uint64_t)range_start + range_len, b0 + payload_max);
      uint32_t origin = b0 > window ? (uint32_t)b0 - window : 0u;
      origin &= ~4095u;  // tile staging needs word alignment
      if (origin < data_base) origin = (uint32_t)data_base;
      const uint32_t count = b1 - origin;
      const uint32_t tiles = (count + SORT_TILE - 1) / SORT_TILE;
      mark(B200_ST_SORT);
      SortArgs sa;
      sa.data = d_all + origin;
      sa.count = count;
      sa.hist = d_hist.as<uint32_t>();
      sa.digit_base = d_digit.as<uint32_t>();
      sa.num_tiles = tiles;
      sa.hash_type = P.hash_type;
      sa.key_bits = P.key_bits;
      for (int pass = 0; pass < 2; ++pass) {
        sa.pass = pass;
        sa.in = pass == 0 ? nullptr : d_sortA.as<uint32_t>();
        sa.outw = pass == 0 ? d_sortA.as<uint32_t>() : d_sortB.as<uint32_t>();
        k_sort_hist<<<tiles, SORT_THREADS, 0, stream>>>(sa);
        k_scan_rows<<<256, 256, 0, stream>>>(sa.hist, tiles, d_digit.as<uint32_t>() + 256);
        k_scan_digits<<<1, 256, 0, stream>>>(d_digit.as<uint32_t>() + 256, d_digit.as<uint32_t>());
        k_sort_scatter<<<tiles, SORT_THREADS, 0, stream>>>(sa);
        launches += 4;
      }
      MatchArgs ma;
      ma.data = d_all;
      ma.sorted = d_sortB.as<uint32_t>();
      ma.count = count;
      ma.origin = origin;
      ma.payload_begin = (uint32_t)b0 - origin;
      ma.n = range_start + range_len;  // matches may not run past the end of this range
      ma.best = W.best - range_start;  // best[] is indexed by range-relative position
      ma.hash_type = P.hash_type;
      ma.key_bits = P.key_bits;
      ma.depth = P.depth;
      ma.lcap = P.lcap;
      ma.max_backward = P.max_backward;
      const size_t smem = (size_t)(MATCH_THREADS + P.depth) * 6 * 4;
      mark(B200_ST_MATCH);
      k_match<<<(count + MATCH_THREADS - 1) / MATCH_THREADS, MATCH_THREADS, smem, stream>>>(ma);
      launches += 1;
    }
    mark(B200_ST_PARSE);
    k_parse<<<(W.num_units + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, stream>>>(W);
    mark(B200_ST_FINALIZE);
    k_fin_count<<<W.num_mb, 1024, 0, stream>>>(W);
    k_fin_write<<<(W.num_units + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, stream>>>(W);
    k_fin_dist<<<W.num_mb, 1024, 0, stream>>>(W);
    k_ctx_decide<<<W.num_mb, 256, 0, stream>>>(W);
    {
      dim3 g((W.cmd_cap + 255) / 256, W.num_mb);
      cudaMemsetAsync(W.long_tab, 0, (size_t)W.num_mb * W.long_cap * sizeof(uint2), stream);
      k_symbols<<<g, 256, 0, stream>>>(W);
      k_symbols_long<<<dim3(LONG_GRID, W.num_mb), 256, 0, stream>>>(W);
    }
    mark(B200_ST_SPLIT);
    {
      dim3 g(W.num_mb, 3);
      if (P.split) k_split_greedy<<<g, SPLIT_THREADS, 0, stream>>>(W);
      else k_split_simple<<<g, 512, 0, stream>>>(W);
    }
    mark(B200_ST_HEADER);
    {
      dim3 g(W.max_lit_trees + W.max_cmd_types + W.max_dist_types, W.num_mb);
      k_trees<<<g, 32, 0, stream>>>(W);
    }
    k_header<<<W.num_mb, 32, 0, stream>>>(W);
    mark(B200_ST_EMIT);
    {
      dim3 g((W.cmd_cap + 255) / 256, W.num_mb);
      k_bitlen_long<<<dim3(LONG_GRID, W.num_mb), 256, 0, stream>>>(W);
      k_bitlen<<<g, 256, 0, stream>>>(W);
      k_bitscan<<<W.num_mb, 1024, 0, stream>>>(W);
      k_layout<<<1, 32, 0, stream>>>(W, first ? 1 : 0, last ? 1 : 0, byte_align_end ? 1 : 0);
      k_emit_header<<<W.num_mb, 256, 0, stream>>>(W);
      k_emit_body<<<g, 256, 0, stream>>>(W);
      k_emit_long<<<dim3(LONG_GRID, W.num_mb), 256, 0, stream>>>(W);
      dim3 gr(64, W.num_mb);
      k_emit_raw<<<gr, 256, 0, stream>>>(W);
    }
    mark(-1);
    launches += 18;
    CUDA_OK(cudaGetLastError());
    return true;
  }

  // Whole-stream compression of n bytes already resident at d_data[0..n) (padded).  Output to d_outw.
  bool compress_resident(int quality, int lgwin, uint64_t size_hint, size_t n, uint32_t* d_outw, uint64_t out_cap_bytes,
                         bool first, bool last, bool byte_align_end, uint32_t range_start, uint32_t range_len) {
    EncParams P;
    fill_params(&P, quality, lgwin, size_hint ? size_hint : n);
    (void)n;
    for (uint64_t s = range_start; s < (uint64_t)range_start + range_len; s += kChunk) {
      uint32_t len = (uint32_t)std::min<uint64_t>(kChunk, (uint64_t)range_start + range_len - s);
      bool f = first && s == range_start;
      bool l = s + len == (uint64_t)range_start + range_len;
      if (!run_chunk(P, (uint32_t)s, len, d_outw, out_cap_bytes, f, last && l, byte_align_end && l)) return false;
    }
    return true;
  }
};

// k_layout takes flags; declared here because it needs the final signature
namespace bro {}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

B200Encoder* b200_encoder_create(int device) {
  B200Encoder* e = new B200Encoder();
  if (!e->init(device)) {
    delete e;
    return nullptr;
  }
  return e;
}
void b200_encoder_destroy(B200Encoder* e) {
  if (!e) return;
  e->destroy();
  delete e;
}
int b200_encoder_set_option(B200Encoder* e, int option, uint32_t value) {
  if (!e) return 0;
  switch (option) {
    case B200_OPT_UNIT: e->unit = value; return 1;
    case B200_OPT_MB_UNITS: e->mb_units = value; return 1;
    case B200_OPT_LCAP: e->lcap = value > 255 ? 255 : value; return 1;
    case B200_OPT_RLE_OPT: e->use_rle_opt = (int)value; return 1;
    case B200_OPT_SPLIT: e->split = (int)value; return 1;
    case B200_OPT_CTX_MODEL: e->ctx_model = (int)value; return 1;
    case B200_OPT_TIMING: e->timing = value != 0; return 1;
  }
  return 0;
}

size_t b200_max_compressed_size(size_t n) { return n + (n >> 10) * 8 + 4096; }

// Stage the input on the device: from device memory (kind = 1) or host memory (kind = 0).
// Only the bytes a range can see are staged: [base, end) with base = 4 KiB-aligned start of its window halo.
static bool stage_input(B200Encoder* e, const uint8_t* in, size_t base, size_t end, int kind) {
  const size_t n = end - base;
  if (!e->d_data.ensure(n + kPad)) return false;
  e->data_base = base;
  CUDA_OK(cudaMemcpyAsync(e->d_data.p, in + base, n, kind == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
  CUDA_OK(cudaMemsetAsync(e->d_data.as<uint8_t>() + n, 0, kPad, e->stream));
  return true;
}

// Compresses [range_start, range_start+range_len) of an n-byte stream.  in/out are device pointers when
// device_io != 0, host pointers otherwise.  first/last: emit stream header / final empty metablock;
// byte_align: end the range with a padding metablock so that ranges can be concatenated with memcpy.
int b200_encoder_compress_range(B200Encoder* e, int quality, int lgwin, uint64_t size_hint, const uint8_t* in, size_t n,
                                size_t range_start, size_t range_len, int first, int last, int byte_align, uint8_t* out,
                                size_t out_cap, size_t* out_size, int device_io) {
  if (!e || !e->ok || !out_size) return 0;
  if (n >= 0xFFFFF000ull) return 0;  // 32-bit positions
  if (cudaSetDevice(e->device) != cudaSuccess) return 0;
  if (n == 0 || range_len == 0) {
    if (first && last && n == 0) {  // encode.rs:1463-1467
      if (out_cap < 1) return 0;
      uint8_t b = 6;
      if (device_io) { if (cudaMemcpy(out, &b, 1, cudaMemcpyHostToDevice) != cudaSuccess) return 0; }
      else out[0] = b;
      *out_size = 1;
      return 1;
    }
    *out_size = 0;
    return 1;
  }
  e->launches = 0;
  e->mark_stage.clear();
  const size_t need = b200_max_compressed_size(range_len) + 64;
  {
    int lw = lgwin < 10 ? 10 : (lgwin > 24 ? 24 : lgwin);
    size_t window = (size_t)1 << lw;
    size_t base = range_start > window ? ((range_start - window) & ~(size_t)4095) : 0;
    if (!stage_input(e, in, base, range_start + range_len, device_io)) return 0;
  }
  if (!e->d_out.ensure(need)) return 0;
  if (cudaMemsetAsync(e->d_out.p, 0, need, e->stream) != cudaSuccess) return 0;
  if (cudaMemsetAsync(e->d_total.p, 0, 8, e->stream) != cudaSuccess) return 0;
  if (!e->compress_resident(quality, lgwin, size_hint, n, e->d_out.as<uint32_t>(), need, first != 0, last != 0,
                            byte_align != 0, (uint32_t)range_start, (uint32_t)range_len))
    return 0;
  uint64_t total_bits = 0;
  if (cudaMemcpyAsync(&total_bits, e->d_total.p, 8, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) return 0;
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
    fprintf(stderr, "[brotli_b200] kernel failure: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  size_t bytes = (size_t)((total_bits + 7) >> 3);
  if (bytes > out_cap) return 0;
  if (cudaMemcpy(out, e->d_out.p, bytes, device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  *out_size = bytes;
  if (e->timing) e->collect_timings();
  return 1;
}

int b200_encoder_compress(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint8_t* out, size_t out_cap,
                          size_t* out_size, int device_io) {
  return b200_encoder_compress_range(e, quality, lgwin, n, in, n, 0, n, 1, 1, 0, out, out_cap, out_size, device_io);
}

int b200_encoder_last_timings(B200Encoder* e, float* ms, uint32_t* launches) {
  if (!e) return 0;
  for (int i = 0; i < B200_NUM_STAGES; ++i) ms[i] = e->stage_ms[i];
  if (launches) *launches = e->launches;
  return 1;
}

// test hook: device results of the match stage for an n-byte buffer (host in, host out)
int b200_stage_match(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint32_t* best_out) {
  if (!e || !e->ok || n == 0 || n > kChunk) return 0;
  if (cudaSetDevice(e->device) != cudaSuccess) return 0;
  if (!stage_input(e, in, 0, n, 0)) return 0;
  const size_t need = b200_max_compressed_size(n) + 64;
  if (!e->d_out.ensure(need)) return 0;
  cudaMemsetAsync(e->d_out.p, 0, need, e->stream);
  cudaMemsetAsync(e->d_total.p, 0, 8, e->stream);
  if (!e->compress_resident(quality, lgwin, n, n, e->d_out.as<uint32_t>(), need, true, true, false, 0, (uint32_t)n)) return 0;
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return 0;
  return cudaMemcpy(best_out, e->d_best.p, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
}

}  // extern "C"
