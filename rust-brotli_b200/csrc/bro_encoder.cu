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
#include "bro_kernels_hq.cuh"
#include "bro_encoder.h"
#include "bro_dict_data.inc"  // generated at build time by gen_dict.py: kDictData, kDictHash

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
constexpr uint32_t kChunk = BRO_CHUNK_BYTES;   // bytes per pipeline pass (one sort batch for lgwin <= 22)
constexpr uint32_t kBatchMax = 1u << 25;       // positions per sort batch (25-bit packed positions)
constexpr uint32_t kLookahead = 4096;          // input bytes past a chunk's end that must be resident before it runs
constexpr int kMaxLanes = 6;

struct EventPool {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  cudaEvent_t get(bool timing) {
    (void)timing;
    if (used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming);
      ev.push_back(e);
    }
    return ev[used++];
  }
  void reset() { used = 0; }
  void destroy() { for (auto& e : ev) cudaEventDestroy(e); ev.clear(); used = 0; }
};

// One lane = one stream + the per-chunk workspace.  Consecutive chunks alternate between lanes so that the
// latency-bound stages of one chunk (block splitting, Huffman trees, layout) overlap the throughput-bound stages of
// the next one (sort, match, parse) and the host<->device copies.
struct Lane {
  cudaStream_t stream = nullptr;
  DevBuf d_sortA, d_sortB, d_hist, d_digit, d_best, d_raw, d_unit, d_cmds, d_cmd_bits, d_lit_syms, d_cmd_syms, d_dist_syms,
      d_hqm, d_hqn, d_hq_nodes, d_hq_pre, d_hq_scratch, d_bs_meta, d_bs_blockid, d_bs_signal, d_bs_hist, d_bs_icost, d_bs_first, d_bs_fmap, d_bs_bstart,
      d_bs_bh_in, d_bs_bh_work, d_bs_u64, d_bs_u32, d_bs_nsurv, d_cm_in, d_cm_work, d_cm_u64, d_cm_u32, d_cm_nsurv, d_cm_counts, d_cm_maps, d_dist_cost, d_mb, d_split_u8, d_split_u32, d_split_counts, d_hist_lit, d_hist_cmd, d_hist_dist, d_split_codes, d_codes_u8,
      d_codes_u16, d_hdr, d_huff_ws, d_ctxmap_ws, d_tree_ws, d_tree_bits, d_tree_nbits, d_cmd_tile, d_long_tab, d_seg_bits, d_sect_bits, d_sect_nbits;
  EventPool marks;  // timing marks: (event, stage that starts there); -1 ends the last stage
  std::vector<int> mark_stage;
  void release() {
    DevBuf* all[] = {&d_hqm, &d_hqn, &d_hq_nodes, &d_hq_pre, &d_hq_scratch, &d_bs_meta, &d_bs_blockid, &d_bs_signal, &d_bs_hist, &d_bs_icost,
                     &d_bs_first, &d_bs_fmap, &d_bs_bstart, &d_bs_bh_in, &d_bs_bh_work, &d_bs_u64, &d_bs_u32, &d_bs_nsurv, &d_cm_in, &d_cm_work, &d_cm_u64,
                     &d_cm_u32, &d_cm_nsurv, &d_cm_counts, &d_cm_maps, &d_dist_cost, &d_sortA, &d_sortB, &d_hist, &d_digit, &d_best, &d_raw, &d_unit, &d_cmds, &d_cmd_bits, &d_lit_syms,
                     &d_cmd_syms, &d_dist_syms, &d_mb, &d_split_u8, &d_split_u32, &d_split_counts, &d_hist_lit, &d_hist_cmd,
                     &d_hist_dist, &d_split_codes, &d_codes_u8, &d_codes_u16, &d_hdr, &d_huff_ws, &d_ctxmap_ws, &d_tree_ws,
                     &d_tree_bits, &d_tree_nbits, &d_cmd_tile, &d_long_tab, &d_seg_bits, &d_sect_bits, &d_sect_nbits};
    for (auto* b : all) b->release();
    marks.destroy();
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
};

}  // namespace

struct B200Encoder {
  int device = 0;
  bool ok = false;
  // configuration knobs (tests flip these)
  uint32_t unit = 4096, mb_units = 1024, lcap = 64;
  int use_rle_opt = 1, split = 1, ctx_model = 1, use_dict = 1, hq_split = 1, hq_levels = HQ_MAX_LEVELS;
  uint32_t hq_unit = 0;  // parse unit of the shortest-path parse (quality >= 10); 0 = 8 KiB at q10, 16 KiB at q11 (measured: DESIGN.md)
  int hq_thread_units = 0;   // 1: one parse unit per thread instead of one per warp (A/B switch)
  int num_lanes = 4;
  int ondemand = 1;       // q7..q9: search deep buckets where the parse stands (1) or for every position up front (0, A/B)
  int pair_parse = 4;     // parse units per warp for q5 / q6: 4 (default) or 2; 0 = one unit per warp (kept for A/B measurements)
  int shallow_match = 1;  // (the loop version of the depth 16 / 32 scan is gone; the option is accepted and ignored)
  Lane lanes[kMaxLanes];
  cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams
  DevBuf d_dict_words, d_dict_hash, d_dict_lutb, d_dict_lute, d_dict_trg, d_dict_tr;
  DevBuf d_data, d_lut, d_out, d_total;          // d_total: [0] running bit position, [1 + k] position after chunk k
  uint64_t* h_total = nullptr;                   // pinned mirror of d_total[1 + k]
  size_t h_total_cap = 0;
  EventPool sync_events;
  uint64_t data_base = 0;  // absolute stream position of d_data[0]
  float stage_ms[B200_NUM_STAGES];
  uint32_t launches = 0;
  bool timing = false;

  bool init(int dev) {
    device = dev;
    CUDA_OK(cudaSetDevice(device));
    // (descending stream priorities per lane, to stagger the chunks, were measured: 12.1 ms vs 11.4 ms with equal priority)
    for (auto& L : lanes) CUDA_OK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    if (!d_lut.ensure(65536 * 4)) return false;
    std::vector<uint32_t> lut(65536);
    lut[0] = 0;
    for (uint32_t i = 1; i < 65536; ++i) lut[i] = (uint32_t)llround(std::log2((double)i) * 65536.0);
    CUDA_OK(cudaMemcpy(d_lut.p, lut.data(), 65536 * 4, cudaMemcpyHostToDevice));
    if (!d_dict_words.ensure(sizeof(kDictData) + 64) || !d_dict_hash.ensure(sizeof(kDictHash))) return false;
    CUDA_OK(cudaMemset(d_dict_words.p, 0, sizeof(kDictData) + 64));
    CUDA_OK(cudaMemcpy(d_dict_words.p, kDictData, sizeof(kDictData), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dict_hash.p, kDictHash, sizeof(kDictHash), cudaMemcpyHostToDevice));
    if (!d_dict_lutb.ensure(sizeof(kDictLutBuckets)) || !d_dict_lute.ensure(sizeof(kDictLutEntries)) || !d_dict_trg.ensure(sizeof(kDictTrGroups)) ||
        !d_dict_tr.ensure(sizeof(kDictTransforms)))
      return false;
    CUDA_OK(cudaMemcpy(d_dict_lutb.p, kDictLutBuckets, sizeof(kDictLutBuckets), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dict_lute.p, kDictLutEntries, sizeof(kDictLutEntries), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dict_trg.p, kDictTrGroups, sizeof(kDictTrGroups), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dict_tr.p, kDictTransforms, sizeof(kDictTransforms), cudaMemcpyHostToDevice));
    CUDA_OK(cudaDeviceSetLimit(cudaLimitStackSize, 4096));
    CUDA_OK(cudaFuncSetAttribute(k_split_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLIT_SMEM_WORDS * 4));
    for (int i = 0; i < B200_NUM_STAGES; ++i) stage_ms[i] = 0;
    ok = true;
    return true;
  }
  void destroy() {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    for (auto& L : lanes) L.release();
    DevBuf* all[] = {&d_data, &d_lut, &d_out, &d_total, &d_dict_words, &d_dict_hash, &d_dict_lutb, &d_dict_lute, &d_dict_trg, &d_dict_tr};
    for (auto* b : all) b->release();
    if (h_total) cudaFreeHost(h_total);
    sync_events.destroy();
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
  }
  bool ensure_totals(size_t chunks) {
    if (!d_total.ensure((chunks + 2) * 8)) return false;
    if (chunks + 2 > h_total_cap) {
      if (h_total) cudaFreeHost(h_total);
      h_total = nullptr;
      h_total_cap = chunks + 64;
      CUDA_OK(cudaHostAlloc((void**)&h_total, h_total_cap * 8, cudaHostAllocDefault));
    }
    return true;
  }

  void fill_params(EncParams* P, int quality, int lgwin, uint64_t size_hint) const {
    memset(P, 0, sizeof(*P));
    quality = b200_effective_quality(quality);
    if (lgwin < 10) lgwin = 10;
    if (lgwin > 24) lgwin = 24;
    P->quality = quality;
    P->lgwin = lgwin;
    uint32_t hint = size_hint > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)size_hint;
    P->size_hint = hint;
    // ChooseHasher, encode.rs:834-893 (H40-42 are not implemented there and fall back to H6 with default params)
    if (quality >= 10) { P->hash_type = 5; P->key_bits = 15; P->hash_len = 4; P->depth = 256; P->n_last = 16; }  // bucket lists for k_match_all (with the long-prefix levels on, 64..1024 give the same size +-0.02 %)
    else if (quality == 9) { P->hash_type = 9; P->key_bits = 15; P->hash_len = 4; P->depth = 256; P->n_last = 16; }
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
    P->use_dict = use_dict;
    P->hq_split = hq_split;
    P->hq_levels = quality >= 10 ? hq_levels : 0;
    P->hq_warm = 1;
    if (quality >= 10) {  // same metablock span, larger parse units
      const uint32_t span = P->unit * P->mb_units;
      P->unit = bmin(hq_unit ? hq_unit : hq_default_unit(quality, hint), span);
      P->mb_units = span / P->unit;
      P->lcap = HQ_LCAP;
    }
  }

  // device buffers of lane L for one chunk of `c` bytes
  bool ensure_chunk(Lane& L, uint32_t c, const EncParams& P, Workspace* W) {
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
    W->dist_A = (P.quality >= 10 && P.hq_split) ? BRO_DIST_A_MAX : 64u;
    W->hdr_cap = P.split ? (384u << 10) : (16u << 10);
    const bool hq = P.quality >= 10;
    if (!hq && !L.d_best.ensure(((size_t)c + 64) * 4)) return false;
    if (hq) {
      if (!L.d_hqm.ensure(((size_t)c + 64) * HQ_MAXM * sizeof(HqMatch)) || !L.d_hqn.ensure((size_t)c + 64)) return false;
      if (!L.d_hq_nodes.ensure((size_t)NU * (P.unit + 1) * sizeof(ZNode)) || !L.d_hq_pre.ensure((size_t)NU * (P.unit + 1) * 4)) return false;
      if (!L.d_hq_scratch.ensure((size_t)NU * HQ_SCRATCH_WORDS * 4)) return false;
    }
    if (!L.d_raw.ensure((size_t)NU * cu * sizeof(RawCmd))) return false;
    if (!L.d_unit.ensure((size_t)NU * 7 * 4)) return false;
    if (!L.d_cmds.ensure((size_t)NM * cmd_cap * sizeof(GCmd))) return false;
    if (!L.d_cmd_bits.ensure((size_t)NM * cmd_cap * 4)) return false;
    W->tile_cap = cmd_cap / 256 + 2;
    if (!L.d_cmd_tile.ensure((size_t)NM * W->tile_cap * 4)) return false;
    W->long_cap = mb_span / LONG_INS + 1;
    if (!L.d_long_tab.ensure((size_t)NM * W->long_cap * 8) || !L.d_seg_bits.ensure((size_t)NM * W->long_cap * 4)) return false;
    if (!L.d_lit_syms.ensure(((size_t)c + 64) * 2)) return false;
    if (!L.d_cmd_syms.ensure((size_t)NM * cmd_cap * 2)) return false;
    if (!L.d_dist_syms.ensure((size_t)NM * cmd_cap * 2)) return false;
    if (!L.d_mb.ensure((size_t)NM * sizeof(MBDesc))) return false;
    const size_t blk_total = (size_t)W->lit_blk_cap + W->cmd_blk_cap + W->dist_blk_cap;
    if (!L.d_split_u8.ensure((size_t)NM * blk_total)) return false;
    if (!L.d_split_u32.ensure((size_t)NM * blk_total * 2 * 4)) return false;
    if (!L.d_split_counts.ensure((size_t)NM * 6 * 4)) return false;
    if (!L.d_hist_lit.ensure((size_t)NM * (W->max_lit_trees + 13) * 256 * 4)) return false;
    if (!L.d_hist_cmd.ensure((size_t)NM * (W->max_cmd_types + 1) * 704 * 4)) return false;
    if (!L.d_hist_dist.ensure((size_t)NM * (W->max_dist_types + 1) * W->dist_A * 4)) return false;
    if (!L.d_split_codes.ensure((size_t)NM * 3 * sizeof(SplitCode))) return false;
    const size_t code_syms = (size_t)W->max_lit_trees * 256 + (size_t)W->max_cmd_types * 704 + (size_t)W->max_dist_types * W->dist_A;
    if (!L.d_codes_u8.ensure((size_t)NM * code_syms)) return false;
    if (!L.d_codes_u16.ensure((size_t)NM * code_syms * 2)) return false;
    if (!L.d_hdr.ensure((size_t)NM * W->hdr_cap)) return false;
    if (!L.d_huff_ws.ensure((size_t)NM * sizeof(HuffStoreWs))) return false;
    if (!L.d_ctxmap_ws.ensure((size_t)NM * (256 * 64 + 1024) * 4)) return false;
    if (!L.d_cm_maps.ensure((size_t)NM * (CM_LIT_MAX + CM_DIST_MAX)) || !L.d_cm_counts.ensure((size_t)NM * 2 * 4)) return false;
    W->lit_cmap = L.d_cm_maps.as<uint8_t>();
    W->dist_cmap = W->lit_cmap + (size_t)NM * CM_LIT_MAX;
    W->cm_counts = L.d_cm_counts.as<uint32_t>();
    const size_t tree_cap = (size_t)W->max_lit_trees + W->max_cmd_types + W->max_dist_types;
    if (!L.d_tree_bits.ensure((size_t)NM * tree_cap * TREE_SLOT_BYTES)) return false;
    if (!L.d_tree_nbits.ensure((size_t)NM * tree_cap * 4)) return false;
    if (!L.d_sect_bits.ensure((size_t)NM * HDR_SECTIONS * SECT_BYTES) || !L.d_sect_nbits.ensure((size_t)NM * HDR_SECTIONS * 4)) return false;
    // sort scratch
    const uint32_t nb = std::min<uint64_t>((uint64_t)c + (1ull << P.lgwin) + 4096, kBatchMax);
    const uint32_t tiles = (nb + SORT_TILE - 1) / SORT_TILE;
    if (!L.d_sortA.ensure((size_t)nb * 4 + 64)) return false;
    if (!L.d_sortB.ensure((size_t)nb * 4 + 64)) return false;
    if (!L.d_hist.ensure((size_t)256 * tiles * 4)) return false;
    if (!L.d_digit.ensure(512 * 4)) return false;
    // wire pointers
    W->lut = d_lut.as<uint32_t>();
    W->dict.words = d_dict_words.as<uint8_t>();
    W->dict.hash = d_dict_hash.as<uint16_t>();
    W->dict.lut_buckets = d_dict_lutb.as<uint16_t>();
    W->dict.lut_entries = d_dict_lute.as<uint32_t>();
    W->dict.tr_groups = d_dict_trg.as<uint8_t>();
    W->dict.transforms = d_dict_tr.as<uint8_t>();
    W->dict.num_tr_groups = BRO_DICT_NUM_TR_GROUPS;
    W->best = L.d_best.as<uint32_t>();
    W->hqm = L.d_hqm.as<HqMatch>();
    W->hqn = L.d_hqn.as<uint8_t>();
    W->raw = L.d_raw.as<RawCmd>();
    uint32_t* up = L.d_unit.as<uint32_t>();
    W->unit_ncmd = up; W->unit_tail = up + NU; W->unit_ncopy = up + 2 * (size_t)NU;
    W->unit_cmd_off = up + 3 * (size_t)NU; W->unit_lit_off = up + 4 * (size_t)NU; W->unit_ndist = up + 5 * (size_t)NU; W->unit_dist_off = up + 6 * (size_t)NU;
    W->cmds = L.d_cmds.as<GCmd>();
    W->cmd_bits = L.d_cmd_bits.as<uint32_t>();
    W->cmd_tile = L.d_cmd_tile.as<uint32_t>();
    W->long_tab = L.d_long_tab.as<uint2>();
    W->seg_bits = L.d_seg_bits.as<uint32_t>();
    W->lit_syms = L.d_lit_syms.as<uint16_t>();
    W->cmd_syms = L.d_cmd_syms.as<uint16_t>();
    W->dist_syms = L.d_dist_syms.as<uint16_t>();
    W->mb = L.d_mb.as<MBDesc>();
    uint8_t* t8 = L.d_split_u8.as<uint8_t>();
    W->lit_types = t8; W->cmd_types = t8 + (size_t)NM * W->lit_blk_cap; W->dist_types = W->cmd_types + (size_t)NM * W->cmd_blk_cap;
    uint32_t* t32 = L.d_split_u32.as<uint32_t>();
    W->lit_lengths = t32; t32 += (size_t)NM * W->lit_blk_cap;
    W->lit_starts = t32; t32 += (size_t)NM * W->lit_blk_cap;
    W->cmd_lengths = t32; t32 += (size_t)NM * W->cmd_blk_cap;
    W->cmd_starts = t32; t32 += (size_t)NM * W->cmd_blk_cap;
    W->dist_lengths = t32; t32 += (size_t)NM * W->dist_blk_cap;
    W->dist_starts = t32;
    W->split_counts = L.d_split_counts.as<uint32_t>();
    W->lit_hist = L.d_hist_lit.as<uint32_t>(); W->cmd_hist = L.d_hist_cmd.as<uint32_t>(); W->dist_hist = L.d_hist_dist.as<uint32_t>();
    W->split_codes = L.d_split_codes.as<SplitCode>();
    uint8_t* c8 = L.d_codes_u8.as<uint8_t>();
    W->lit_depth = c8; W->cmd_depth = c8 + (size_t)NM * W->max_lit_trees * 256;
    W->dist_depth = W->cmd_depth + (size_t)NM * W->max_cmd_types * 704;
    uint16_t* c16 = L.d_codes_u16.as<uint16_t>();
    W->lit_code = c16; W->cmd_code = c16 + (size_t)NM * W->max_lit_trees * 256;
    W->dist_code = W->cmd_code + (size_t)NM * W->max_cmd_types * 704;
    W->hdr = L.d_hdr.as<uint8_t>();
    W->huff_ws = L.d_huff_ws.as<HuffStoreWs>();
    W->ctxmap_ws = L.d_ctxmap_ws.as<uint32_t>();
    W->tree_ws = L.d_tree_ws.as<HuffStoreWs>();
    W->tree_bits = L.d_tree_bits.as<uint8_t>();
    W->tree_nbits = L.d_tree_nbits.as<uint32_t>();
    W->sect_bits = L.d_sect_bits.as<uint8_t>();
    W->sect_nbits = L.d_sect_nbits.as<uint32_t>();
    W->total_bits = d_total.as<uint64_t>();
    return true;
  }

  // workspaces of the quality >= 10 histogram stage (BrotliSplitBlock + context-map clustering) for one chunk
  bool ensure_hq_split(Lane& L, const Workspace& W, BsWs* B, CmWs* M) {
    const uint32_t NM = W.num_mb;
    const uint32_t mb_span = W.P.unit * W.P.mb_units;
    memset(B, 0, sizeof(*B));
    memset(M, 0, sizeof(*M));
    B->cap[0] = mb_span; B->cap[1] = W.cmd_cap; B->cap[2] = W.cmd_cap;
    B->maxb[0] = W.lit_blk_cap; B->maxb[1] = W.cmd_blk_cap; B->maxb[2] = W.dist_blk_cap;
    for (int i = 0; i < 3; ++i) B->segc[i] = B->cap[i] / BS_SEG + 1;
    B->cap_sum = B->cap[0] + B->cap[1] + B->cap[2];
    B->maxb_sum = B->maxb[0] + B->maxb[1] + B->maxb[2];
    B->segc_sum = B->segc[0] + B->segc[1] + B->segc[2];
    B->dist_A = W.dist_A;
    B->hist_stride = 100u * (256u + 704u + W.dist_A);
    B->bh_sum = B->maxb[0] * 256 + B->maxb[1] * 704 + B->maxb[2] * W.dist_A;
    B->nsurv_stride = std::max(B->maxb[0], std::max(B->maxb[1], B->maxb[2])) / 64 + 2;
    const size_t nb = (size_t)NM * B->maxb_sum;
    if (!L.d_bs_meta.ensure((size_t)NM * 3 * sizeof(BsMeta)) || !L.d_bs_blockid.ensure((size_t)NM * B->cap_sum + 64) ||
        !L.d_bs_signal.ensure((size_t)NM * B->cap_sum * 16 + 64) || !L.d_bs_hist.ensure((size_t)NM * B->hist_stride * 4) ||
        !L.d_bs_icost.ensure((size_t)NM * B->hist_stride * 4) || !L.d_bs_first.ensure((size_t)NM * 3 * 128 * 4) ||
        !L.d_bs_fmap.ensure((size_t)NM * B->segc_sum * 129) || !L.d_bs_bstart.ensure((nb + NM * 3) * 4 + 64) ||
        !L.d_bs_bh_in.ensure((size_t)NM * B->bh_sum * 4) || !L.d_bs_bh_work.ensure((size_t)NM * B->bh_sum * 4) || !L.d_bs_u64.ensure(nb * 2 * 8) ||
        !L.d_bs_u32.ensure(nb * 4 * 4) || !L.d_bs_nsurv.ensure((size_t)NM * 3 * B->nsurv_stride * 4))
      return false;
    B->meta = L.d_bs_meta.as<BsMeta>();
    B->blockid = L.d_bs_blockid.as<uint8_t>();
    B->signal = L.d_bs_signal.as<uint32_t>();
    B->hist = L.d_bs_hist.as<uint32_t>();
    B->icost = L.d_bs_icost.as<uint32_t>();
    B->firstpos = L.d_bs_first.as<uint32_t>();
    B->fmap = L.d_bs_fmap.as<uint8_t>();
    B->enter = B->fmap + (size_t)NM * B->segc_sum * 128;
    B->bstart = L.d_bs_bstart.as<uint32_t>();
    B->bh_in = L.d_bs_bh_in.as<uint32_t>();
    B->bh_work = L.d_bs_bh_work.as<uint32_t>();
    B->ccost = L.d_bs_u64.as<uint64_t>();
    B->bd = reinterpret_cast<int64_t*>(B->ccost + nb);
    B->csize = L.d_bs_u32.as<uint32_t>(); B->hsym = B->csize + nb; B->clusters = B->hsym + nb; B->bj = B->clusters + nb;
    B->nsurv = L.d_bs_nsurv.as<uint32_t>();
    const size_t nc = (size_t)NM * (CM_LIT_MAX + CM_DIST_MAX);
    const size_t hl = (size_t)NM * CM_LIT_MAX * 256, hd = (size_t)NM * CM_DIST_MAX * W.dist_A;
    if (!L.d_cm_in.ensure((hl + hd) * 4) || !L.d_cm_work.ensure((hl + hd) * 4) || !L.d_cm_u64.ensure(nc * 2 * 8) || !L.d_cm_u32.ensure(nc * 4 * 4) ||
        !L.d_cm_nsurv.ensure((size_t)NM * 2 * CM_NSURV_STRIDE * 4))
      return false;
    M->in_lit = L.d_cm_in.as<uint32_t>(); M->in_dist = M->in_lit + hl;
    M->work_lit = L.d_cm_work.as<uint32_t>(); M->work_dist = M->work_lit + hl;
    M->cost = L.d_cm_u64.as<uint64_t>();
    M->bd = reinterpret_cast<int64_t*>(M->cost + nc);
    M->size = L.d_cm_u32.as<uint32_t>(); M->sym = M->size + nc; M->clusters = M->sym + nc; M->bj = M->clusters + nc;
    M->nsurv = L.d_cm_nsurv.as<uint32_t>();
    M->counts = W.cm_counts;
    M->lit_cmap = W.lit_cmap;
    M->dist_cmap = W.dist_cmap;
    return true;
  }
  // BrotliSplitBlock + BrotliBuildMetaBlock's clustering for every metablock of the chunk (replaces k_split_greedy)
  bool run_hq_split(Lane& L, const Workspace& W) {
    BsWs B;
    CmWs M;
    if (!ensure_hq_split(L, W, &B, &M)) return false;
    cudaStream_t st = L.stream;
    const uint32_t NM = W.num_mb;
    const dim3 g3(NM, 3), gx3(64, NM, 3), g2(NM, 2), gx2(64, NM, 2);
    k_bs_setup<<<g3, 128, 0, st>>>(W, B);
    k_bs_sample<<<dim3(32, NM, 3), 256, 0, st>>>(W, B);
    for (int it = 0; it < 3; ++it) {
      k_bs_icost<<<g3, 256, 0, st>>>(W, B);
      k_bs_forward<<<dim3(B.segc[0], NM, 3), 32, 0, st>>>(W, B);
      k_bs_bfunc<<<dim3(B.segc[0], NM, 3), 32, 0, st>>>(W, B);
      k_bs_bchain<<<g3, 32, 0, st>>>(W, B);
      k_bs_bwrite<<<dim3(B.segc[0], NM, 3), 32, 0, st>>>(W, B);
      k_bs_remap<<<g3, 256, 0, st>>>(W, B);
      k_bs_rehist<<<gx3, 256, 0, st>>>(W, B);
    }
    k_bs_blocks<<<g3, 1024, 0, st>>>(W, B);
    CUDA_OK(cudaMemsetAsync(B.bh_in, 0, (size_t)NM * B.bh_sum * 4, st));
    k_bs_bhist<<<gx3, 256, 0, st>>>(W, B);
    k_bs_cl_prepare<<<gx3, CL_WARPS * 32, 0, st>>>(W, B);
    k_bs_cl_batch<<<dim3(128, NM, 3), CLB_WARPS * 32, 0, st>>>(W, B);
    k_bs_cl_final<<<g3, CLB_WARPS * 32, 0, st>>>(W, B);
    k_bs_cl_assign<<<gx3, CL_WARPS * 32, 0, st>>>(W, B);
    k_bs_types<<<g3, 32, 0, st>>>(W, B);
    k_cm_zero<<<dim3(64, NM), 256, 0, st>>>(W, M);
    k_cm_hist<<<gx3, 256, 0, st>>>(W, M);
    k_cm_cl_prepare<<<gx2, CL_WARPS * 32, 0, st>>>(W, M);
    k_cm_cl_batch<<<dim3(256, NM, 2), CLB_WARPS * 32, 0, st>>>(W, M);
    k_cm_cl_final<<<g2, CLB_WARPS * 32, 0, st>>>(W, M);
    k_cm_cl_assign<<<gx2, CL_WARPS * 32, 0, st>>>(W, M);
    k_cm_reindex<<<g2, 256, 0, st>>>(W, M);
    k_cm_rebuild<<<gx2, 256, 0, st>>>(W, M);
    launches += 2 + 21 + 3 + 5 + 8;
    return true;
  }

  void mark(Lane& L, int stage) {
    if (!timing) return;
    cudaEventRecord(L.marks.get(true), L.stream);
    L.mark_stage.push_back(stage);
  }
  void reset_timings() {
    for (auto& L : lanes) { L.marks.reset(); L.mark_stage.clear(); }
  }
  // per-stage sums of event-bracketed time on each lane's own stream (with two lanes stages of different chunks overlap,
  // so the sum over stages can exceed the wall time)
  void collect_timings() {
    for (int i = 0; i < B200_NUM_STAGES; ++i) stage_ms[i] = 0;
    for (auto& L : lanes) {
      for (size_t i = 0; i + 1 < L.mark_stage.size(); ++i) {
        if (L.mark_stage[i] < 0) continue;
        float ms = 0;
        cudaEventElapsedTime(&ms, L.marks.ev[i], L.marks.ev[i + 1]);
        stage_ms[L.mark_stage[i]] += ms;
      }
    }
    reset_timings();
  }

  // Enqueues, on lane L, the compression of data[range_start, range_start + range_len) of the stream resident at d_data
  // (absolute positions); its metablocks are appended to the output at the running bit position.  `after_layout` (may
  // be null) is the previous chunk's layout event; `layout_done` is recorded when this chunk's layout is final.
  bool run_chunk(Lane& L, const EncParams& Pstream, uint32_t range_start, uint32_t range_len, uint32_t* d_outw,
                 uint64_t out_cap_bytes, bool first, bool last, bool byte_align_end, uint32_t chunk_idx,
                 cudaEvent_t after_layout, cudaEvent_t layout_done) {
    cudaStream_t stream = L.stream;
    Workspace W;
    memset(&W, 0, sizeof(W));
    EncParams P = Pstream;
    P.n = range_len;
    P.abs_base = range_start;
    if (!ensure_chunk(L, range_len, P, &W)) return false;
    W.P = P;
    W.data = d_data.as<uint8_t>() + (range_start - data_base);
    W.out = d_outw;
    W.out_cap_bytes = out_cap_bytes;
    const uint8_t* d_all = d_data.as<uint8_t>() - data_base;  // indexable by absolute position >= data_base
    k_init_mb<<<(W.num_mb + 63) / 64, 64, 0, stream>>>(W);
    // the literal context decision needs the input only: it runs here, under the throughput-bound stages of the other lanes,
    // instead of in the latency-bound tail of the chunk
    k_ctx_decide<<<W.num_mb, 256, 0, stream>>>(W);
    // ---- sort + match, batch by batch ----
    const uint32_t window = 1u << P.lgwin;
    const uint32_t payload_max = kBatchMax - window - 4096;
    // q7..q9 (bucket depth >= 64): the parse searches the buckets on demand when the chunk is a single sort batch
    // (measured, profiles/r02u_q9_ab.log: depth >= 128 -- q8, q9 and the lgwin <= 16 configurations -- gains 1.7x..2.8x on JSON logs
    // and periodic data and is within +-10 % on text; depth 64 (q7) and inputs of a few units are faster up front.  ondemand = 2
    // forces the on-demand path for every deep configuration, 0 switches it off.)
    const bool od_shape = P.quality < 10 && (P.depth == 64 || P.depth == 128 || P.depth == 256) && range_len <= payload_max &&
                          (P.n_last == 4 || P.n_last == 10 || P.n_last == 16);
    const bool od = od_shape && (ondemand > 1 || (ondemand == 1 && P.depth >= 128 && range_len >= ((uint32_t)4 << 20)));
    DeepArgs da;
    memset(&da, 0, sizeof(da));
    for (uint64_t b0 = range_start; b0 < (uint64_t)range_start + range_len; b0 += payload_max) {
      const uint32_t b1 = (uint32_t)std::min<uint64_t>((uint64_t)range_start + range_len, b0 + payload_max);
      uint32_t origin = b0 > window ? (uint32_t)b0 - window : 0u;
      origin &= ~4095u;  // tile staging needs word alignment
      if (origin < data_base) origin = (uint32_t)data_base;
      const uint32_t count = b1 - origin;
      const uint32_t tiles = (count + SORT_TILE - 1) / SORT_TILE;
      mark(L, B200_ST_SORT);
      SortArgs sa;
      sa.data = d_all + origin;
      sa.count = count;
      sa.hist = L.d_hist.as<uint32_t>();
      sa.digit_base = L.d_digit.as<uint32_t>();
      sa.num_tiles = tiles;
      sa.hash_type = P.hash_type;
      sa.key_bits = P.key_bits;
      for (int pass = 0; pass < 2; ++pass) {
        sa.pass = pass;
        sa.in = pass == 0 ? nullptr : L.d_sortA.as<uint32_t>();
        sa.outw = pass == 0 ? L.d_sortA.as<uint32_t>() : L.d_sortB.as<uint32_t>();
        k_sort_hist<false><<<tiles, SORT_THREADS, 0, stream>>>(sa);
        k_scan_rows<<<256, 256, 0, stream>>>(sa.hist, tiles, L.d_digit.as<uint32_t>() + 256);
        k_scan_digits<<<1, 256, 0, stream>>>(L.d_digit.as<uint32_t>() + 256, L.d_digit.as<uint32_t>());
        k_sort_scatter<false><<<tiles, SORT_THREADS, 0, stream>>>(sa);
        launches += 4;
      }
      MatchArgs ma;
      ma.data = d_all;
      ma.sorted = L.d_sortB.as<uint32_t>();
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
      ma.dict = W.dict;
      ma.use_dict = P.use_dict;
      const size_t smem = (size_t)(MATCH_THREADS + P.depth) * 6 * 4;
      mark(L, B200_ST_MATCH);
      const uint32_t mgrid = (count + MATCH_THREADS - 1) / MATCH_THREADS;
      if (od) {  // ranks into best[], signatures into the free half of the sort ping-pong
        da.m = ma;
        da.sig = L.d_sortA.as<uint32_t>();
        k_rank_sig<<<(count + 255) / 256, 256, 0, stream>>>(ma, L.d_sortA.as<uint32_t>());
      } else
      if (P.quality >= 10) {  // all matches of every position
        MatchAllArgs aa;
        aa.m = ma;
        aa.hqm = W.hqm - (size_t)range_start * HQ_MAXM;
        aa.hqn = W.hqn - range_start;
        aa.quality = P.quality;
        aa.level = 0;
        aa.last_pass = P.hq_levels == 0;
        if (P.depth == 256) k_match_all<256><<<mgrid, MATCH_THREADS, (size_t)(MATCH_THREADS + 256) * 3 * 4, stream>>>(aa);
        else if (P.depth == 1024) k_match_all<1024><<<mgrid, MATCH_THREADS, (size_t)(MATCH_THREADS + 1024) * 3 * 4, stream>>>(aa);
        else { fprintf(stderr, "[brotli_b200] unsupported bucket depth %d\n", P.depth); return false; }
        for (int lv = 0; lv < P.hq_levels; ++lv) {  // long-prefix levels: the batch re-sorted by the level's hash, lists merged
          SortArgs sl = sa;
          sl.hash_type = BRO_HASH_LEVEL0 + lv;
          for (int pass = 0; pass < 2; ++pass) {
            sl.pass = pass;
            sl.in = pass == 0 ? nullptr : L.d_sortA.as<uint32_t>();
            sl.outw = pass == 0 ? L.d_sortA.as<uint32_t>() : L.d_sortB.as<uint32_t>();
            k_sort_hist<true><<<tiles, SORT_THREADS, 0, stream>>>(sl);
            k_scan_rows<<<256, 256, 0, stream>>>(sl.hist, tiles, L.d_digit.as<uint32_t>() + 256);
            k_scan_digits<<<1, 256, 0, stream>>>(L.d_digit.as<uint32_t>() + 256, L.d_digit.as<uint32_t>());
            k_sort_scatter<true><<<tiles, SORT_THREADS, 0, stream>>>(sl);
            launches += 4;
          }
          aa.level = lv;
          aa.last_pass = lv + 1 == P.hq_levels;
          k_match_level<HQ_LEVEL_DEPTH><<<mgrid, MATCH_THREADS, (size_t)(MATCH_THREADS + HQ_LEVEL_DEPTH) * 3 * 4, stream>>>(aa);
          launches += 1;
        }
      } else
      switch (P.depth) {  // bucket depth = 1 << block_bits: 16 (q5) .. 256 (q9, and lgwin <= 16)
        case 16:
          k_match_shallow<16><<<mgrid, MATCH_THREADS, smem, stream>>>(ma);
          break;
        case 32:
          k_match_shallow<32><<<mgrid, MATCH_THREADS, smem, stream>>>(ma);
          break;
        case 64: k_match_deep<64><<<mgrid, MATCH_THREADS, smem, stream>>>(ma); break;
        case 128: k_match_deep<128><<<mgrid, MATCH_THREADS, smem, stream>>>(ma); break;
        case 256: k_match_deep<256><<<mgrid, MATCH_THREADS, smem, stream>>>(ma); break;
        default: fprintf(stderr, "[brotli_b200] unsupported bucket depth %d\n", P.depth); return false;
      }
      launches += 1;
    }
    mark(L, B200_ST_PARSE);
    if (P.quality >= 10) {  // shortest-path parse, one unit per warp
      ZopfliArgs za;
      za.nodes = L.d_hq_nodes.as<ZNode>();
      za.pre = L.d_hq_pre.as<uint32_t>();
      za.scratch = L.d_hq_scratch.as<uint32_t>();
      for (int phase = 1; phase <= (P.quality >= 11 ? 2 : 1); ++phase) {
        if (hq_thread_units) k_zopfli<<<(W.num_units + 31) / 32, 32, 0, stream>>>(W, za, 1u, phase);
        else k_zopfli<<<W.num_units, 32, 0, stream>>>(W, za, 32u, phase);
      }
    } else
    if (od) {
      const uint32_t pg = (W.num_units + PARSE_WARPS - 1) / PARSE_WARPS;
#define B200_OD_LAUNCH(NLV) \
      switch (P.depth) { \
        case 64: k_parse_ondemand<NLV, 64><<<pg, PARSE_WARPS * 32, 0, stream>>>(W, da); break; \
        case 128: k_parse_ondemand<NLV, 128><<<pg, PARSE_WARPS * 32, 0, stream>>>(W, da); break; \
        default: k_parse_ondemand<NLV, 256><<<pg, PARSE_WARPS * 32, 0, stream>>>(W, da); break; \
      }
      if (P.n_last == 4) { B200_OD_LAUNCH(4) }
      else if (P.n_last == 10) { B200_OD_LAUNCH(10) }
      else { B200_OD_LAUNCH(16) }
#undef B200_OD_LAUNCH
    } else
    if (pair_parse == 4 && P.n_last == 4 && P.hash_type != 9)  // four units per warp (q5, q6)
      k_parse_pair<4><<<(W.num_units + 4 * PARSE_WARPS - 1) / (4 * PARSE_WARPS), PARSE_WARPS * 32, 0, stream>>>(W);
    else if (pair_parse && P.n_last == 4 && P.hash_type != 9)  // two units per warp
      k_parse_pair<2><<<(W.num_units + 2 * PARSE_WARPS - 1) / (2 * PARSE_WARPS), PARSE_WARPS * 32, 0, stream>>>(W);
    else
      k_parse<<<(W.num_units + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, stream>>>(W);
    mark(L, B200_ST_FINALIZE);
    k_fin_count<<<W.num_mb, 1024, 0, stream>>>(W);
    k_fin_write<<<(W.num_units + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, stream>>>(W);
    k_fin_dist<<<W.num_mb, 1024, 0, stream>>>(W);
    if (P.quality >= 10 && P.hq_split) {  // NPOSTFIX / NDIRECT of every metablock (metablock.rs:152-207), commands re-coded
      if (!L.d_dist_cost.ensure((size_t)W.num_mb * 64 * 8)) return false;
      k_dist_cost<<<dim3(64, W.num_mb), 256, 0, stream>>>(W, L.d_dist_cost.as<uint64_t>());
      k_dist_apply<<<dim3(64, W.num_mb), 256, 0, stream>>>(W, L.d_dist_cost.as<uint64_t>());
      launches += 2;
    }
    {
      dim3 g((W.cmd_cap + 255) / 256, W.num_mb);
      cudaMemsetAsync(W.long_tab, 0, (size_t)W.num_mb * W.long_cap * sizeof(uint2), stream);
      k_symbols<<<g, 256, 0, stream>>>(W);
      k_symbols_long<<<dim3(LONG_GRID, W.num_mb), 256, 0, stream>>>(W);
    }
    mark(L, B200_ST_SPLIT);
    {
      dim3 g(W.num_mb, 3);
      if (P.quality >= 10 && P.hq_split) { if (!run_hq_split(L, W)) return false; }
      else if (P.split) k_split_greedy<<<g, SPLIT_THREADS, SPLIT_SMEM_WORDS * 4, stream>>>(W);
      else k_split_simple<<<g, 512, 0, stream>>>(W);
    }
    mark(L, B200_ST_HEADER);
    {
      dim3 g(W.max_lit_trees + W.max_cmd_types + W.max_dist_types + HDR_SECTIONS, W.num_mb);
      k_trees<<<g, 32, 0, stream>>>(W);
    }
    k_header<<<W.num_mb, 32, 0, stream>>>(W);
    mark(L, B200_ST_EMIT);
    {
      dim3 g((W.cmd_cap + 255) / 256, W.num_mb);
      k_bitlen_long<<<dim3(LONG_GRID, W.num_mb), 256, 0, stream>>>(W);
      k_bitlen<<<g, 256, 0, stream>>>(W);
      k_bitscan<<<W.num_mb, 1024, 0, stream>>>(W);
      if (after_layout) CUDA_OK(cudaStreamWaitEvent(stream, after_layout, 0));  // bit positions chain through the chunks
      k_layout<<<1, 32, 0, stream>>>(W, first ? 1 : 0, last ? 1 : 0, byte_align_end ? 1 : 0, d_total.as<uint64_t>() + 1 + chunk_idx);
      CUDA_OK(cudaEventRecord(layout_done, stream));
      k_emit_header<<<W.num_mb, 256, 0, stream>>>(W);
      k_emit_body<<<g, 256, 0, stream>>>(W);
      k_emit_long<<<dim3(LONG_GRID, W.num_mb), 256, 0, stream>>>(W);
      dim3 gr(64, W.num_mb);
      k_emit_raw<<<gr, 256, 0, stream>>>(W);
    }
    mark(L, -1);
    launches += 19;
    CUDA_OK(cudaGetLastError());
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
// The quality the device path runs for a requested one: 5..9 hash-chain family (encode.rs:834-893), 10 / 11 shortest-path parse;
// q0..q4 (BasicHasher H2..H54, fragment compressors) are not built and run as 5.
int b200_effective_quality(int requested_quality) {
  if (requested_quality < 5) return 5;
  if (requested_quality > 11) return 11;
  return requested_quality;
}

B200Encoder* b200_encoder_create(int device) {
  B200Encoder* e = new B200Encoder();
  if (!e->init(device)) {
    delete e;
    return nullptr;
  }
  return e;
}
int b200_encoder_device(const B200Encoder* e) { return e ? e->device : -1; }
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
    case B200_OPT_DICT: e->use_dict = (int)value; return 1;
    case B200_OPT_SHALLOW_MATCH: e->shallow_match = (int)value; return 1;
    case B200_OPT_PAIR_PARSE: e->pair_parse = (int)value; return 1;
    case B200_OPT_ONDEMAND: e->ondemand = (int)value; return 1;
    case B200_OPT_HQ_LEVELS: e->hq_levels = value > HQ_MAX_LEVELS ? HQ_MAX_LEVELS : (int)value; return 1;
    case B200_OPT_HQ_SPLIT: e->hq_split = (int)value; return 1;
    case B200_OPT_HQ_UNIT: e->hq_unit = value; return 1;
    case B200_OPT_HQ_THREAD_UNITS: e->hq_thread_units = (int)value; return 1;
    case B200_OPT_LANES: e->num_lanes = value < 1 ? 1 : (value > (uint32_t)kMaxLanes ? kMaxLanes : (int)value); return 1;
  }
  return 0;
}

size_t b200_max_compressed_size(size_t n) { return n + (n >> 10) * 8 + 4096; }

// Compresses [range_start, range_start+range_len) of an n-byte stream.  in/out are device pointers when
// device_io != 0, host pointers otherwise.  first/last: emit stream header / final empty metablock;
// byte_align: end the range with a padding metablock so that ranges can be concatenated with memcpy.
//
// Pipeline: the input is staged chunk by chunk on a copy stream, chunks alternate between two compute lanes, and the
// finished part of the output is copied back while later chunks are still running.
static bool compress_range_impl(B200Encoder* e, int quality, int lgwin, uint64_t size_hint, const uint8_t* in, size_t n,
                                size_t range_start, size_t range_len, bool first, bool last, bool byte_align, uint8_t* out,
                                size_t out_cap, size_t* out_size, int device_io, bool keep_on_device) {
  e->launches = 0;
  e->reset_timings();
  e->sync_events.reset();
  EncParams P;
  e->fill_params(&P, quality, lgwin, size_hint ? size_hint : n);
  const size_t window = (size_t)1 << P.lgwin;
  const size_t base = range_start > window ? ((range_start - window) & ~(size_t)4095) : 0;
  const size_t end = range_start + range_len;
  const size_t staged = end - base;
  const size_t need = b200_max_compressed_size(range_len) + 64;
  std::vector<std::pair<size_t, size_t>> chunks;  // (absolute start, length)
  for (size_t done = 0; done < range_len;) {
    const size_t len = chunk_len_at(done, range_len);
    chunks.emplace_back(range_start + done, len);
    done += len;
  }
  const size_t nchunks = chunks.size();
  // device_io: 0 host in / host out, 1 device in / device out, 2 host in / device out, 3 device in / host out
  const bool in_dev = device_io == 1 || device_io == 3, out_dev = device_io == 1 || device_io == 2;
  const cudaMemcpyKind in_kind = in_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out_kind = out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (!e->d_data.ensure(staged + kPad) || !e->d_out.ensure(need) || !e->ensure_totals(nchunks)) return false;
  e->data_base = base;
  uint8_t* dd = e->d_data.as<uint8_t>();
  CUDA_OK(cudaMemsetAsync(e->d_out.p, 0, need, e->s_in));
  CUDA_OK(cudaMemsetAsync(e->d_total.p, 0, 8, e->s_in));
  CUDA_OK(cudaMemsetAsync(dd + staged, 0, kPad, e->s_in));
  std::vector<cudaEvent_t> ev_done(nchunks);
  cudaEvent_t prev_layout = nullptr;
  size_t copied = base;  // absolute position up to which the input is staged
  for (size_t k = 0; k < nchunks; ++k) {
    const size_t s = chunks[k].first, len = chunks[k].second;
    // stage the input this chunk can see: its window halo (first chunk), its own bytes, a short look-ahead
    const size_t upto = std::min(end, s + len + kLookahead);
    if (upto > copied) {
      CUDA_OK(cudaMemcpyAsync(dd + (copied - base), in + copied, upto - copied, in_kind, e->s_in));
      copied = upto;
    }
    cudaEvent_t ev_in = e->sync_events.get(false);
    CUDA_OK(cudaEventRecord(ev_in, e->s_in));
    Lane& L = e->lanes[k % (size_t)e->num_lanes];
    CUDA_OK(cudaStreamWaitEvent(L.stream, ev_in, 0));
    cudaEvent_t ev_layout = e->sync_events.get(false);
    const bool f = first && k == 0, l = k + 1 == nchunks;
    if (!e->run_chunk(L, P, (uint32_t)s, (uint32_t)len, e->d_out.as<uint32_t>(), need, f, last && l, byte_align && l, (uint32_t)k,
                      prev_layout, ev_layout))
      return false;
    prev_layout = ev_layout;
    CUDA_OK(cudaMemcpyAsync(e->h_total + k, e->d_total.as<uint64_t>() + 1 + k, 8, cudaMemcpyDeviceToHost, L.stream));
    ev_done[k] = e->sync_events.get(false);
    CUDA_OK(cudaEventRecord(ev_done[k], L.stream));
  }
  // drain: as each chunk finishes, every output byte below its end bit position is final
  size_t done_bytes = 0;
  for (size_t k = 0; k < nchunks; ++k) {
    if (cudaEventSynchronize(ev_done[k]) != cudaSuccess) {
      fprintf(stderr, "[brotli_b200] kernel failure: %s\n", cudaGetErrorString(cudaGetLastError()));
      return false;
    }
    const uint64_t tb = e->h_total[k];
    const size_t upto = k + 1 == nchunks ? (size_t)((tb + 7) >> 3) : (size_t)(tb >> 3);
    if (upto > out_cap) return false;
    if (!keep_on_device && upto > done_bytes)
      CUDA_OK(cudaMemcpyAsync(out + done_bytes, e->d_out.as<uint8_t>() + done_bytes, upto - done_bytes, out_kind, e->s_out));
    done_bytes = upto;
  }
  CUDA_OK(cudaStreamSynchronize(e->s_out));
  CUDA_OK(cudaStreamSynchronize(e->s_in));
  *out_size = done_bytes;
  if (e->timing) e->collect_timings();
  return true;
}

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
      if (device_io == 1 || device_io == 2) { if (cudaMemcpy(out, &b, 1, cudaMemcpyHostToDevice) != cudaSuccess) return 0; }
      else out[0] = b;
      *out_size = 1;
      return 1;
    }
    *out_size = 0;
    return 1;
  }
  if (!compress_range_impl(e, quality, lgwin, size_hint, in, n, range_start, range_len, first != 0, last != 0, byte_align != 0,
                           out, out_cap, out_size, device_io, false)) {
    cudaDeviceSynchronize();  // leave no work in flight behind a failed call
    return 0;
  }
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

// test hook: device results of the match stage for an n-byte buffer (host in, host out); n <= one chunk
int b200_stage_match(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint32_t* best_out) {
  if (!e || !e->ok || n == 0 || n > kChunk) return 0;
  if (cudaSetDevice(e->device) != cudaSuccess) return 0;
  size_t got = 0;
  if (!compress_range_impl(e, quality, lgwin, n, in, n, 0, n, true, true, false, nullptr, b200_max_compressed_size(n) + 64, &got, 0, true)) {
    cudaDeviceSynchronize();
    return 0;
  }
  return cudaMemcpy(best_out, e->lanes[0].d_best.p, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
}

// test hook (quality >= 10): matches per position, per-unit results and raw commands of an n-byte buffer (n <= one chunk)
int b200_stage_hq(B200Encoder* e, int quality, int lgwin, const uint8_t* in, size_t n, uint8_t* hqn, uint32_t* hqm, uint32_t* units,
                  uint32_t* raw) {
  if (!e || !e->ok || n == 0 || n > kChunk || quality < 10) return 0;
  if (cudaSetDevice(e->device) != cudaSuccess) return 0;
  size_t got = 0;
  if (!compress_range_impl(e, quality, lgwin, n, in, n, 0, n, true, true, false, nullptr, b200_max_compressed_size(n) + 64, &got, 0, true)) {
    cudaDeviceSynchronize();
    return 0;
  }
  EncParams P;
  e->fill_params(&P, quality, lgwin, n);
  const uint32_t nu = (uint32_t)((n + P.unit - 1) / P.unit);
  Lane& L = e->lanes[0];
  bool ok = cudaMemcpy(hqn, L.d_hqn.p, n, cudaMemcpyDeviceToHost) == cudaSuccess;
  ok = ok && cudaMemcpy(hqm, L.d_hqm.p, n * HQ_MAXM * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  ok = ok && cudaMemcpy(units, L.d_unit.p, (size_t)nu * 3 * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
  ok = ok && cudaMemcpy(raw, L.d_raw.p, (size_t)nu * (P.unit / 2 + 1) * 12, cudaMemcpyDeviceToHost) == cudaSuccess;
  return ok ? 1 : 0;
}

}  // extern "C"
