// tools/gpu_model.cpp -- TEST/DEVELOPMENT INFRASTRUCTURE (not part of the product library).
//
// Sequential CPU model of the B200 pipeline: every stage below is the deterministic specification of one
// CUDA kernel in rust-brotli_b200/csrc (same integer arithmetic, same tie-breaking), built from the same
// host/device headers.  It exists because the development container has no GPU: the algorithm (and its
// compressed-size behaviour) is developed here, and on the GPU box tests assert that the kernels reproduce
// the model's stream bit-for-bit.
//
// Build: g++ -O2 -shared -fPIC -I rust-brotli_b200/csrc tools/gpu_model.cpp -o tools/libgpu_model.so
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bro_common.cuh"
#include "bro_finalize.cuh"
#include "bro_huffman.cuh"
#include "bro_meta.cuh"
#include "bro_parse.cuh"
#include "bro_split.cuh"

using namespace bro;

namespace {

struct PlainOrWriter {  // same interface as the device atomic-OR writer
  uint8_t* out;
  uint64_t pos;
  void put(uint32_t n, uint64_t v) {
    for (uint32_t i = 0; i < n; ++i, ++pos)
      if ((v >> i) & 1) out[pos >> 3] |= (uint8_t)(1u << (pos & 7));
  }
  void skip(uint32_t n) { pos += n; }
};

struct Model {
  EncParams P;
  std::vector<uint8_t> data;
  std::vector<uint32_t> lut;
  std::vector<uint32_t> best;
  std::vector<RawCmd> raw;
  std::vector<uint32_t> unit_ncmd, unit_tail, unit_ncopy;
  uint32_t data_shift = 0;  // range start inside `data`
  const uint8_t* d() const { return data.data() + data_shift; }
};

void stage_match(Model& M) {
  const EncParams& P = M.P;
  const uint32_t N = P.n;
  const uint8_t* d = M.data.data();
  M.best.assign(N + 1, 0);
  const uint32_t nb = 1u << P.key_bits;
  const uint32_t D = (uint32_t)P.depth;
  std::vector<uint32_t> ring((size_t)nb * D);
  std::vector<uint32_t> cnt(nb, 0);
  for (uint32_t p = 0; p < N; ++p) {
    uint32_t key = hash_key(P.hash_type, P.key_bits, d + p);
    uint32_t maxl = bmin(P.lcap, N - p);
    uint32_t c = cnt[key];
    uint32_t best_score = BRO_MIN_SCORE, best_len = 0, best_dist = 0;
    uint32_t max_backward = bmin(p, P.max_backward);
    if (N - p >= 8) {  // keys of the last 7 positions would depend on bytes past the range: no bucket match
      uint32_t lim = bmin(c, D);
      for (uint32_t k = 1; k <= lim; ++k) {
        uint32_t cand = ring[(size_t)key * D + ((c - k) & (D - 1))];
        uint32_t backward = p - cand;
        if (backward > max_backward) break;
        if (best_len < maxl && d[cand + best_len] != d[p + best_len]) continue;
        uint32_t len = lcp_bytes(d + cand, d + p, maxl);
        if (len >= 4) {
          uint32_t score = score_regular(P.hash_type, len, backward);
          if (score > best_score) {
            best_score = score; best_len = len; best_dist = backward;
          }
          if (len == maxl) break;
        }
      }
    }
    M.best[p] = best_len ? ((best_dist << 8) | best_len) : 0;
    ring[(size_t)key * D + (c & (D - 1))] = p;
    cnt[key] = c + 1;
  }
}

void stage_parse(Model& M) {
  const EncParams& P = M.P;
  const uint32_t NU = (P.n + P.unit - 1) / P.unit;
  const uint32_t CU = P.unit / 2 + 1;
  M.raw.assign((size_t)NU * CU, RawCmd{0, 0, 0});
  M.unit_ncmd.assign(NU, 0);
  M.unit_tail.assign(NU, 0);
  M.unit_ncopy.assign(NU, 0);
  for (uint32_t u = 0; u < NU; ++u) {
    uint32_t s = u * P.unit, e = bmin(P.n, s + P.unit), tail, ncopy;
    M.unit_ncmd[u] = parse_unit(P, M.d(), M.best.data(), s, e, &M.raw[(size_t)u * CU], &tail, &ncopy);
    M.unit_tail[u] = tail;
    M.unit_ncopy[u] = ncopy;
  }
}

struct MetaBlock {
  uint32_t start, len;
  std::vector<GCmd> cmds;
  uint32_t nlit, ndist;
  int ctx_map_id;
  // splits
  SplitResult lit, cmd, dist;
  std::vector<uint8_t> lit_depth, cmd_depth, dist_depth;
  std::vector<uint16_t> lit_code, cmd_code, dist_code;
  SplitCode lit_sc, cmd_sc, dist_sc;
  std::vector<uint8_t> hdr;
  uint64_t hdr_bits, body_bits;
  std::vector<uint64_t> cmd_bitpos;
  bool raw;
};

void stage_finalize(Model& M, MetaBlock& mb, uint32_t u0, uint32_t u1) {
  const EncParams& P = M.P;
  UnitView V;
  V.raw = M.raw.data(); V.ncmd = M.unit_ncmd.data(); V.tail = M.unit_tail.data();
  V.cu = P.unit / 2 + 1; V.unit = P.unit; V.n = P.n;
  // kernel 1: per-unit counts; then exclusive scans (commands, literals); kernel 2: write; scan distances; kernel 3: add
  std::vector<uint32_t> cmd_off(u1 - u0 + 1, 0), lit_off(u1 - u0 + 1, 0), dist_off(u1 - u0 + 1, 0);
  for (uint32_t u = u0; u < u1; ++u) {
    uint32_t ulen = bmin(P.n, (u + 1) * P.unit) - u * P.unit;
    cmd_off[u - u0 + 1] = cmd_off[u - u0] + unit_final_ncmd(V, u0, u1, u);
    lit_off[u - u0 + 1] = lit_off[u - u0] + (ulen - M.unit_ncopy[u]);
  }
  mb.cmds.assign(cmd_off[u1 - u0], GCmd());
  for (uint32_t u = u0; u < u1; ++u)
    dist_off[u - u0 + 1] = dist_off[u - u0] + finalize_unit(V, u0, u1, u, lit_off[u - u0], mb.cmds.data() + cmd_off[u - u0]);
  for (uint32_t u = u0; u < u1; ++u)
    for (uint32_t i = cmd_off[u - u0]; i < cmd_off[u - u0 + 1]; ++i) mb.cmds[i].dist_idx += dist_off[u - u0];
  mb.nlit = lit_off[u1 - u0];
  mb.ndist = dist_off[u1 - u0];
  uint64_t cover = 0;
  for (auto& c : mb.cmds) cover += c.insert_len + c.copy_len;
  if (cover != mb.len) fprintf(stderr, "model: coverage mismatch %llu vs %u\n", (unsigned long long)cover, mb.len);
}

void stage_ctx_decide(Model& M, MetaBlock& mb) {
  const EncParams& P = M.P;
  mb.ctx_map_id = CTXMAP_NONE;
  if (!P.ctx_model) return;
  mb.ctx_map_id = decide_literal_context_map(P.quality, P.size_hint, M.d(), mb.start, mb.len, M.lut.data());
}

void stage_split_and_histograms(Model& M, MetaBlock& mb) {
  const EncParams& P = M.P;
  const uint8_t* d = M.d();
  const uint32_t nctx = ctxmap_num_contexts(mb.ctx_map_id);
  // symbol streams
  std::vector<uint16_t> lits(mb.nlit);   // literal | ctx << 8
  std::vector<uint16_t> cmds(mb.cmds.size()), dists(mb.ndist);
  for (size_t i = 0; i < mb.cmds.size(); ++i) {
    const GCmd& c = mb.cmds[i];
    cmds[i] = c.cmd_prefix;
    uint32_t pos = c.pos;
    for (uint32_t j = 0; j < c.insert_len; ++j) {
      uint32_t p = pos + j;
      uint8_t p1 = (P.abs_base || p >= 1) ? d[(int64_t)p - 1] : 0, p2 = (P.abs_base || p >= 2) ? d[(int64_t)p - 2] : 0;
      uint32_t cx = mb.ctx_map_id ? ctxmap_lookup(mb.ctx_map_id, context_utf8(p1, p2)) : 0;
      lits[c.lit_idx + j] = (uint16_t)(d[p] | (cx << 8));
    }
    if (c.as_cmd().has_distance()) dists[c.dist_idx] = c.dist_prefix & 0x3ff;
  }
  const bool split = P.split != 0;
  greedy_split(lits.data(), mb.nlit, 256, nctx, 512, 400, split, M.lut.data(), &mb.lit);
  greedy_split(cmds.data(), (uint32_t)cmds.size(), 704, 1, 1024, 500, split, M.lut.data(), &mb.cmd);
  greedy_split(dists.data(), mb.ndist, 64, 1, 512, 100, split, M.lut.data(), &mb.dist);
}

SplitView view_of(const SplitResult& r) {
  SplitView v;
  v.num_types = r.num_types;
  v.num_blocks = (uint32_t)r.types.size();
  v.types = r.types.data();
  v.lengths = r.lengths.data();
  v.starts = r.starts.data();
  return v;
}

void stage_header(Model& M, MetaBlock& mb) {
  const EncParams& P = M.P;
  const uint32_t nctx = ctxmap_num_contexts(mb.ctx_map_id);
  HuffStoreWs* ws = new HuffStoreWs;
  std::vector<uint8_t> good(704);
  if (P.use_rle_opt) {
    for (uint32_t t = 0; t < mb.lit.num_types * nctx; ++t) huff_optimize_counts_for_rle(256, &mb.lit.histograms[(size_t)t * 256], good.data());
    for (uint32_t t = 0; t < mb.cmd.num_types; ++t) huff_optimize_counts_for_rle(704, &mb.cmd.histograms[(size_t)t * 704], good.data());
    for (uint32_t t = 0; t < mb.dist.num_types; ++t) huff_optimize_counts_for_rle(64, &mb.dist.histograms[(size_t)t * 64], good.data());
  }
  size_t cap = 4096 + 1024 * (size_t)(mb.lit.num_types * nctx + mb.cmd.num_types + mb.dist.num_types) +
               64 * (size_t)(mb.lit.types.size() + 1);
  mb.hdr.assign(cap, 0);
  BitWriter bw;
  bw.init(mb.hdr.data());
  store_compressed_metablock_header(bw, false, mb.len);
  SplitView lv = view_of(mb.lit), cv = view_of(mb.cmd), dv = view_of(mb.dist);
  store_block_split_code(bw, lv, &mb.lit_sc, ws);
  store_block_split_code(bw, cv, &mb.cmd_sc, ws);
  store_block_split_code(bw, dv, &mb.dist_sc, ws);
  bw.put(2, 0);  // NPOSTFIX
  bw.put(4, 0);  // NDIRECT
  for (uint32_t i = 0; i < mb.lit.num_types; ++i) bw.put(2, 2);  // CONTEXT_UTF8
  if (mb.ctx_map_id == CTXMAP_NONE) store_trivial_context_map(bw, mb.lit.num_types, 6, ws);
  else {
    std::vector<uint32_t> rle((size_t)mb.lit.num_types << 6);
    store_static_literal_context_map(bw, mb.lit.num_types, mb.ctx_map_id, rle.data(), ws);
  }
  store_trivial_context_map(bw, mb.dist.num_types, 2, ws);
  uint32_t nlt = mb.lit.num_types * nctx;
  mb.lit_depth.assign((size_t)nlt * 256, 0); mb.lit_code.assign((size_t)nlt * 256, 0);
  mb.cmd_depth.assign((size_t)mb.cmd.num_types * 704, 0); mb.cmd_code.assign((size_t)mb.cmd.num_types * 704, 0);
  mb.dist_depth.assign((size_t)mb.dist.num_types * 64, 0); mb.dist_code.assign((size_t)mb.dist.num_types * 64, 0);
  for (uint32_t t = 0; t < nlt; ++t)
    huff_build_and_store(bw, &mb.lit.histograms[(size_t)t * 256], 256, 256, ws, &mb.lit_depth[(size_t)t * 256], &mb.lit_code[(size_t)t * 256]);
  for (uint32_t t = 0; t < mb.cmd.num_types; ++t)
    huff_build_and_store(bw, &mb.cmd.histograms[(size_t)t * 704], 704, 704, ws, &mb.cmd_depth[(size_t)t * 704], &mb.cmd_code[(size_t)t * 704]);
  for (uint32_t t = 0; t < mb.dist.num_types; ++t)
    huff_build_and_store(bw, &mb.dist.histograms[(size_t)t * 64], 64, 64, ws, &mb.dist_depth[(size_t)t * 64], &mb.dist_code[(size_t)t * 64]);
  bw.flush_partial();
  mb.hdr_bits = bw.bit_pos();
  if ((mb.hdr_bits + 7) / 8 > cap) { fprintf(stderr, "model: header overflow\n"); abort(); }
  delete ws;
}

MetaCodes codes_of(const MetaBlock& mb, const SplitView& lv, const SplitView& cv, const SplitView& dv) {
  MetaCodes mc;
  mc.lit = lv; mc.cmd = cv; mc.dist = dv;
  mc.lit_sc = &mb.lit_sc; mc.cmd_sc = &mb.cmd_sc; mc.dist_sc = &mb.dist_sc;
  mc.lit_depth = mb.lit_depth.data(); mc.lit_code = mb.lit_code.data();
  mc.cmd_depth = mb.cmd_depth.data(); mc.cmd_code = mb.cmd_code.data();
  mc.dist_depth = mb.dist_depth.data(); mc.dist_code = mb.dist_code.data();
  mc.ctx_map_id = mb.ctx_map_id;
  mc.nctx = ctxmap_num_contexts(mb.ctx_map_id);
  return mc;
}

void stage_bitlen(Model& M, MetaBlock& mb) {
  SplitView lv = view_of(mb.lit), cv = view_of(mb.cmd), dv = view_of(mb.dist);
  MetaCodes mc = codes_of(mb, lv, cv, dv);
  mb.cmd_bitpos.resize(mb.cmds.size());
  uint64_t total = 0;
  for (size_t i = 0; i < mb.cmds.size(); ++i) {
    CountWriter w{0};
    emit_command(w, mc, mb.cmds[i].as_cmd(), (uint32_t)i, mb.cmds[i].lit_idx, mb.cmds[i].dist_idx, M.d(), mb.cmds[i].pos, M.P.abs_base);
    mb.cmd_bitpos[i] = total;
    total += w.bits;
  }
  mb.body_bits = total;
}

}  // namespace

extern "C" {

struct ModelStats {
  uint64_t num_metablocks, num_raw_metablocks, num_commands, num_literals, header_bits, body_bits;
  uint32_t lit_types_total, cmd_types_total, dist_types_total, ctx_ids[4];
};

void gpu_model_default_params(EncParams* P, int quality, int lgwin, uint32_t n, uint32_t size_hint) {
  memset(P, 0, sizeof(*P));
  if (quality < 5) quality = 5;
  if (quality > 9) quality = 9;
  if (lgwin < 10) lgwin = 10;
  if (lgwin > 24) lgwin = 24;
  if (size_hint == 0) size_hint = n;
  P->quality = quality;
  P->lgwin = lgwin;
  P->n = n;
  P->size_hint = size_hint;
  if (quality == 9) { P->hash_type = 9; P->key_bits = 15; P->hash_len = 4; P->depth = 256; P->n_last = 16; }
  else if (lgwin <= 16) { P->hash_type = 6; P->key_bits = 15; P->hash_len = 5; P->depth = 256; P->n_last = 16; }
  else if (size_hint > (1u << 22) && lgwin >= 19) {
    P->hash_type = 6; P->key_bits = 15; P->hash_len = 5; P->depth = 1 << (quality - 1);
    P->n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  } else {
    P->hash_type = 5; P->key_bits = (quality < 7 && size_hint <= (1u << 20)) ? 14 : 15; P->hash_len = 4;
    P->depth = 1 << (quality - 1);
    P->n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  }
  P->lcap = 64;
  P->unit = 4096;
  P->mb_units = 1024;  // 4 MiB metablocks
  P->max_backward = (1u << lgwin) - 16;
  P->use_rle_opt = 1;
  P->split = 1;
  P->ctx_model = 1;
}

void gpu_model_fill_lut(uint32_t* lut) {
  lut[0] = 0;
  for (uint32_t i = 1; i < 65536; ++i) lut[i] = (uint32_t)llround(std::log2((double)i) * 65536.0);
}

// Runs the whole pipeline model.  Returns compressed size in bytes (0 on failure).
size_t gpu_model_compress_range(const EncParams* Pin, const uint8_t* input, uint32_t range_start, uint32_t range_len,
                                int first, int last, int byte_align, uint8_t* out, size_t out_cap, ModelStats* st,
                                uint32_t* best_out);

size_t gpu_model_compress(const EncParams* Pin, const uint8_t* input, uint8_t* out, size_t out_cap, ModelStats* st,
                          uint32_t* best_out /* optional [n] */) {
  return gpu_model_compress_range(Pin, input, 0, Pin->n, 1, 1, 0, out, out_cap, st, best_out);
}

static uint64_t model_chunk(const EncParams* Pin, const uint8_t* input, uint32_t range_start, uint32_t range_len,
                            int first, int last, int byte_align, uint8_t* out, size_t out_cap, ModelStats* st,
                            uint32_t* best_out, uint64_t bitpos_in);

// Mirrors b200_encoder_compress_range: Pin->n is the size of the whole stream `input`.  Like the device encoder the
// range is compressed as independent chunks of BRO_CHUNK_BYTES that append to one bit stream.
size_t gpu_model_compress_range(const EncParams* Pin, const uint8_t* input, uint32_t range_start, uint32_t range_len,
                                int first, int last, int byte_align, uint8_t* out, size_t out_cap, ModelStats* st,
                                uint32_t* best_out) {
  const uint32_t stream_n = Pin->n;
  if (st) memset(st, 0, sizeof(*st));
  if (stream_n == 0 || range_len == 0) {
    if (first && last && stream_n == 0) { out[0] = 6; return 1; }
    return 0;
  }
  memset(out, 0, out_cap);
  uint64_t bits = 0;
  const uint64_t end = (uint64_t)range_start + range_len;
  for (uint64_t s = range_start; s < end; s += BRO_CHUNK_BYTES) {
    const uint32_t len = (uint32_t)std::min<uint64_t>(BRO_CHUNK_BYTES, end - s);
    const bool l = s + len == end;
    bits = model_chunk(Pin, input, (uint32_t)s, len, first && s == range_start, last && l, byte_align && l, out, out_cap, st,
                       best_out ? best_out + (s - range_start) : nullptr, bits);
    if (bits == ~0ull) return 0;
  }
  return (size_t)((bits + 7) >> 3);
}

static uint64_t model_chunk(const EncParams* Pin, const uint8_t* input, uint32_t range_start, uint32_t range_len,
                            int first, int last, int byte_align, uint8_t* out, size_t out_cap, ModelStats* st,
                            uint32_t* best_out, uint64_t bitpos_in) {
  Model M;
  M.P = *Pin;
  // match stage over the whole prefix (absolute positions), then shift everything to range-relative
  M.P.n = range_start + range_len;
  M.P.abs_base = 0;
  M.data.assign((size_t)M.P.n + 320, 0);
  memcpy(M.data.data(), input, M.P.n);
  M.lut.resize(65536);
  gpu_model_fill_lut(M.lut.data());
  stage_match(M);
  if (range_start) {
    M.best.erase(M.best.begin(), M.best.begin() + range_start);
    M.data_shift = range_start;
  }
  M.P.n = range_len;
  M.P.abs_base = range_start;
  const EncParams& P = M.P;
  const uint32_t N = P.n;
  if (best_out) memcpy(best_out, M.best.data(), (size_t)N * 4);
  stage_parse(M);
  const uint32_t NU = (N + P.unit - 1) / P.unit;
  const uint32_t NM = (NU + P.mb_units - 1) / P.mb_units;
  std::vector<MetaBlock> mbs(NM);
  for (uint32_t m = 0; m < NM; ++m) {
    MetaBlock& mb = mbs[m];
    uint32_t u0 = m * P.mb_units, u1 = bmin(NU, u0 + P.mb_units);
    mb.start = u0 * P.unit;
    mb.len = bmin(N, u1 * P.unit) - mb.start;
    stage_finalize(M, mb, u0, u1);
    stage_ctx_decide(M, mb);
    stage_split_and_histograms(M, mb);
    stage_header(M, mb);
    stage_bitlen(M, mb);
  }
  // layout: stream header, metablocks, final empty metablock
  PlainOrWriter w{out, bitpos_in};
  if (first) {
    if (P.lgwin == 16) w.put(1, 0);
    else if (P.lgwin == 17) w.put(7, 1);
    else if (P.lgwin > 17) w.put(4, (uint64_t)(((P.lgwin - 17) << 1) | 1));
    else w.put(7, (uint64_t)(((P.lgwin - 8) << 4) | 1));
  }
  for (uint32_t m = 0; m < NM; ++m) {
    MetaBlock& mb = mbs[m];
    uint64_t comp_bits = mb.hdr_bits + mb.body_bits;
    uint64_t raw_hdr = raw_metablock_header_bits(mb.len);
    uint64_t raw_bits = ((w.pos + raw_hdr + 7) & ~7ull) - w.pos + 8ull * mb.len;
    mb.raw = comp_bits > raw_bits;
    if ((w.pos + bmax(comp_bits, raw_bits)) / 8 + 16 > out_cap) return ~0ull;
    if (st) {
      st->num_metablocks++;
      st->num_raw_metablocks += mb.raw;
      st->num_commands += mb.cmds.size();
      st->num_literals += mb.nlit;
      st->header_bits += mb.hdr_bits;
      st->body_bits += mb.body_bits;
      st->lit_types_total += mb.lit.num_types;
      st->cmd_types_total += mb.cmd.num_types;
      st->dist_types_total += mb.dist.num_types;
      st->ctx_ids[mb.ctx_map_id]++;
    }
    if (mb.raw) {
      uint32_t lg = mb.len == 1 ? 1u : log2_floor_nz(mb.len - 1) + 1u;
      uint32_t mnibbles = (lg < 16 ? 16u : lg + 3u) / 4u;
      w.put(1, 0);
      w.put(2, mnibbles - 4);
      w.put(mnibbles * 4, mb.len - 1);
      w.put(1, 1);
      w.pos = (w.pos + 7) & ~7ull;
      memcpy(out + (w.pos >> 3), M.d() + mb.start, mb.len);
      w.pos += 8ull * mb.len;
    } else {
      uint64_t base = w.pos;
      for (uint64_t b = 0; b < mb.hdr_bits; ++b)
        if ((mb.hdr[b >> 3] >> (b & 7)) & 1) out[(base + b) >> 3] |= (uint8_t)(1u << ((base + b) & 7));
      base += mb.hdr_bits;
      SplitView lv = view_of(mb.lit), cv = view_of(mb.cmd), dv = view_of(mb.dist);
      MetaCodes mc = codes_of(mb, lv, cv, dv);
      for (size_t i = 0; i < mb.cmds.size(); ++i) {
        PlainOrWriter cw{out, base + mb.cmd_bitpos[i]};
        emit_command(cw, mc, mb.cmds[i].as_cmd(), (uint32_t)i, mb.cmds[i].lit_idx, mb.cmds[i].dist_idx, M.d(), mb.cmds[i].pos, M.P.abs_base);
      }
      w.pos = base + mb.body_bits;
    }
  }
  if (last) {
    w.put(1, 1);  // ISLAST
    w.put(1, 1);  // ISLASTEMPTY
  } else if (byte_align && (w.pos & 7)) {
    w.put(6, 6);
    w.pos = (w.pos + 7) & ~7ull;
  }
  return w.pos;
}

}  // extern "C"
