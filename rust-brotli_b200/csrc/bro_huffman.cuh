// bro_huffman.cuh -- prefix codes: count smoothing, length-limited code lengths, canonical codes, code description.
//
// What the brotli format (RFC 7932 section 3) fixes: canonical code assignment from the lengths, the maximum length (15; 5 for the
// code-length code), the code-length alphabet (0..15 literal, 16 = repeat the previous non-zero length, 17 = repeat zero, with
// compounding repeats), the storage order of the code-length code and its fixed variable-length code, and the simple-code form
// for <= 4 symbols.  Everything else is this encoder's own:
//   * huff_smooth_counts   -- which counts are flattened so that the lengths run-length code well (the reference's counterpart
//                             is BrotliOptimizeHuffmanCountsForRle, entropy_encode.rs:211);
//   * huff_lengths_sorted  -- Huffman code lengths by a two-queue merge over the sorted counts with parent links (depths are
//                             read off the links), and a Kraft-sum repair when the tree is deeper than the limit (the reference
//                             rebuilds the tree with clamped counts until it fits, entropy_encode.rs:133-210);
//   * huff_rle_lengths     -- the run-length form of the length sequence: repeat counts are written in bijective base 4 / 8,
//                             and a run is split "literal + repeats" exactly when that needs fewer symbols.
// All routines are sequential (one GPU thread or the CPU model); k_trees replaces the sort by a warp-wide bitonic sort and
// runs the rest on one lane out of shared memory.
#pragma once
#include "bro_common.cuh"

namespace bro {

#define HUFF_MAX_SYMS 704u

struct HuffWs {  // scratch of one code (shared memory in k_trees)
  uint64_t key[1024];                  // used symbols: count << 16 | symbol, sorted ascending (1024: bitonic padding)
  uint32_t weight[2 * HUFF_MAX_SYMS];  // leaves in sorted order, then internal nodes in creation order
  uint16_t parent[2 * HUFF_MAX_SYMS];
  uint16_t node_depth[2 * HUFF_MAX_SYMS];
  uint8_t rle_sym[HUFF_MAX_SYMS];      // run-length form of the code lengths
  uint8_t rle_extra[HUFF_MAX_SYMS];
};
typedef HuffWs HuffStoreWs;

// Count smoothing before the prefix code is built (the job of BrotliOptimizeHuffmanCountsForRle, entropy_encode.rs:211-345: trade a few
// body bits for a code-length sequence that run-length codes well).  Own formulation, derived from the two costs involved:
// replacing the counts c_i of a run by their mean m costs about sum (c_i - m)^2 / (2 m ln 2) body bits, and saves about 2..3
// header bits per symbol once the run is long enough for a repeat code.  So a run (consecutive used symbols) is grown greedily
// while the next count stays within two standard deviations of the run's mean under a Poisson model, (c - m)^2 <= 4 m + slack
// (slack = min(16, total / 256): in a large histogram counts below ~10 get an absolute slack of 4), and a run of at least
// HUFF_SMOOTH_RUN symbols is set to its rounded mean.
#ifndef HUFF_SMOOTH_RUN
#define HUFF_SMOOTH_RUN 4u
#endif
#ifndef HUFF_SMOOTH_DENSITY_NUM
#define HUFF_SMOOTH_DENSITY_NUM 1u
#define HUFF_SMOOTH_DENSITY_DEN 4u
#endif
BRO_HD_NOINLINE void huff_smooth_counts(uint32_t length, uint32_t* counts) {
  uint32_t used = 0;
  uint64_t total = 0;
  for (uint32_t i = 0; i < length; ++i) { used += counts[i] != 0; total += counts[i]; }
  if (used < 16) return;  // small codes are stored with few bits anyway
  const uint64_t slack = total >> 8 < 16 ? total >> 8 : 16;  // the absolute slack shrinks with the histogram: in a small one every count matters
  while (length != 0 && counts[length - 1] == 0) --length;  // the unused tail of the alphabet is not coded at all
  for (uint32_t i = 0; i < length;) {
    uint64_t sum = counts[i], n = 1, nz = counts[i] != 0;
    uint32_t j = i + 1;
    for (; j < length; ++j) {
      const int64_t dev = (int64_t)counts[j] * (int64_t)n - (int64_t)sum;      // n * (c - mean)
      if ((uint64_t)(dev * dev) > n * (4 * sum + slack * n)) break;             // (c - m)^2 > 4 m + slack
      sum += counts[j];
      nz += counts[j] != 0;
      ++n;
    }
    // an unused symbol is a count of 0 under the same test, so sparse regions (a few symbols seen once or twice) form runs too;
    // they are filled only if at least HUFF_SMOOTH_DENSITY of their symbols are in use -- otherwise zero runs code better
    if (n >= HUFF_SMOOTH_RUN && nz * HUFF_SMOOTH_DENSITY_DEN >= n * HUFF_SMOOTH_DENSITY_NUM) {
      uint32_t mean = (uint32_t)((sum + n / 2) / n);
      if (mean == 0) mean = 1;
      for (uint32_t k = i; k < j; ++k) counts[k] = mean;
    }
    i = j;
  }
}

// used symbols as sort keys (count << 16 | symbol); returns their number
BRO_HD uint32_t huff_collect_keys(const uint32_t* counts, uint32_t length, uint64_t* key) {
  uint32_t n = 0;
  for (uint32_t i = 0; i < length; ++i) if (counts[i]) key[n++] = ((uint64_t)counts[i] << 16) | i;
  return n;
}
BRO_HD_NOINLINE void huff_sort_keys(uint64_t* key, uint32_t n) {  // ascending; shell sort with the 3x+1 gap sequence
  uint32_t gap = 1;
  while (gap < n / 3) gap = 3 * gap + 1;
  for (; gap >= 1; gap /= 3) {
    for (uint32_t i = gap; i < n; ++i) {
      const uint64_t v = key[i];
      uint32_t j = i;
      for (; j >= gap && key[j - gap] > v; j -= gap) key[j] = key[j - gap];
      key[j] = v;
    }
  }
}
// Code lengths of the n >= 2 symbols in ws->key (sorted ascending), at most `limit` bits; depth[] is written for those symbols
// only.
BRO_HD_NOINLINE void huff_lengths_sorted(HuffWs* ws, uint32_t n, uint32_t limit, uint8_t* depth) {
  uint32_t* weight = ws->weight;
  uint16_t* parent = ws->parent;
  uint16_t* nd = ws->node_depth;
  for (uint32_t i = 0; i < n; ++i) weight[i] = (uint32_t)(ws->key[i] >> 16);
  // two-queue merge: leaves 0..n-1 ascending, internal nodes n..2n-2 are created in non-decreasing weight order, so the two
  // smallest nodes are always at the heads of the two queues (a leaf wins a tie: shallower trees)
  uint32_t leaf = 0, inner = n, next = n;
  for (uint32_t k = 0; k + 1 < n; ++k) {
    uint32_t pick[2];
    for (int t = 0; t < 2; ++t) {
      if (leaf < n && (inner >= next || weight[leaf] <= weight[inner])) pick[t] = leaf++;
      else pick[t] = inner++;
    }
    weight[next] = weight[pick[0]] + weight[pick[1]];
    parent[pick[0]] = (uint16_t)next;
    parent[pick[1]] = (uint16_t)next;
    ++next;
  }
  const uint32_t root = next - 1;
  nd[root] = 0;
  uint32_t maxd = 0;
  for (uint32_t v = root; v-- > 0;) {  // a parent is always created after its children: one backward sweep
    nd[v] = (uint16_t)(nd[parent[v]] + 1);
    if (v < n && nd[v] > maxd) maxd = nd[v];
  }
  if (maxd > limit) {
    // Kraft-sum repair.  Clamp to the limit, then lengthen the least frequent symbols that still have room until the code is
    // feasible again (each extra bit on a symbol of length l frees 2^(limit - l - 1) units of 2^-limit), then hand any slack
    // back to the most frequent symbols that can use it.
    int64_t excess = -((int64_t)1 << limit);
    for (uint32_t i = 0; i < n; ++i) {
      if (nd[i] > limit) nd[i] = (uint16_t)limit;
      excess += (int64_t)1 << (limit - nd[i]);
    }
    for (uint32_t i = 0; excess > 0 && i < n;) {
      if (nd[i] >= limit) { ++i; continue; }
      excess -= (int64_t)1 << (limit - nd[i] - 1);
      ++nd[i];
    }
    for (uint32_t i = n; i-- > 0 && excess < 0;) {
      while (nd[i] > 1 && ((int64_t)1 << (limit - nd[i])) <= -excess && (i + 1 == n || nd[i] - 1 >= nd[i + 1])) {
        excess += (int64_t)1 << (limit - nd[i]);
        --nd[i];
      }
    }
  }
  for (uint32_t i = 0; i < n; ++i) depth[ws->key[i] & 0xFFFFu] = (uint8_t)nd[i];
}
// Code lengths of a histogram.  depth[] must be zero on entry (it is written for used symbols only).
BRO_HD_NOINLINE void huff_code_lengths(const uint32_t* counts, uint32_t length, uint32_t limit, HuffWs* ws, uint8_t* depth) {
  const uint32_t n = huff_collect_keys(counts, length, ws->key);
  if (n == 0) return;
  if (n == 1) { depth[ws->key[0] & 0xFFFFu] = 1; return; }
  huff_sort_keys(ws->key, n);
  huff_lengths_sorted(ws, n, limit, depth);
}

BRO_HD uint16_t reverse_bits(uint32_t num_bits, uint32_t bits) {
#ifdef __CUDA_ARCH__
  return (uint16_t)(__brev(bits) >> (32u - num_bits));
#endif
  uint32_t r = 0;
  for (uint32_t i = 0; i < num_bits; ++i) {
    r = (r << 1) | (bits & 1u);
    bits >>= 1;
  }
  return (uint16_t)r;
}
// canonical codes (RFC 7932 section 3.2), bit-reversed because the stream is written LSB first
BRO_HD_NOINLINE void huff_depths_to_codes(const uint8_t* depth, uint32_t len, uint16_t* codes) {
  uint32_t per_len[16], next_code[16];
  for (int l = 0; l < 16; ++l) per_len[l] = 0;
  for (uint32_t i = 0; i < len; ++i) ++per_len[depth[i]];
  uint32_t code = 0;
  per_len[0] = 0;
  for (int l = 1; l < 16; ++l) {
    code = (code + per_len[l - 1]) << 1;
    next_code[l] = code;
  }
  for (uint32_t i = 0; i < len; ++i) {
    const uint32_t l = depth[i];
    if (l) codes[i] = reverse_bits(l, next_code[l]++);
  }
}

// ---- run-length form of a code length sequence ----
// A run of r repeats written with k compounding repeat symbols covers r - 2 = sum d_j B^(k-j) with digits d_j in 1..B (B = 4 for
// symbol 16, 8 for symbol 17): bijective base-B numeration; the extra bits of a symbol are d_j - 1.
BRO_HD uint32_t huff_repeat_symbols(uint32_t r, uint32_t B) {  // symbols needed for r >= 3 repeats
  uint32_t k = 0;
  for (uint32_t x = r - 2; x > 0; x = (x - 1) / B) ++k;  // (x - d) / B with d = ((x - 1) % B) + 1
  return k;
}
BRO_HD void huff_put_repeats(uint32_t r, uint32_t B, uint8_t symbol, uint32_t* n, uint8_t* sym, uint8_t* extra) {
  const uint32_t k = huff_repeat_symbols(r, B);
  uint32_t x = r - 2;
  for (uint32_t j = k; j-- > 0;) {  // least significant digit is written last
    const uint32_t d = ((x - 1) % B) + 1;
    sym[*n + j] = symbol;
    extra[*n + j] = (uint8_t)(d - 1);
    x = (x - d) / B;
  }
  *n += k;
}
// Emits `r` further occurrences of the current value (the repeat symbol repeats `value`): as literals when that is not longer,
// else as one compound repeat, or as a literal followed by a compound repeat of r - 1 when the shorter run needs one symbol less.
BRO_HD void huff_put_run(uint8_t value, uint32_t r, uint32_t* n, uint8_t* sym, uint8_t* extra) {
  const uint32_t B = value ? 4u : 8u;
  const uint8_t rep = value ? 16 : 17;
  while (r > 0) {
    if (r < 3) { sym[*n] = value; extra[*n] = 0; ++*n; --r; continue; }
    if (r > 3 && huff_repeat_symbols(r - 1, B) < huff_repeat_symbols(r, B)) { sym[*n] = value; extra[*n] = 0; ++*n; --r; }
    huff_put_repeats(r, B, rep, n, sym, extra);
    r = 0;
  }
}
BRO_HD_NOINLINE uint32_t huff_rle_lengths(const uint8_t* depth, uint32_t length, uint8_t* sym, uint8_t* extra) {
  while (length > 0 && depth[length - 1] == 0) --length;  // trailing unused symbols are implied
  uint32_t n = 0;
  uint8_t prev_nonzero = 8;  // the format's initial "previous length"
  for (uint32_t i = 0; i < length;) {
    const uint8_t v = depth[i];
    uint32_t r = 1;
    while (i + r < length && depth[i + r] == v) ++r;
    i += r;
    if (v != 0 && v != prev_nonzero) {  // a repeat symbol repeats the last non-zero length: a new value needs one literal first
      sym[n] = v; extra[n] = 0; ++n;
      --r;
      prev_nonzero = v;
    }
    huff_put_run(v, r, &n, sym, extra);
  }
  return n;
}

// Complex prefix code description (RFC 7932 section 3.5): the code-length code (lengths in the format's storage order, with
// its fixed variable-length code), then the run-length coded lengths.
BRO_HD_NOINLINE void huff_store_complex(BitWriter& bw, const uint8_t* depths, uint32_t num, HuffWs* ws) {
  uint8_t cl_depth[18];
  uint16_t cl_bits[18];
  uint32_t cl_hist[18];
  for (int i = 0; i < 18; ++i) { cl_depth[i] = 0; cl_bits[i] = 0; cl_hist[i] = 0; }
  const uint32_t nsym = huff_rle_lengths(depths, num, ws->rle_sym, ws->rle_extra);
  for (uint32_t i = 0; i < nsym; ++i) ++cl_hist[ws->rle_sym[i]];
  uint32_t used = 0, only = 0;
  for (uint32_t i = 0; i < 18; ++i) if (cl_hist[i]) { if (!used) only = i; ++used; }
  huff_code_lengths(cl_hist, 18, 5, ws, cl_depth);
  huff_depths_to_codes(cl_depth, 18, cl_bits);
  {
    static constexpr uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};  // format: storage order
    static constexpr uint8_t kVlcBits[6] = {0, 7, 3, 2, 1, 15};                                            // format: fixed code of a length 0..5
    static constexpr uint8_t kVlcLen[6] = {2, 4, 3, 2, 2, 4};
    uint32_t last = 18;  // lengths after the last non-zero one are implied
    if (used > 1) while (last > 0 && cl_depth[kOrder[last - 1]] == 0) --last;
    uint32_t skip = 0;   // HSKIP: the first two or three (zero) lengths can be skipped
    if (cl_depth[kOrder[0]] == 0 && cl_depth[kOrder[1]] == 0) skip = cl_depth[kOrder[2]] == 0 ? 3u : 2u;
    bw.put(2, skip);
    for (uint32_t i = skip; i < last; ++i) {
      const uint32_t l = cl_depth[kOrder[i]];
      bw.put(kVlcLen[l], kVlcBits[l]);
    }
  }
  if (used == 1) cl_depth[only] = 0;  // a code with a single symbol takes no bits
  for (uint32_t i = 0; i < nsym; ++i) {
    const uint32_t s = ws->rle_sym[i];
    bw.put(cl_depth[s], cl_bits[s]);
    if (s == 16) bw.put(2, ws->rle_extra[i]);
    else if (s == 17) bw.put(3, ws->rle_extra[i]);
  }
}

// Builds the code of a histogram (depth / codes get histogram_length entries) and appends its description to bw:
// 0 or 1 used symbols -> the one-symbol form, <= 4 -> the simple form, else the complex form (RFC 7932 sections 3.4, 3.5).
BRO_HD_NOINLINE void huff_build_and_store(BitWriter& bw, const uint32_t* histogram, uint32_t histogram_length,
                                          uint32_t alphabet_size, HuffWs* ws, uint8_t* depth, uint16_t* codes) {
  uint32_t used = 0, first4[4] = {0, 0, 0, 0}, sym_bits = 0;
  for (uint32_t i = 0; i < histogram_length && used < 5; ++i)
    if (histogram[i]) { if (used < 4) first4[used] = i; ++used; }
  for (uint32_t c = alphabet_size - 1; c; c >>= 1) ++sym_bits;
  for (uint32_t i = 0; i < histogram_length; ++i) { depth[i] = 0; codes[i] = 0; }
  if (used <= 1) {
    bw.put(4, 1);  // simple code, one symbol
    bw.put(sym_bits, first4[0]);
    return;
  }
  huff_code_lengths(histogram, histogram_length, 15, ws, depth);
  huff_depths_to_codes(depth, histogram_length, codes);
  if (used > 4) { huff_store_complex(bw, depth, histogram_length, ws); return; }
  // simple code: the symbols in order of code length (the lengths themselves are implied by NSYM and the tree-select bit)
  for (uint32_t i = 1; i < used; ++i) {
    const uint32_t v = first4[i];
    uint32_t j = i;
    for (; j > 0 && depth[first4[j - 1]] > depth[v]; --j) first4[j] = first4[j - 1];
    first4[j] = v;
  }
  bw.put(2, 1);
  bw.put(2, used - 1);
  for (uint32_t i = 0; i < used; ++i) bw.put(sym_bits, first4[i]);
  if (used == 4) bw.put(1, depth[first4[0]] == 1 ? 1u : 0u);
}

}  // namespace bro
