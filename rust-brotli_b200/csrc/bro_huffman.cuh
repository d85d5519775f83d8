// bro_huffman.cuh -- prefix-code construction and serialisation, one histogram per GPU thread.
// Semantics follow the reference: entropy_encode.rs:27-56,71-116,133-210 (length-limited tree by
// count clamping), :211-345 (count smoothing for RLE), :347-525 (code-length RLE), :546-575 (canonical
// codes); brotli_bit_stream.rs:764-911,1401-1498 (serialisation).
#pragma once
#include "bro_common.cuh"

namespace bro {

struct HuffNode {
  uint32_t count;
  int16_t left;
  int16_t right_or_value;
};
// workspace: at least 2 * n + 2 nodes

BRO_HD bool huff_sort_less(const HuffNode& a, const HuffNode& b) {
  if (a.count != b.count) return a.count < b.count;
  return a.right_or_value > b.right_or_value;
}

BRO_HD_NOINLINE bool huff_set_depth(int p0, HuffNode* pool, uint8_t* depth, int max_depth) {
  int stack[16];
  int level = 0;
  int p = p0;
  stack[0] = -1;
  for (;;) {
    if (pool[p].left >= 0) {
      level++;
      if (level > max_depth) return false;
      stack[level] = pool[p].right_or_value;
      p = pool[p].left;
      continue;
    } else {
      depth[pool[p].right_or_value] = (uint8_t)level;
    }
    while (level >= 0 && stack[level] == -1) level--;
    if (level < 0) return true;
    p = stack[level];
    stack[level] = -1;
  }
}

BRO_HD_NOINLINE void huff_sort(HuffNode* items, uint32_t n) {
  static constexpr uint32_t gaps[6] = {132, 57, 23, 10, 4, 1};
  if (n < 13) {
    for (uint32_t i = 1; i < n; ++i) {
      HuffNode tmp = items[i];
      uint32_t k = i, j = i - 1;
      while (huff_sort_less(tmp, items[j])) {
        items[k] = items[j];
        k = j;
        if (!j--) break;
      }
      items[k] = tmp;
    }
  } else {
    for (int g = n < 57 ? 2 : 0; g < 6; ++g) {
      uint32_t gap = gaps[g];
      for (uint32_t i = gap; i < n; ++i) {
        uint32_t j = i;
        HuffNode tmp = items[i];
        for (; j >= gap && huff_sort_less(tmp, items[j - gap]); j -= gap) items[j] = items[j - gap];
        items[j] = tmp;
      }
    }
  }
}

// depth[] must be zero for unused symbols on entry (it is only written for used ones).
BRO_HD_NOINLINE void huff_create_tree(const uint32_t* data, uint32_t length, int tree_limit, HuffNode* tree,
                                      uint8_t* depth) {
  HuffNode sentinel;
  sentinel.count = 0xFFFFFFFFu;
  sentinel.left = -1;
  sentinel.right_or_value = -1;
  for (uint32_t count_limit = 1;; count_limit *= 2) {
    uint32_t n = 0;
    for (uint32_t i = length; i != 0;) {
      --i;
      if (data[i]) {
        tree[n].count = bmax(data[i], count_limit);
        tree[n].left = -1;
        tree[n].right_or_value = (int16_t)i;
        ++n;
      }
    }
    if (n == 1) {
      depth[tree[0].right_or_value] = 1;
      return;
    }
    huff_sort(tree, n);
    tree[n] = sentinel;
    tree[n + 1] = sentinel;
    uint32_t i = 0, j = n + 1;
    for (uint32_t k = n - 1; k != 0; --k) {
      uint32_t left, right;
      if (tree[i].count <= tree[j].count) left = i++; else left = j++;
      if (tree[i].count <= tree[j].count) right = i++; else right = j++;
      uint32_t j_end = 2 * n - k;
      tree[j_end].count = tree[left].count + tree[right].count;
      tree[j_end].left = (int16_t)left;
      tree[j_end].right_or_value = (int16_t)right;
      tree[j_end + 1] = sentinel;
    }
    if (huff_set_depth((int)(2 * n - 1), tree, depth, tree_limit)) return;
  }
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
BRO_HD_NOINLINE void huff_depths_to_codes(const uint8_t* depth, uint32_t len, uint16_t* codes) {
  uint16_t bl_count[16], next_code[16];
  for (int i = 0; i < 16; ++i) bl_count[i] = 0;
  for (uint32_t i = 0; i < len; ++i) bl_count[depth[i]]++;
  bl_count[0] = 0;
  next_code[0] = 0;
  int code = 0;
  for (int i = 1; i < 16; ++i) {
    code = (code + bl_count[i - 1]) << 1;
    next_code[i] = (uint16_t)code;
  }
  for (uint32_t i = 0; i < len; ++i)
    if (depth[i]) codes[i] = reverse_bits(depth[i], next_code[depth[i]]++);
}

// entropy_encode.rs:211-345; good_for_rle: workspace of `length` bytes
BRO_HD_NOINLINE void huff_optimize_counts_for_rle(uint32_t length, uint32_t* counts, uint8_t* good_for_rle) {
  uint32_t nonzero_count = 0;
  const uint32_t streak_limit = 1240;
  for (uint32_t i = 0; i < length; ++i) if (counts[i]) ++nonzero_count;
  if (nonzero_count < 16) return;
  while (length != 0 && counts[length - 1] == 0) --length;
  if (length == 0) return;
  {
    uint32_t nonzeros = 0, smallest_nonzero = 1u << 30;
    for (uint32_t i = 0; i < length; ++i) {
      if (counts[i] != 0) {
        ++nonzeros;
        if (smallest_nonzero > counts[i]) smallest_nonzero = counts[i];
      }
    }
    if (nonzeros < 5) return;
    if (smallest_nonzero < 4) {
      uint32_t zeros = length - nonzeros;
      if (zeros < 6)
        for (uint32_t i = 1; i + 1 < length; ++i)
          if (counts[i - 1] != 0 && counts[i] == 0 && counts[i + 1] != 0) counts[i] = 1;
    }
    if (nonzeros < 28) return;
  }
  for (uint32_t i = 0; i < length; ++i) good_for_rle[i] = 0;
  {
    uint32_t symbol = counts[0];
    uint32_t step = 0;
    for (uint32_t i = 0; i <= length; ++i) {
      if (i == length || counts[i] != symbol) {
        if ((symbol == 0 && step >= 5) || (symbol != 0 && step >= 7))
          for (uint32_t k = 0; k < step; ++k) good_for_rle[i - k - 1] = 1;
        step = 1;
        if (i != length) symbol = counts[i];
      } else {
        ++step;
      }
    }
  }
  uint64_t stride = 0, sum = 0;
  uint64_t limit = 256ull * ((uint64_t)counts[0] + counts[1] + counts[2]) / 3 + 420;
  for (uint32_t i = 0; i <= length; ++i) {
    bool brk = (i == length) || good_for_rle[i] || (i != 0 && good_for_rle[i - 1]);
    if (!brk) brk = (256ull * counts[i] - limit + streak_limit) >= 2ull * streak_limit;  // unsigned wrap intended
    if (brk) {
      if (stride >= 4 || (stride >= 3 && sum == 0)) {
        uint64_t count = (sum + stride / 2) / stride;
        if (count == 0) count = 1;
        if (sum == 0) count = 0;
        for (uint32_t k = 0; k < stride; ++k) counts[i - k - 1] = (uint32_t)count;
      }
      stride = 0;
      sum = 0;
      if (i + 2 < length) limit = 256ull * ((uint64_t)counts[i] + counts[i + 1] + counts[i + 2]) / 3 + 420;
      else if (i < length) limit = 256ull * counts[i];
      else limit = 0;
    }
    ++stride;
    if (i != length) {
      sum += counts[i];
      if (stride >= 4) limit = (256ull * sum + stride / 2) / stride;
      if (stride == 4) limit += 120;
    }
  }
}

// ---- code-length RLE (entropy_encode.rs:347-525) ----
BRO_HD void cl_reverse(uint8_t* v, uint32_t start, uint32_t end) {
  --end;
  while (start < end) {
    uint8_t t = v[start];
    v[start] = v[end];
    v[end] = t;
    ++start;
    --end;
  }
}
BRO_HD_NOINLINE void cl_write_reps(uint8_t prev, uint8_t value, uint32_t reps, uint32_t* n, uint8_t* tree, uint8_t* extra) {
  if (prev != value) { tree[*n] = value; extra[*n] = 0; ++*n; --reps; }
  if (reps == 7) { tree[*n] = value; extra[*n] = 0; ++*n; --reps; }
  if (reps < 3) {
    for (uint32_t i = 0; i < reps; ++i) { tree[*n] = value; extra[*n] = 0; ++*n; }
  } else {
    uint32_t start = *n;
    reps -= 3;
    for (;;) {
      tree[*n] = 16; extra[*n] = (uint8_t)(reps & 3); ++*n;
      reps >>= 2;
      if (reps == 0) break;
      --reps;
    }
    cl_reverse(tree, start, *n);
    cl_reverse(extra, start, *n);
  }
}
BRO_HD_NOINLINE void cl_write_zero_reps(uint32_t reps, uint32_t* n, uint8_t* tree, uint8_t* extra) {
  if (reps == 11) { tree[*n] = 0; extra[*n] = 0; ++*n; --reps; }
  if (reps < 3) {
    for (uint32_t i = 0; i < reps; ++i) { tree[*n] = 0; extra[*n] = 0; ++*n; }
  } else {
    uint32_t start = *n;
    reps -= 3;
    for (;;) {
      tree[*n] = 17; extra[*n] = (uint8_t)(reps & 7); ++*n;
      reps >>= 3;
      if (reps == 0) break;
      --reps;
    }
    cl_reverse(tree, start, *n);
    cl_reverse(extra, start, *n);
  }
}
BRO_HD_NOINLINE void cl_write_tree(const uint8_t* depth, uint32_t length, uint32_t* n, uint8_t* tree, uint8_t* extra) {
  uint8_t previous_value = 8;
  bool use_nz = false, use_z = false;
  uint32_t new_length = length;
  for (uint32_t i = 0; i < length; ++i) {
    if (depth[length - i - 1] == 0) --new_length; else break;
  }
  if (length > 50) {  // decide_over_rle_use
    uint32_t total_reps_zero = 0, total_reps_non_zero = 0, count_reps_zero = 1, count_reps_non_zero = 1;
    for (uint32_t i = 0; i < new_length;) {
      uint8_t value = depth[i];
      uint32_t reps = 1;
      for (uint32_t k = i + 1; k < new_length && depth[k] == value; ++k) ++reps;
      if (reps >= 3 && value == 0) { total_reps_zero += reps; ++count_reps_zero; }
      if (reps >= 4 && value != 0) { total_reps_non_zero += reps; ++count_reps_non_zero; }
      i += reps;
    }
    use_nz = total_reps_non_zero > count_reps_non_zero * 2;
    use_z = total_reps_zero > count_reps_zero * 2;
  }
  for (uint32_t i = 0; i < new_length;) {
    uint8_t value = depth[i];
    uint32_t reps = 1;
    if ((value != 0 && use_nz) || (value == 0 && use_z))
      for (uint32_t k = i + 1; k < new_length && depth[k] == value; ++k) ++reps;
    if (value == 0) cl_write_zero_reps(reps, n, tree, extra);
    else { cl_write_reps(previous_value, value, reps, n, tree, extra); previous_value = value; }
    i += reps;
  }
}

// workspace for serialising one code: RLE symbols + extra bits
struct HuffStoreWs {
  uint8_t rle[704];
  uint8_t extra[704];
  HuffNode nodes[2 * 704 + 2];
};

BRO_HD_NOINLINE void huff_store_complex(BitWriter& bw, const uint8_t* depths, uint32_t num, HuffStoreWs* ws) {
  uint32_t tree_size = 0;
  uint8_t cl_depth[18];
  uint16_t cl_bits[18];
  uint32_t histogram[18];
  for (int i = 0; i < 18; ++i) { cl_depth[i] = 0; cl_bits[i] = 0; histogram[i] = 0; }
  cl_write_tree(depths, num, &tree_size, ws->rle, ws->extra);
  for (uint32_t i = 0; i < tree_size; ++i) ++histogram[ws->rle[i]];
  int num_codes = 0;
  uint32_t code = 0;
  for (uint32_t i = 0; i < 18; ++i) {
    if (histogram[i]) {
      if (num_codes == 0) { code = i; num_codes = 1; }
      else if (num_codes == 1) { num_codes = 2; break; }
    }
  }
  huff_create_tree(histogram, 18, 5, ws->nodes, cl_depth);
  huff_depths_to_codes(cl_depth, 18, cl_bits);
  {  // brotli_bit_stream.rs:764-808
    static constexpr uint8_t kStorageOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    static constexpr uint8_t kSymbols[6] = {0, 7, 3, 2, 1, 15};
    static constexpr uint8_t kLengths[6] = {2, 4, 3, 2, 2, 4};
    uint32_t skip_some = 0, codes_to_store = 18;
    if (num_codes > 1)
      for (; codes_to_store > 0; --codes_to_store)
        if (cl_depth[kStorageOrder[codes_to_store - 1]] != 0) break;
    if (cl_depth[kStorageOrder[0]] == 0 && cl_depth[kStorageOrder[1]] == 0) {
      skip_some = 2;
      if (cl_depth[kStorageOrder[2]] == 0) skip_some = 3;
    }
    bw.put(2, skip_some);
    for (uint32_t i = skip_some; i < codes_to_store; ++i) {
      uint32_t l = cl_depth[kStorageOrder[i]];
      bw.put(kLengths[l], kSymbols[l]);
    }
  }
  if (num_codes == 1) cl_depth[code] = 0;
  for (uint32_t i = 0; i < tree_size; ++i) {
    uint32_t s = ws->rle[i];
    bw.put(cl_depth[s], cl_bits[s]);
    if (s == 16) bw.put(2, ws->extra[i]);
    else if (s == 17) bw.put(3, ws->extra[i]);
  }
}

// brotli_bit_stream.rs:1445-1498.  depth/codes arrays have histogram_length entries and are fully written.
BRO_HD_NOINLINE void huff_build_and_store(BitWriter& bw, const uint32_t* histogram, uint32_t histogram_length,
                                          uint32_t alphabet_size, HuffStoreWs* ws, uint8_t* depth, uint16_t* codes) {
  uint32_t count = 0, s4[4] = {0, 0, 0, 0}, max_bits = 0;
  for (uint32_t i = 0; i < histogram_length; ++i) {
    if (histogram[i]) {
      if (count < 4) s4[count] = i;
      else if (count > 4) break;
      count++;
    }
  }
  for (uint32_t c = alphabet_size - 1; c; c >>= 1) ++max_bits;
  for (uint32_t i = 0; i < histogram_length; ++i) { depth[i] = 0; codes[i] = 0; }
  if (count <= 1) {
    bw.put(4, 1);
    bw.put(max_bits, s4[0]);
    return;
  }
  huff_create_tree(histogram, histogram_length, 15, ws->nodes, depth);
  huff_depths_to_codes(depth, histogram_length, codes);
  if (count <= 4) {  // StoreSimpleHuffmanTree brotli_bit_stream.rs:1401-1443
    bw.put(2, 1);
    bw.put(2, count - 1);
    for (uint32_t i = 0; i < count; ++i)
      for (uint32_t j = i + 1; j < count; ++j)
        if (depth[s4[j]] < depth[s4[i]]) { uint32_t t = s4[j]; s4[j] = s4[i]; s4[i] = t; }
    for (uint32_t i = 0; i < count; ++i) bw.put(max_bits, s4[i]);
    if (count == 4) bw.put(1, depth[s4[0]] == 1 ? 1u : 0u);
  } else {
    huff_store_complex(bw, depth, histogram_length, ws);
  }
}

}  // namespace bro
