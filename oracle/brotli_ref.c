/*
 * oracle/brotli_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of the reference's (dropbox/rust-brotli 8.0.4) compression hot path for
 * qualities 4..9 with the hash-chain hashers H5 / H6 / H9, the greedy+lazy LZ77 parse, the greedy
 * (streaming) block splitter with static literal contexts, Huffman construction and the bit-stream
 * writer.  Every function cites the reference file:line it follows.  It exists so that
 *   (a) the CUDA path has a per-stage CPU checker (hash keys, command codes, entropy, Huffman depths),
 *   (b) compressed SIZE parity can be asserted against a faithful model of the reference encoder,
 *   (c) bench.py has a CPU baseline ("port") to time beside the GPU.
 *
 * Pin status: the Rust reference cannot be built in this image (no cargo).  This restatement is pinned
 * by (1) round-tripping every stream through the system RFC 7932 decoder (libbrotlidec 1.1.0), and
 * (2) compressed-size agreement with Google's C encoder libbrotlienc 1.1.0 (the code the reference was
 * ported from) -- see tests/test_oracle.py and tests/golden/.  The reference's own in-tree KATs that this
 * path can reach are checked there too.  Not restated: static-dictionary matching
 * (backward_references/mod.rs:1896-1988) -- SURVEY.md section 8(f) row 4.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINZ(a, b) ((a) < (b) ? (a) : (b))
#define MAXZ(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------------
 * log2 helpers: util.rs:13-25 (FastLog2u16 = logs_16 table, FastLog2 = logs_8 table below 256, else
 * f32::log2), floatX = f32 (util.rs:9-10).
 * ---------------------------------------------------------------------------------------------- */
static float g_log2_u16[65536];
static int g_tables_ready = 0;
static void init_tables(void) {
  if (g_tables_ready) return;
  g_log2_u16[0] = 0.0f;
  for (int i = 1; i < 65536; ++i) g_log2_u16[i] = (float)log2((double)i);
  g_tables_ready = 1;
}
static inline float FastLog2u16(uint16_t v) { return g_log2_u16[v]; }
static inline float FastLog2(uint64_t v) { return v < 256 ? g_log2_u16[v] : log2f((float)v); }
static inline uint32_t Log2FloorNonZero(uint64_t n) { return 63u - (uint32_t)__builtin_clzll(n); }

/* ------------------------------------------------------------------------------------------------
 * Format constants (RFC 7932; constants.rs:2-18, brotli_bit_stream.rs:635).
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t kInsBase[24] = {0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98,
                                      130, 194, 322, 578, 1090, 2114, 6210, 22594};
static const uint32_t kInsExtra[24] = {0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24};
static const uint32_t kCopyBase[24] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54,
                                       70, 102, 134, 198, 326, 582, 1094, 2118};
static const uint32_t kCopyExtra[24] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24};
static const uint32_t kBlockLenOffset[26] = {1, 5, 9, 13, 17, 25, 33, 41, 49, 65, 81, 97, 113, 145, 177, 209,
                                             241, 305, 369, 497, 753, 1265, 2289, 4337, 8433, 16625};
static const uint32_t kBlockLenNBits[26] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5,
                                            6, 6, 7, 8, 9, 10, 11, 12, 13, 24};

/* UTF8 literal context lookup (RFC 7932 section 7.1; constants.rs:228).  lut0 is indexed by the previous
 * byte, lut1 by the byte before it; context id = lut0[p1] | lut1[p2]. */
static uint8_t g_utf8_lut0[256], g_utf8_lut1[256];
static void init_context_luts(void) {
  static const uint8_t ascii0[128] = {
      0,  0,  0,  0,  0,  0,  0,  0,  0,  4,  4,  0,  0,  4,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,
      0,  0,  0,  0,  0,  0,  8,  12, 16, 12, 12, 20, 12, 16, 24, 28, 12, 12, 32, 12, 36, 12, 44, 44, 44, 44,
      44, 44, 44, 44, 44, 44, 32, 32, 24, 40, 28, 12, 12, 48, 52, 52, 52, 48, 52, 52, 52, 48, 52, 52, 52, 52,
      52, 48, 52, 52, 52, 52, 52, 48, 52, 52, 52, 52, 52, 24, 12, 28, 12, 12, 12, 56, 60, 60, 60, 56, 60, 60,
      60, 56, 60, 60, 60, 60, 60, 56, 60, 60, 60, 60, 60, 56, 60, 60, 60, 60, 60, 24, 12, 28, 12, 0};
  for (int c = 0; c < 256; ++c) {
    if (c < 128) g_utf8_lut0[c] = ascii0[c];
    else if (c < 192) g_utf8_lut0[c] = (uint8_t)(c & 1);
    else g_utf8_lut0[c] = (uint8_t)(2 + (c & 1));
    uint8_t v;
    if (c < 32) v = 0;
    else if (c < 128) {
      if (c == 32 || c == 127) v = 0;
      else if (c >= '0' && c <= '9') v = 2;
      else if (c >= 'A' && c <= 'Z') v = 2;
      else if (c >= 'a' && c <= 'z') v = 3;
      else v = 1;
    } else if (c < 224) v = 0;
    else v = 2;
    g_utf8_lut1[c] = v;
  }
}
static inline uint8_t ContextUTF8(uint8_t p1, uint8_t p2) { return g_utf8_lut0[p1] | g_utf8_lut1[p2]; }

/* ------------------------------------------------------------------------------------------------
 * Bit writer: brotli_bit_stream.rs:742-757 (LSB-first, up to 56 bits per call).
 * ---------------------------------------------------------------------------------------------- */
static inline void WriteBits(unsigned n_bits, uint64_t bits, size_t* pos, uint8_t* array) {
  uint8_t* p = &array[*pos >> 3];
  uint64_t v = (uint64_t)p[0];
  v |= bits << (*pos & 7);
  for (int i = 0; i < 8; ++i) p[i] = (uint8_t)(v >> (8 * i));
  *pos += n_bits;
}
static inline void JumpToByteBoundary(size_t* pos, uint8_t* array) {
  *pos = (*pos + 7u) & ~(size_t)7u;
  array[*pos >> 3] = 0;
}

/* ------------------------------------------------------------------------------------------------
 * Command record and prefix codes: command.rs:11-21, :48-68, :71-121, :134-173, :273-297.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t insert_len_;
  uint32_t copy_len_;   /* low 25 bits length, high 7 bits (copy_code - copy_len) */
  uint32_t dist_extra_;
  uint16_t cmd_prefix_;
  uint16_t dist_prefix_; /* low 10 bits symbol, high 6 bits number of extra bits */
} Command;

static inline uint32_t CommandCopyLen(const Command* c) { return c->copy_len_ & 0x01ffffffu; }
static inline uint32_t CommandCopyLenCode(const Command* c) { /* brotli_bit_stream.rs:1924 */
  uint32_t modifier = c->copy_len_ >> 25;
  int32_t delta = (int8_t)(uint8_t)(modifier | ((modifier & 0x40u) << 1));
  return (uint32_t)((int32_t)(c->copy_len_ & 0x01ffffffu) + delta);
}

size_t oracle_compute_distance_code(size_t distance, size_t max_distance, const int* dist_cache) {
  if (distance <= max_distance) { /* command.rs:48-68 */
    size_t distance_plus_3 = distance + 3;
    size_t offset0 = distance_plus_3 - (size_t)dist_cache[0];
    size_t offset1 = distance_plus_3 - (size_t)dist_cache[1];
    if (distance == (size_t)dist_cache[0]) return 0;
    if (distance == (size_t)dist_cache[1]) return 1;
    if (offset0 < 7) return (0x09750468 >> (4 * offset0)) & 0xF;
    if (offset1 < 7) return (0x0FDB1ACE >> (4 * offset1)) & 0xF;
    if (distance == (size_t)dist_cache[2]) return 2;
    if (distance == (size_t)dist_cache[3]) return 3;
  }
  return distance + 16 - 1;
}

uint16_t oracle_insert_length_code(size_t insertlen) { /* command.rs:71-88 */
  if (insertlen < 6) return (uint16_t)insertlen;
  if (insertlen < 130) {
    uint32_t nbits = Log2FloorNonZero(insertlen - 2) - 1u;
    return (uint16_t)((nbits << 1) + ((insertlen - 2) >> nbits) + 2);
  }
  if (insertlen < 2114) return (uint16_t)(Log2FloorNonZero(insertlen - 66) + 10);
  if (insertlen < 6210) return 21;
  if (insertlen < 22594) return 22;
  return 23;
}
uint16_t oracle_copy_length_code(size_t copylen) { /* command.rs:91-104 */
  if (copylen < 10) return (uint16_t)(copylen - 2);
  if (copylen < 134) {
    uint32_t nbits = Log2FloorNonZero(copylen - 6) - 1u;
    return (uint16_t)((nbits << 1) + ((copylen - 6) >> nbits) + 4);
  }
  if (copylen < 2118) return (uint16_t)(Log2FloorNonZero(copylen - 70) + 12);
  return 23;
}
uint16_t oracle_combine_length_codes(uint16_t inscode, uint16_t copycode, int use_last_distance) {
  uint16_t bits64 = (uint16_t)((copycode & 0x7u) | ((inscode & 0x7u) << 3)); /* command.rs:107-121 */
  if (use_last_distance && inscode < 8 && copycode < 16) return (copycode < 8) ? bits64 : (uint16_t)(bits64 | 64);
  int sub_offset = 2 * ((copycode >> 3) + 3 * (inscode >> 3));
  int offset = (sub_offset << 5) + 0x40 + ((0x520D40 >> sub_offset) & 0xC0);
  return (uint16_t)(offset | bits64);
}
static inline uint16_t GetLengthCode(size_t insertlen, size_t copylen, int use_last_distance) {
  return oracle_combine_length_codes(oracle_insert_length_code(insertlen), oracle_copy_length_code(copylen),
                                     use_last_distance);
}
/* command.rs:134-173 with NPOSTFIX = 0, NDIRECT = 0 (encode.rs:2169-2190 for non-FONT modes). */
void oracle_prefix_encode_copy_distance(size_t distance_code, uint16_t* code, uint32_t* extra_bits) {
  if (distance_code < 16) {
    *code = (uint16_t)distance_code;
    *extra_bits = 0;
  } else {
    uint64_t dist = (1ull << 2) + (distance_code - 16);
    uint64_t bucket = Log2FloorNonZero(dist) - 1u;
    uint64_t prefix = (dist >> bucket) & 1;
    uint64_t offset = (2 + prefix) << bucket;
    uint64_t nbits = bucket;
    *code = (uint16_t)((nbits << 10) | (16 + 2 * (nbits - 1) + prefix));
    *extra_bits = (uint32_t)(dist - offset);
  }
}
static void CommandInit(Command* c, size_t insertlen, size_t copylen, size_t copylen_code, size_t distance_code) {
  c->insert_len_ = (uint32_t)insertlen; /* command.rs:273-297 */
  int8_t delta = (int8_t)((int)copylen_code - (int)copylen);
  c->copy_len_ = (uint32_t)copylen | ((uint32_t)(uint8_t)delta << 25);
  oracle_prefix_encode_copy_distance(distance_code, &c->dist_prefix_, &c->dist_extra_);
  c->cmd_prefix_ = GetLengthCode(insertlen, copylen_code, (c->dist_prefix_ & 0x3ff) == 0);
}
static void CommandInitInsert(Command* c, size_t insertlen) { /* command.rs:38-44 */
  c->insert_len_ = (uint32_t)insertlen;
  c->copy_len_ = 4u << 25;
  c->dist_extra_ = 0;
  c->dist_prefix_ = (1u << 10) | 16;
  c->cmd_prefix_ = GetLengthCode(insertlen, 4, 0);
}
static uint32_t CommandRestoreDistanceCode(const Command* c) { /* command.rs:176-200, npostfix = ndirect = 0 */
  uint32_t dcode = c->dist_prefix_ & 0x3ffu;
  if (dcode < 16) return dcode;
  uint32_t nbits = c->dist_prefix_ >> 10;
  uint32_t hcode = dcode - 16;
  uint32_t offset = ((2u + (hcode & 1u)) << nbits) - 4u;
  return offset + c->dist_extra_ + 16;
}
/* test hook: full Command::init (command.rs:273) */
void oracle_command_init(size_t insertlen, size_t copylen, size_t distance_code, uint32_t out[5]) {
  Command c;
  CommandInit(&c, insertlen, copylen, copylen, distance_code);
  out[0] = c.insert_len_; out[1] = c.copy_len_; out[2] = c.dist_extra_; out[3] = c.cmd_prefix_; out[4] = c.dist_prefix_;
}

/* ------------------------------------------------------------------------------------------------
 * Hashers.  AdvHasher H5 / H6: backward_references/mod.rs:919-1149 (specialisations), :1470-1813.
 * H9: backward_references/mod.rs:598-917.  Hasher choice: encode.rs:834-893.
 * ---------------------------------------------------------------------------------------------- */
#define kHashMul32 0x1e35a7bdu
#define kHashMul64Long 0x1fe35a7bd3579bd3ull

typedef struct {
  int type; /* 5, 6 or 9 */
  int bucket_bits, block_bits, hash_len, n_last;
  uint32_t block_mask;
  uint64_t hash_mask;
  uint16_t* num;
  uint32_t* buckets;
  int use_dictionary;                            /* params.use_dictionary (encode.rs:559-562: off for catable) */
  size_t dict_num_lookups, dict_num_matches;     /* Struct1 of the hasher (mod.rs:161-169) */
} Hasher;

typedef struct {
  size_t len, len_x_code, distance;
  uint64_t score;
} SearchResult;

static inline uint32_t HashBytes(const Hasher* h, const uint8_t* p) {
  if (h->type == 6) { /* mod.rs:1138-1140, shift 64 - bucket_bits (encode.rs:1066-1067) */
    uint64_t v;
    memcpy(&v, p, 8);
    return (uint32_t)(((v & h->hash_mask) * kHashMul64Long) >> (64 - h->bucket_bits));
  }
  uint32_t v; /* mod.rs:990-991 (H5), mod.rs:610-613 (H9): load32 * kHashMul32 >> (32 - bucket_bits) */
  memcpy(&v, p, 4);
  return (uint32_t)(v * kHashMul32) >> (32 - h->bucket_bits);
}
/* test hook: hash keys for n positions (buffer must have 8 readable bytes past the last position) */
void oracle_hash_keys(int type, int bucket_bits, int hash_len, const uint8_t* data, size_t n, uint32_t* keys) {
  Hasher h;
  memset(&h, 0, sizeof(h));
  h.type = type; h.bucket_bits = bucket_bits; h.hash_len = hash_len;
  h.hash_mask = hash_len >= 8 ? ~0ull : (~0ull >> (64 - 8 * hash_len));
  for (size_t i = 0; i < n; ++i) keys[i] = HashBytes(&h, data + i);
}

static inline size_t HashTypeLength(const Hasher* h) { return h->type == 6 ? 8 : 4; }
static inline size_t StoreLookahead(const Hasher* h) { return h->type == 6 ? 8 : 4; }

static inline void HasherStore(Hasher* h, const uint8_t* data, size_t ix) { /* mod.rs:1644-1656, 866-875 */
  uint32_t key = HashBytes(h, data + ix);
  size_t minor = h->num[key] & h->block_mask;
  h->buckets[((size_t)key << h->block_bits) + minor] = (uint32_t)ix;
  h->num[key] = (uint16_t)(h->num[key] + 1);
}
static inline void HasherStoreRange(Hasher* h, const uint8_t* data, size_t a, size_t b) {
  for (size_t i = a; i < b; ++i) HasherStore(h, data, i);
}

static inline size_t FindMatchLengthWithLimit(const uint8_t* s1, const uint8_t* s2, size_t limit) {
  size_t i = 0; /* static_dict.rs:125-132 */
  while (i < limit && s1[i] == s2[i]) ++i;
  return i;
}
static inline size_t FindMatchLengthWithLimitMin4(const uint8_t* s1, const uint8_t* s2, size_t limit) {
  uint32_t a, b; /* static_dict.rs:134-147 */
  memcpy(&a, s1, 4);
  memcpy(&b, s2, 4);
  if (a != b) return 0;
  if (limit <= 4 || s1[4] != s2[4]) return MINZ(limit, 4);
  return FindMatchLengthWithLimit(s1 + 5, s2 + 5, limit - 5) + 5;
}

/* Scores for H5/H6: mod.rs:1871-1889 with literal_byte_score 540 (>>2 = 135); penalty :1151-1154. */
static inline uint64_t ScoreLastDistance(size_t len) { return 135ull * len + 1920 + 15; }
static inline uint64_t Score(size_t len, size_t backward) {
  return 1920ull + 135ull * len - 30ull * Log2FloorNonZero(backward);
}
static inline uint64_t PenaltyLastDistance(size_t i) { return 39ull + ((0x1ca10ull >> (i & 0xe)) & 0xe); }

/* adv_prepare_distance_cache: mod.rs:632-651 */
static void PrepareDistanceCache(int* dc, int num) {
  if (num > 4) {
    int last = dc[0];
    dc[4] = last - 1; dc[5] = last + 1; dc[6] = last - 2; dc[7] = last + 2; dc[8] = last - 3; dc[9] = last + 3;
    if (num > 10) {
      int next = dc[1];
      dc[10] = next - 1; dc[11] = next + 1; dc[12] = next - 2; dc[13] = next + 2; dc[14] = next - 3; dc[15] = next + 3;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Static dictionary: SearchInStaticDictionary / TestStaticDictionaryItem, mod.rs:1890-1988.  The dictionary bytes and
 * kStaticDictionaryHash come from the system's brotli libraries at build time (oracle/gen_dict_data.py).
 * ---------------------------------------------------------------------------------------------- */
#include "dict_data.inc"
#define kCutoffTransformsCount 10u
#define kCutoffTransforms 0x071B520ADA2D3200ull
#define kMaxDistance 0x3fffffcu /* params.dist.max_distance, large_window off (encode.rs:318-325) */
static inline uint32_t Hash14(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return (v * kHashMul32) >> (32 - 14);
}
static int TestStaticDictionaryItem(size_t item, const uint8_t* data, size_t max_length, size_t max_backward, int h9,
                                    SearchResult* out) { /* mod.rs:1895-1938 */
  size_t len = item & 0x1f, dist = item >> 5;
  size_t offset = kDictOffsets[len] + len * dist;
  if (len > max_length) return 0;
  size_t matchlen = FindMatchLengthWithLimit(data, kBrotliDictionaryData + offset, len);
  if (matchlen + kCutoffTransformsCount <= len || matchlen == 0) return 0;
  uint64_t cut = len - matchlen;
  size_t transform_id = (size_t)((cut << 2) + ((kCutoffTransforms >> (cut * 6)) & 0x3f));
  size_t backward = max_backward + dist + 1 + (transform_id << kDictSizeBits[len]);
  if (backward > kMaxDistance) return 0;
  uint64_t score = h9 ? ((1920ull * 4 + 540ull * matchlen - 120ull * Log2FloorNonZero(backward)) >> 2)
                      : Score(matchlen, backward);
  if (score < out->score) return 0;
  out->len = matchlen;
  out->len_x_code = len ^ matchlen;
  out->distance = backward;
  out->score = score;
  return 1;
}
static int SearchInStaticDictionary(Hasher* h, const uint8_t* data, size_t max_length, size_t max_backward,
                                    SearchResult* out) { /* mod.rs:1940-1988, shallow = false */
  int found = 0;
  if (h->dict_num_matches < (h->dict_num_lookups >> 7)) return 0;
  size_t key = (size_t)Hash14(data) << 1;
  for (int i = 0; i < 2; ++i, ++key) {
    size_t item = kStaticDictionaryHash[key];
    h->dict_num_lookups++;
    if (item != 0 && TestStaticDictionaryItem(item, data, max_length, max_backward, h->type == 9, out)) {
      h->dict_num_matches++;
      found = 1;
    }
  }
  return found;
}

/* AdvHasher::FindLongestMatch, mod.rs:1684-1812 (ring mask dropped: the oracle works on a flat buffer). */
static int FindLongestMatchAdv(Hasher* h, const uint8_t* data, const int* dist_cache, size_t cur_ix,
                               size_t max_length, size_t max_backward, SearchResult* out) {
  int found = 0;
  uint64_t best_score = out->score;
  size_t best_len = out->len;
  const uint8_t* cur = data + cur_ix;
  out->len = 0;
  out->len_x_code = 0;
  for (int i = 0; i < h->n_last; ++i) {
    size_t backward = (size_t)dist_cache[i];
    size_t prev_ix = cur_ix - backward;
    if (prev_ix >= cur_ix || backward > max_backward) continue;
    if (cur[best_len] != data[prev_ix + best_len]) continue;
    size_t len = FindMatchLengthWithLimit(data + prev_ix, cur, max_length);
    if (len >= 3 || (len == 2 && i < 2)) {
      uint64_t score = ScoreLastDistance(len);
      if (best_score < score) {
        if (i != 0) score -= PenaltyLastDistance((size_t)i);
        if (best_score < score) {
          best_score = score; best_len = len;
          out->len = len; out->distance = backward; out->score = score;
          found = 1;
        }
      }
    }
  }
  uint32_t key = HashBytes(h, cur);
  uint32_t* bucket = h->buckets + ((size_t)key << h->block_bits);
  uint16_t num_copy = h->num[key];
  if (num_copy != 0) {
    int block_size = 1 << h->block_bits;
    size_t down = (size_t)MAXZ((int)num_copy - block_size, 0);
    for (size_t i = num_copy; i > down;) {
      --i;
      size_t prev_ix = bucket[i & h->block_mask];
      size_t backward = cur_ix - prev_ix;
      if (cur[best_len] != data[prev_ix + best_len]) {
        if (backward > max_backward) break;
        continue;
      }
      if (backward > max_backward) break;
      size_t len = FindMatchLengthWithLimitMin4(data + prev_ix, cur, max_length);
      if (len != 0) {
        uint64_t score = Score(len, backward);
        if (best_score < score) {
          best_score = score; best_len = len;
          out->len = len; out->distance = backward; out->score = score;
          found = 1;
        }
      }
    }
  }
  bucket[num_copy & h->block_mask] = (uint32_t)cur_ix;
  h->num[key] = (uint16_t)(num_copy + 1);
  if (!found && h->use_dictionary) found = SearchInStaticDictionary(h, cur, max_length, max_backward, out); /* :1797-1810 */
  return found;
}

/* H9: mod.rs:598-917.  bucket_bits 15, block 256, 16 last distances; scores (mod.rs:685-708):
 *   score = (30*8*8*4 + 540*len - 120*log2(dist)) >> 2 ; last distance: (540*len + kDistanceShortCodeCost[i]) >> 2 */
static const uint32_t kDistanceShortCodeCost[16] = {
    /* mod.rs:657-683: (BROTLI_SCORE_BASE + 60 / + ...) table as published in the reference */
    1920 * 4 + 60, 1920 * 4 - 95, 1920 * 4 - 117, 1920 * 4 - 127, 1920 * 4 - 93, 1920 * 4 - 93, 1920 * 4 - 96,
    1920 * 4 - 96,  1920 * 4 - 99, 1920 * 4 - 99,  1920 * 4 - 105, 1920 * 4 - 105, 1920 * 4 - 115, 1920 * 4 - 115,
    1920 * 4 - 125, 1920 * 4 - 125};
static const uint8_t kDistanceCacheIndex[16] = {0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1};
static const int8_t kDistanceCacheOffset[16] = {0, 0, 0, 0, -1, 1, -2, 2, -3, 3, -1, 1, -2, 2, -3, 3};

static int FindLongestMatchH9(Hasher* h, const uint8_t* data, const int* dist_cache, size_t cur_ix,
                              size_t max_length, size_t max_backward, SearchResult* out) {
  int found = 0;
  size_t best_len = out->len;
  uint64_t best_score = out->score;
  const uint8_t* cur = data + cur_ix;
  out->len_x_code = 0;
  for (int i = 0; i < 16; ++i) { /* mod.rs:751-800 */
    int idx = kDistanceCacheIndex[i];
    size_t backward = (size_t)(dist_cache[idx] + kDistanceCacheOffset[i]);
    size_t prev_ix = cur_ix - backward;
    if (prev_ix >= cur_ix) continue;
    if (backward > max_backward) continue;
    if (cur[best_len] != data[prev_ix + best_len]) continue;
    size_t len = FindMatchLengthWithLimit(data + prev_ix, cur, max_length);
    if (len >= 3 || (len == 2 && i < 2)) {
      uint64_t score = (540ull * len + kDistanceShortCodeCost[i]) >> 2;
      if (best_score < score) {
        best_score = score; best_len = len;
        out->len = len; out->distance = backward; out->score = score;
        found = 1;
      }
    }
  }
  if (max_length >= 4) { /* mod.rs:801-858 */
    uint32_t key = HashBytes(h, cur);
    uint32_t* bucket = h->buckets + ((size_t)key << h->block_bits);
    uint16_t num_copy = h->num[key];
    int block_size = 1 << h->block_bits;
    size_t down = (num_copy > block_size) ? (size_t)(num_copy - block_size) : 0;
    for (size_t i = num_copy; i > down;) {
      --i;
      size_t prev_ix = bucket[i & h->block_mask];
      size_t backward = cur_ix - prev_ix;
      if (backward > max_backward) break;
      if (cur[best_len] != data[prev_ix + best_len]) continue;
      size_t len = FindMatchLengthWithLimit(data + prev_ix, cur, max_length);
      if (len >= 4) {
        uint64_t score = (1920ull * 4 + 540ull * len - 120ull * Log2FloorNonZero(backward)) >> 2;
        if (best_score < score) {
          best_score = score; best_len = len;
          out->len = len; out->distance = backward; out->score = score;
          found = 1;
        }
      }
    }
    bucket[num_copy & h->block_mask] = (uint32_t)cur_ix;
    h->num[key] = (uint16_t)(num_copy + 1);
  }
  if (!found && h->use_dictionary) found = SearchInStaticDictionary(h, cur, max_length, max_backward, out); /* :862-875 */
  return found;
}

static inline int FindLongestMatch(Hasher* h, const uint8_t* data, const int* dc, size_t cur_ix, size_t max_length,
                                   size_t max_backward, SearchResult* out) {
  return h->type == 9 ? FindLongestMatchH9(h, data, dc, cur_ix, max_length, max_backward, out)
                      : FindLongestMatchAdv(h, data, dc, cur_ix, max_length, max_backward, out);
}

/* ------------------------------------------------------------------------------------------------
 * Greedy + lazy parse: CreateBackwardReferences, backward_references/mod.rs:2376-2552.
 * ---------------------------------------------------------------------------------------------- */
static void CreateBackwardReferences(size_t num_bytes, size_t position, const uint8_t* data, int quality, int lgwin,
                                     Hasher* hasher, int* dist_cache, size_t* last_insert_len, Command* commands,
                                     size_t* num_commands, size_t* num_literals) {
  const size_t max_backward_limit = ((size_t)1 << lgwin) - 16;
  size_t insert_length = *last_insert_len;
  const size_t pos_end = position + num_bytes;
  const size_t store_end = num_bytes >= StoreLookahead(hasher) ? position + num_bytes - StoreLookahead(hasher) + 1 : position;
  const size_t window = quality < 9 ? 64 : 512; /* LiteralSpreeLengthForSparseSearch mod.rs:150-152 */
  size_t apply_random_heuristics = position + window;
  const uint64_t kMinScore = 30 * 8 * 8 + 100;
  Command* out = commands + *num_commands;
  size_t new_commands = 0;
  PrepareDistanceCache(dist_cache, hasher->n_last);
  while (position + HashTypeLength(hasher) < pos_end) {
    size_t max_length = pos_end - position;
    size_t max_distance = MINZ(position, max_backward_limit);
    SearchResult sr = {0, 0, 0, kMinScore};
    if (FindLongestMatch(hasher, data, dist_cache, position, max_length, max_distance, &sr)) {
      int delayed = 0;
      max_length--;
      for (;; max_length--) {
        SearchResult sr2 = {0, 0, 0, kMinScore};
        sr2.len = quality < 5 ? MINZ(sr.len - 1, max_length) : 0;
        max_distance = MINZ(position + 1, max_backward_limit);
        int found = FindLongestMatch(hasher, data, dist_cache, position + 1, max_length, max_distance, &sr2);
        if (found && sr2.score >= sr.score + 175) {
          position++;
          insert_length++;
          sr = sr2;
          if (++delayed < 4 && position + HashTypeLength(hasher) < pos_end) continue;
        }
        break;
      }
      apply_random_heuristics = position + 2 * sr.len + window;
      max_distance = MINZ(position, max_backward_limit);
      {
        size_t distance_code = oracle_compute_distance_code(sr.distance, max_distance, dist_cache);
        if (sr.distance <= max_distance && distance_code > 0) {
          dist_cache[3] = dist_cache[2];
          dist_cache[2] = dist_cache[1];
          dist_cache[1] = dist_cache[0];
          dist_cache[0] = (int)sr.distance;
          PrepareDistanceCache(dist_cache, hasher->n_last);
        }
        CommandInit(&out[new_commands++], insert_length, sr.len, sr.len ^ sr.len_x_code, distance_code);
      }
      *num_literals += insert_length;
      insert_length = 0;
      HasherStoreRange(hasher, data, position + 2, MINZ(position + sr.len, store_end));
      position += sr.len;
    } else {
      insert_length++;
      position++;
      if (position > apply_random_heuristics) {
        size_t kMargin = MAXZ(StoreLookahead(hasher) - 1, 4);
        if (position + 16 >= pos_end - kMargin) {
          insert_length += pos_end - position;
          position = pos_end;
        } else if (position > apply_random_heuristics + 4 * window) {
          for (int i = 0; i < 4; ++i) HasherStore(hasher, data, position + (size_t)i * 4); /* Store4Vec4 */
          insert_length += 16;
          position += 16;
        } else {
          for (int i = 0; i < 4; ++i) HasherStore(hasher, data, position + (size_t)i * 2); /* StoreEvenVec4 */
          insert_length += 8;
          position += 8;
        }
      }
    }
  }
  insert_length += pos_end - position;
  *last_insert_len = insert_length;
  *num_commands += new_commands;
}

/* ------------------------------------------------------------------------------------------------
 * Entropy: bit_cost.rs:13-42 (note the u16 truncation of counts at :21,:26 -- restated as is).
 * ---------------------------------------------------------------------------------------------- */
static float ShannonEntropy(const uint32_t* population, size_t size, size_t* total) {
  size_t sum = 0;
  float retval = 0.0f;
  for (size_t i = 0; i < size; ++i) {
    size_t p = population[i];
    sum += p;
    retval -= (float)p * FastLog2u16((uint16_t)p);
  }
  if (sum) retval += (float)sum * FastLog2(sum);
  *total = sum;
  return retval;
}
float oracle_bits_entropy(const uint32_t* population, size_t size) {
  init_tables();
  size_t sum;
  float r = ShannonEntropy(population, size, &sum);
  if (r < (float)sum) r = (float)sum;
  return r;
}
#define BitsEntropy oracle_bits_entropy

/* BrotliPopulationCost: bit_cost.rs:76-211 (used at q>=10 and by tests of the GPU cost kernel). */
float oracle_population_cost(const uint32_t* data, size_t data_size) {
  init_tables();
  static const float kOneSymbolHistogramCost = 12, kTwoSymbolHistogramCost = 20, kThreeSymbolHistogramCost = 28,
                     kFourSymbolHistogramCost = 37;
  size_t total_count = 0;
  for (size_t i = 0; i < data_size; ++i) total_count += data[i];
  if (total_count == 0) return kOneSymbolHistogramCost;
  int count = 0;
  size_t s[5];
  for (size_t i = 0; i < data_size; ++i) {
    if (data[i] > 0) {
      s[count] = i;
      if (++count > 4) break;
    }
  }
  if (count == 1) return kOneSymbolHistogramCost;
  if (count == 2) return kTwoSymbolHistogramCost + (float)total_count;
  if (count == 3) {
    uint32_t h0 = data[s[0]], h1 = data[s[1]], h2 = data[s[2]];
    uint32_t histomax = MAXZ(h0, MAXZ(h1, h2));
    return kThreeSymbolHistogramCost + 2.0f * (float)(h0 + h1 + h2) - (float)histomax;
  }
  if (count == 4) {
    uint32_t histo[4];
    for (int i = 0; i < 4; ++i) histo[i] = data[s[i]];
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j)
        if (histo[j] > histo[i]) { uint32_t t = histo[j]; histo[j] = histo[i]; histo[i] = t; }
    uint32_t h23 = histo[2] + histo[3];
    uint32_t histomax = MAXZ(h23, histo[0]);
    return kFourSymbolHistogramCost + 3.0f * (float)h23 + 2.0f * (float)(histo[0] + histo[1]) - (float)histomax;
  }
  {
    float bits = 0.0f;
    size_t max_depth = 1;
    uint32_t depth_histo[18];
    memset(depth_histo, 0, sizeof(depth_histo));
    const float log2total = FastLog2(total_count);
    for (size_t i = 0; i < data_size;) {
      if (data[i] > 0) {
        float log2p = log2total - FastLog2u16((uint16_t)data[i]);
        size_t depth = (size_t)(log2p + 0.5f);
        bits += (float)data[i] * log2p;
        if (depth > 15) depth = 15;
        if (depth > max_depth) max_depth = depth;
        depth_histo[depth]++;
        ++i;
      } else {
        uint32_t reps = 1;
        for (size_t k = i + 1; k < data_size && data[k] == 0; ++k) ++reps;
        i += reps;
        if (i == data_size) break;
        if (reps < 3) depth_histo[0] += reps;
        else {
          reps -= 2;
          while (reps > 0) {
            depth_histo[17]++;
            bits += 3;
            reps >>= 3;
          }
        }
      }
    }
    bits += (float)(18 + 2 * max_depth);
    bits += BitsEntropy(depth_histo, 18);
    return bits;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Huffman: entropy_encode.rs:27-56 (SetDepth), :71-116 (sort), :133-210 (CreateHuffmanTree),
 * :211-345 (OptimizeHuffmanCountsForRle), :347-525 (WriteHuffmanTree), :546-575 (depths -> codes).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t total_count_;
  int16_t index_left_;
  int16_t index_right_or_value_;
} HuffmanTree;

static int SetDepth(int p0, HuffmanTree* pool, uint8_t* depth, int max_depth) {
  int stack[16];
  int level = 0;
  int p = p0;
  stack[0] = -1;
  for (;;) {
    if (pool[p].index_left_ >= 0) {
      level++;
      if (level > max_depth) return 0;
      stack[level] = pool[p].index_right_or_value_;
      p = pool[p].index_left_;
      continue;
    } else {
      depth[pool[p].index_right_or_value_] = (uint8_t)level;
    }
    while (level >= 0 && stack[level] == -1) level--;
    if (level < 0) return 1;
    p = stack[level];
    stack[level] = -1;
  }
}
static inline int SortCmp(const HuffmanTree* a, const HuffmanTree* b) {
  if (a->total_count_ != b->total_count_) return a->total_count_ < b->total_count_;
  return a->index_right_or_value_ > b->index_right_or_value_;
}
static void SortHuffmanTreeItems(HuffmanTree* items, size_t n) {
  static const size_t gaps[6] = {132, 57, 23, 10, 4, 1};
  if (n < 13) {
    for (size_t i = 1; i < n; ++i) {
      HuffmanTree tmp = items[i];
      size_t k = i, j = i - 1;
      while (SortCmp(&tmp, &items[j])) {
        items[k] = items[j];
        k = j;
        if (!j--) break;
      }
      items[k] = tmp;
    }
  } else {
    for (int g = n < 57 ? 2 : 0; g < 6; ++g) {
      size_t gap = gaps[g];
      for (size_t i = gap; i < n; ++i) {
        size_t j = i;
        HuffmanTree tmp = items[i];
        for (; j >= gap && SortCmp(&tmp, &items[j - gap]); j -= gap) items[j] = items[j - gap];
        items[j] = tmp;
      }
    }
  }
}
void oracle_create_huffman_tree(const uint32_t* data, size_t length, int tree_limit, uint8_t* depth) {
  HuffmanTree* tree = (HuffmanTree*)malloc(sizeof(HuffmanTree) * (2 * length + 2));
  const HuffmanTree sentinel = {0xFFFFFFFFu, -1, -1};
  for (uint32_t count_limit = 1;; count_limit *= 2) {
    size_t n = 0;
    for (size_t i = length; i != 0;) {
      --i;
      if (data[i]) {
        HuffmanTree t = {MAXZ(data[i], count_limit), -1, (int16_t)i};
        tree[n++] = t;
      }
    }
    if (n == 1) {
      depth[tree[0].index_right_or_value_] = 1;
      break;
    }
    SortHuffmanTreeItems(tree, n);
    tree[n] = sentinel;
    tree[n + 1] = sentinel;
    size_t i = 0, j = n + 1;
    for (size_t k = n - 1; k != 0; --k) {
      size_t left, right;
      if (tree[i].total_count_ <= tree[j].total_count_) left = i++; else left = j++;
      if (tree[i].total_count_ <= tree[j].total_count_) right = i++; else right = j++;
      size_t j_end = 2 * n - k;
      tree[j_end].total_count_ = tree[left].total_count_ + tree[right].total_count_;
      tree[j_end].index_left_ = (int16_t)left;
      tree[j_end].index_right_or_value_ = (int16_t)right;
      tree[j_end + 1] = sentinel;
    }
    if (SetDepth((int)(2 * n - 1), tree, depth, tree_limit)) break;
  }
  free(tree);
}
#define CreateHuffmanTree oracle_create_huffman_tree

void oracle_optimize_huffman_counts_for_rle(size_t length, uint32_t* counts) {
  uint8_t good_for_rle[704];
  size_t nonzero_count = 0, stride, limit, sum;
  const size_t streak_limit = 1240;
  for (size_t i = 0; i < length; ++i) if (counts[i]) ++nonzero_count;
  if (nonzero_count < 16) return;
  while (length != 0 && counts[length - 1] == 0) --length;
  if (length == 0) return;
  {
    size_t nonzeros = 0;
    uint32_t smallest_nonzero = 1u << 30;
    for (size_t i = 0; i < length; ++i) {
      if (counts[i] != 0) {
        ++nonzeros;
        if (smallest_nonzero > counts[i]) smallest_nonzero = counts[i];
      }
    }
    if (nonzeros < 5) return;
    if (smallest_nonzero < 4) {
      size_t zeros = length - nonzeros;
      if (zeros < 6)
        for (size_t i = 1; i < length - 1; ++i)
          if (counts[i - 1] != 0 && counts[i] == 0 && counts[i + 1] != 0) counts[i] = 1;
    }
    if (nonzeros < 28) return;
  }
  memset(good_for_rle, 0, sizeof(good_for_rle));
  {
    uint32_t symbol = counts[0];
    size_t step = 0;
    for (size_t i = 0; i <= length; ++i) {
      if (i == length || counts[i] != symbol) {
        if ((symbol == 0 && step >= 5) || (symbol != 0 && step >= 7))
          for (size_t k = 0; k < step; ++k) good_for_rle[i - k - 1] = 1;
        step = 1;
        if (i != length) symbol = counts[i];
      } else {
        ++step;
      }
    }
  }
  stride = 0;
  limit = 256 * (counts[0] + counts[1] + counts[2]) / 3 + 420;
  sum = 0;
  for (size_t i = 0; i <= length; ++i) {
    if (i == length || good_for_rle[i] || (i != 0 && good_for_rle[i - 1]) ||
        (256 * (size_t)counts[i] - limit + streak_limit) >= 2 * streak_limit) {
      if (stride >= 4 || (stride >= 3 && sum == 0)) {
        size_t count = (sum + stride / 2) / stride;
        if (count == 0) count = 1;
        if (sum == 0) count = 0;
        for (size_t k = 0; k < stride; ++k) counts[i - k - 1] = (uint32_t)count;
      }
      stride = 0;
      sum = 0;
      if (i < length - 2) limit = 256 * (counts[i] + counts[i + 1] + counts[i + 2]) / 3 + 420;
      else if (i < length) limit = 256 * (size_t)counts[i];
      else limit = 0;
    }
    ++stride;
    if (i != length) {
      sum += counts[i];
      if (stride >= 4) limit = (256 * sum + stride / 2) / stride;
      if (stride == 4) limit += 120;
    }
  }
}

static void DecideOverRleUse(const uint8_t* depth, size_t length, int* use_nz, int* use_z) {
  size_t total_reps_zero = 0, total_reps_non_zero = 0, count_reps_zero = 1, count_reps_non_zero = 1;
  for (size_t i = 0; i < length;) {
    uint8_t value = depth[i];
    size_t reps = 1;
    for (size_t k = i + 1; k < length && depth[k] == value; ++k) ++reps;
    if (reps >= 3 && value == 0) { total_reps_zero += reps; ++count_reps_zero; }
    if (reps >= 4 && value != 0) { total_reps_non_zero += reps; ++count_reps_non_zero; }
    i += reps;
  }
  *use_nz = total_reps_non_zero > count_reps_non_zero * 2;
  *use_z = total_reps_zero > count_reps_zero * 2;
}
static void ReverseU8(uint8_t* v, size_t start, size_t end) {
  --end;
  while (start < end) { uint8_t t = v[start]; v[start] = v[end]; v[end] = t; ++start; --end; }
}
static void WriteRepetitions(uint8_t previous_value, uint8_t value, size_t repetitions, size_t* tree_size,
                             uint8_t* tree, uint8_t* extra) {
  if (previous_value != value) { tree[*tree_size] = value; extra[*tree_size] = 0; ++*tree_size; --repetitions; }
  if (repetitions == 7) { tree[*tree_size] = value; extra[*tree_size] = 0; ++*tree_size; --repetitions; }
  if (repetitions < 3) {
    for (size_t i = 0; i < repetitions; ++i) { tree[*tree_size] = value; extra[*tree_size] = 0; ++*tree_size; }
  } else {
    size_t start = *tree_size;
    repetitions -= 3;
    for (;;) {
      tree[*tree_size] = 16; extra[*tree_size] = repetitions & 3; ++*tree_size;
      repetitions >>= 2;
      if (repetitions == 0) break;
      --repetitions;
    }
    ReverseU8(tree, start, *tree_size);
    ReverseU8(extra, start, *tree_size);
  }
}
static void WriteRepetitionsZeros(size_t repetitions, size_t* tree_size, uint8_t* tree, uint8_t* extra) {
  if (repetitions == 11) { tree[*tree_size] = 0; extra[*tree_size] = 0; ++*tree_size; --repetitions; }
  if (repetitions < 3) {
    for (size_t i = 0; i < repetitions; ++i) { tree[*tree_size] = 0; extra[*tree_size] = 0; ++*tree_size; }
  } else {
    size_t start = *tree_size;
    repetitions -= 3;
    for (;;) {
      tree[*tree_size] = 17; extra[*tree_size] = repetitions & 7; ++*tree_size;
      repetitions >>= 3;
      if (repetitions == 0) break;
      --repetitions;
    }
    ReverseU8(tree, start, *tree_size);
    ReverseU8(extra, start, *tree_size);
  }
}
static void WriteHuffmanTree(const uint8_t* depth, size_t length, size_t* tree_size, uint8_t* tree, uint8_t* extra) {
  uint8_t previous_value = 8;
  int use_nz = 0, use_z = 0;
  size_t new_length = length;
  for (size_t i = 0; i < length; ++i) {
    if (depth[length - i - 1] == 0) --new_length; else break;
  }
  if (length > 50) DecideOverRleUse(depth, new_length, &use_nz, &use_z);
  for (size_t i = 0; i < new_length;) {
    uint8_t value = depth[i];
    size_t reps = 1;
    if ((value != 0 && use_nz) || (value == 0 && use_z))
      for (size_t k = i + 1; k < new_length && depth[k] == value; ++k) ++reps;
    if (value == 0) WriteRepetitionsZeros(reps, tree_size, tree, extra);
    else { WriteRepetitions(previous_value, value, reps, tree_size, tree, extra); previous_value = value; }
    i += reps;
  }
}
static uint16_t ReverseBits(size_t num_bits, uint16_t bits) {
  uint16_t r = 0;
  for (size_t i = 0; i < num_bits; ++i) { r = (uint16_t)((r << 1) | (bits & 1)); bits >>= 1; }
  return r;
}
void oracle_convert_bit_depths_to_symbols(const uint8_t* depth, size_t len, uint16_t* bits) {
  uint16_t bl_count[16] = {0}, next_code[16];
  int code = 0;
  for (size_t i = 0; i < len; ++i) bl_count[depth[i]]++;
  bl_count[0] = 0;
  next_code[0] = 0;
  for (int i = 1; i < 16; ++i) { code = (code + bl_count[i - 1]) << 1; next_code[i] = (uint16_t)code; }
  for (size_t i = 0; i < len; ++i)
    if (depth[i]) bits[i] = ReverseBits(depth[i], next_code[depth[i]]++);
}
#define ConvertBitDepthsToSymbols oracle_convert_bit_depths_to_symbols

/* ------------------------------------------------------------------------------------------------
 * Huffman tree serialisation: brotli_bit_stream.rs:764-911, :1401-1498.
 * ---------------------------------------------------------------------------------------------- */
static void StoreHuffmanTreeOfHuffmanTreeToBitMask(int num_codes, const uint8_t* cl_depth, size_t* ix, uint8_t* st) {
  static const uint8_t kStorageOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  static const uint8_t kSymbols[6] = {0, 7, 3, 2, 1, 15};
  static const uint8_t kLengths[6] = {2, 4, 3, 2, 2, 4};
  size_t skip_some = 0, codes_to_store = 18;
  if (num_codes > 1)
    for (; codes_to_store > 0; --codes_to_store)
      if (cl_depth[kStorageOrder[codes_to_store - 1]] != 0) break;
  if (cl_depth[kStorageOrder[0]] == 0 && cl_depth[kStorageOrder[1]] == 0) {
    skip_some = 2;
    if (cl_depth[kStorageOrder[2]] == 0) skip_some = 3;
  }
  WriteBits(2, skip_some, ix, st);
  for (size_t i = skip_some; i < codes_to_store; ++i) {
    size_t l = cl_depth[kStorageOrder[i]];
    WriteBits(kLengths[l], kSymbols[l], ix, st);
  }
}
static void StoreHuffmanTree(const uint8_t* depths, size_t num, size_t* ix, uint8_t* st) {
  uint8_t huffman_tree[704], extra_bits[704];
  size_t huffman_tree_size = 0;
  uint8_t cl_depth[18] = {0};
  uint16_t cl_bits[18] = {0};
  uint32_t histogram[18] = {0};
  int num_codes = 0;
  size_t code = 0;
  WriteHuffmanTree(depths, num, &huffman_tree_size, huffman_tree, extra_bits);
  for (size_t i = 0; i < huffman_tree_size; ++i) ++histogram[huffman_tree[i]];
  for (size_t i = 0; i < 18; ++i) {
    if (histogram[i]) {
      if (num_codes == 0) { code = i; num_codes = 1; }
      else if (num_codes == 1) { num_codes = 2; break; }
    }
  }
  CreateHuffmanTree(histogram, 18, 5, cl_depth);
  ConvertBitDepthsToSymbols(cl_depth, 18, cl_bits);
  StoreHuffmanTreeOfHuffmanTreeToBitMask(num_codes, cl_depth, ix, st);
  if (num_codes == 1) cl_depth[code] = 0;
  for (size_t i = 0; i < huffman_tree_size; ++i) {
    size_t s = huffman_tree[i];
    WriteBits(cl_depth[s], cl_bits[s], ix, st);
    if (s == 16) WriteBits(2, extra_bits[i], ix, st);
    else if (s == 17) WriteBits(3, extra_bits[i], ix, st);
  }
}
static void StoreSimpleHuffmanTree(const uint8_t* depths, size_t* symbols, size_t num_symbols, size_t max_bits,
                                   size_t* ix, uint8_t* st) {
  WriteBits(2, 1, ix, st);
  WriteBits(2, num_symbols - 1, ix, st);
  for (size_t i = 0; i < num_symbols; ++i)
    for (size_t j = i + 1; j < num_symbols; ++j)
      if (depths[symbols[j]] < depths[symbols[i]]) { size_t t = symbols[j]; symbols[j] = symbols[i]; symbols[i] = t; }
  for (size_t i = 0; i < num_symbols; ++i) WriteBits((unsigned)max_bits, symbols[i], ix, st);
  if (num_symbols == 4) WriteBits(1, depths[symbols[0]] == 1 ? 1 : 0, ix, st);
}
static void BuildAndStoreHuffmanTree(const uint32_t* histogram, size_t histogram_length, size_t alphabet_size,
                                     uint8_t* depth, uint16_t* bits, size_t* ix, uint8_t* st) {
  size_t count = 0, s4[4] = {0}, max_bits = 0;
  for (size_t i = 0; i < histogram_length; ++i) {
    if (histogram[i]) {
      if (count < 4) s4[count] = i;
      else if (count > 4) break;
      count++;
    }
  }
  for (size_t c = alphabet_size - 1; c; c >>= 1) ++max_bits;
  if (count <= 1) {
    WriteBits(4, 1, ix, st);
    WriteBits((unsigned)max_bits, s4[0], ix, st);
    depth[s4[0]] = 0;
    bits[s4[0]] = 0;
    return;
  }
  memset(depth, 0, histogram_length);
  CreateHuffmanTree(histogram, histogram_length, 15, depth);
  ConvertBitDepthsToSymbols(depth, histogram_length, bits);
  if (count <= 4) StoreSimpleHuffmanTree(depth, s4, count, max_bits, ix, st);
  else StoreHuffmanTree(depth, histogram_length, ix, st);
}
/* test hook: serialise one prefix code exactly as the reference does; returns number of bits written */
size_t oracle_build_and_store_huffman_tree(const uint32_t* histogram, size_t length, size_t alphabet_size,
                                           uint8_t* depth, uint16_t* bits, uint8_t* storage) {
  size_t ix = 0;
  storage[0] = 0;
  BuildAndStoreHuffmanTree(histogram, length, alphabet_size, depth, bits, &ix, storage);
  return ix;
}

/* ------------------------------------------------------------------------------------------------
 * Block splits and the greedy (q4..q9) metablock builder: metablock.rs:385-1021.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  size_t num_types, num_blocks;
  uint8_t* types;
  uint32_t* lengths;
} BlockSplit;

typedef struct {
  size_t alphabet_size, num_contexts, max_block_types, min_block_size;
  float split_threshold;
  size_t num_blocks, target_block_size, block_size, curr_histogram_ix, merge_last_count;
  size_t last_histogram_ix[2];
  float last_entropy[2 * 13];
  BlockSplit* split;
  uint32_t* histograms; /* [histograms_size][alphabet_size] */
  size_t histograms_size;
} Splitter;

static void SplitterInit(Splitter* s, size_t alphabet_size, size_t num_contexts, size_t min_block_size,
                         float split_threshold, size_t num_symbols, BlockSplit* split) {
  size_t max_num_blocks = num_symbols / min_block_size + 1;
  memset(s, 0, sizeof(*s));
  s->alphabet_size = alphabet_size;
  s->num_contexts = num_contexts;
  s->max_block_types = 256 / num_contexts; /* metablock.rs:486; plain splitter uses 256 (:595) */
  s->min_block_size = min_block_size;
  s->split_threshold = split_threshold;
  s->target_block_size = min_block_size;
  s->split = split;
  size_t max_num_types = MINZ(max_num_blocks, s->max_block_types + 1);
  split->types = (uint8_t*)calloc(max_num_blocks, 1);
  split->lengths = (uint32_t*)calloc(max_num_blocks, 4);
  split->num_blocks = max_num_blocks;
  split->num_types = 0;
  s->histograms_size = max_num_types * num_contexts;
  s->histograms = (uint32_t*)calloc(s->histograms_size * alphabet_size, 4);
}
static inline uint32_t* SH(Splitter* s, size_t ix) { return s->histograms + ix * s->alphabet_size; }

/* Covers BlockSplitterFinishBlock (metablock.rs:551-656) and ContextBlockSplitterFinishBlock (:659-792): the
 * plain splitter is the num_contexts == 1 case of the context splitter. */
static void SplitterFinishBlock(Splitter* s, int is_final) {
  const size_t nc = s->num_contexts, A = s->alphabet_size;
  BlockSplit* split = s->split;
  if (s->block_size < s->min_block_size) s->block_size = s->min_block_size;
  if (s->num_blocks == 0) {
    split->lengths[0] = (uint32_t)s->block_size;
    split->types[0] = 0;
    for (size_t i = 0; i < nc; ++i) {
      s->last_entropy[i] = BitsEntropy(SH(s, i), A);
      s->last_entropy[nc + i] = s->last_entropy[i];
    }
    ++s->num_blocks;
    ++split->num_types;
    s->curr_histogram_ix += nc;
    if (s->curr_histogram_ix < s->histograms_size) memset(SH(s, s->curr_histogram_ix), 0, nc * A * 4);
    s->block_size = 0;
  } else if (s->block_size > 0) {
    float entropy[13], combined_entropy[26], diff[2] = {0.0f, 0.0f};
    uint32_t* combined = (uint32_t*)malloc(2 * nc * A * 4);
    for (size_t i = 0; i < nc; ++i) {
      size_t curr = s->curr_histogram_ix + i;
      entropy[i] = BitsEntropy(SH(s, curr), A);
      for (size_t j = 0; j < 2; ++j) {
        size_t jx = j * nc + i;
        size_t last = s->last_histogram_ix[j] + i;
        uint32_t* c = combined + jx * A;
        for (size_t k = 0; k < A; ++k) c[k] = SH(s, curr)[k] + SH(s, last)[k];
        combined_entropy[jx] = BitsEntropy(c, A);
        diff[j] += combined_entropy[jx] - entropy[i] - s->last_entropy[jx];
      }
    }
    if (split->num_types < s->max_block_types && diff[0] > s->split_threshold && diff[1] > s->split_threshold) {
      split->lengths[s->num_blocks] = (uint32_t)s->block_size;
      split->types[s->num_blocks] = (uint8_t)split->num_types;
      s->last_histogram_ix[1] = s->last_histogram_ix[0];
      s->last_histogram_ix[0] = split->num_types * nc;
      for (size_t i = 0; i < nc; ++i) {
        s->last_entropy[nc + i] = s->last_entropy[i];
        s->last_entropy[i] = entropy[i];
      }
      ++s->num_blocks;
      ++split->num_types;
      s->curr_histogram_ix += nc;
      if (s->curr_histogram_ix < s->histograms_size) memset(SH(s, s->curr_histogram_ix), 0, nc * A * 4);
      s->block_size = 0;
      s->merge_last_count = 0;
      s->target_block_size = s->min_block_size;
    } else if (diff[1] < diff[0] - 20.0f) {
      split->lengths[s->num_blocks] = (uint32_t)s->block_size;
      split->types[s->num_blocks] = split->types[s->num_blocks - 2];
      { size_t t = s->last_histogram_ix[0]; s->last_histogram_ix[0] = s->last_histogram_ix[1]; s->last_histogram_ix[1] = t; }
      for (size_t i = 0; i < nc; ++i) {
        memcpy(SH(s, s->last_histogram_ix[0] + i), combined + (nc + i) * A, A * 4);
        s->last_entropy[nc + i] = s->last_entropy[i];
        s->last_entropy[i] = combined_entropy[nc + i];
        memset(SH(s, s->curr_histogram_ix + i), 0, A * 4);
      }
      ++s->num_blocks;
      s->block_size = 0;
      s->merge_last_count = 0;
      s->target_block_size = s->min_block_size;
    } else {
      split->lengths[s->num_blocks - 1] += (uint32_t)s->block_size;
      for (size_t i = 0; i < nc; ++i) {
        memcpy(SH(s, s->last_histogram_ix[0] + i), combined + i * A, A * 4);
        s->last_entropy[i] = combined_entropy[i];
        if (split->num_types == 1) s->last_entropy[nc + i] = s->last_entropy[i];
        memset(SH(s, s->curr_histogram_ix + i), 0, A * 4);
      }
      s->block_size = 0;
      if (++s->merge_last_count > 1) s->target_block_size += s->min_block_size;
    }
    free(combined);
  }
  if (is_final) {
    s->histograms_size = split->num_types * nc;
    split->num_blocks = s->num_blocks;
  }
}
static inline void SplitterAddSymbol(Splitter* s, size_t symbol, size_t context) {
  SH(s, s->curr_histogram_ix + context)[symbol]++;
  if (++s->block_size == s->target_block_size) SplitterFinishBlock(s, 0);
}

typedef struct {
  BlockSplit literal_split, command_split, distance_split;
  uint32_t* literal_context_map; /* num_types << 6, or NULL */
  size_t literal_context_map_size;
  uint32_t *literal_histograms, *command_histograms, *distance_histograms;
  size_t literal_histograms_size, command_histograms_size, distance_histograms_size;
} MetaBlockSplit;

static void BuildMetaBlockGreedy(const uint8_t* data, size_t pos, uint8_t prev_byte, uint8_t prev_byte2,
                                 size_t num_contexts, const uint32_t* static_context_map, const Command* commands,
                                 size_t n_commands, MetaBlockSplit* mb) {
  Splitter lit, cmd, dist;
  size_t num_literals = 0;
  for (size_t i = 0; i < n_commands; ++i) num_literals += commands[i].insert_len_;
  SplitterInit(&lit, 256, num_contexts, 512, 400.0f, num_literals, &mb->literal_split);
  if (num_contexts == 1) lit.max_block_types = 256;
  SplitterInit(&cmd, 704, 1, 1024, 500.0f, n_commands, &mb->command_split);
  cmd.max_block_types = 256;
  SplitterInit(&dist, 64, 1, 512, 100.0f, n_commands, &mb->distance_split);
  dist.max_block_types = 256;
  for (size_t i = 0; i < n_commands; ++i) {
    const Command c = commands[i];
    SplitterAddSymbol(&cmd, c.cmd_prefix_, 0);
    for (size_t j = c.insert_len_; j != 0; --j) {
      uint8_t literal = data[pos];
      size_t ctx = num_contexts == 1 ? 0 : static_context_map[ContextUTF8(prev_byte, prev_byte2)];
      SplitterAddSymbol(&lit, literal, ctx);
      prev_byte2 = prev_byte;
      prev_byte = literal;
      ++pos;
    }
    pos += CommandCopyLen(&c);
    if (CommandCopyLen(&c)) {
      prev_byte2 = data[pos - 2];
      prev_byte = data[pos - 1];
      if (c.cmd_prefix_ >= 128) SplitterAddSymbol(&dist, c.dist_prefix_ & 0x3ff, 0);
    }
  }
  if (getenv("ORACLE_DEBUG")) fprintf(stderr, "before final: lit pending=%zu target=%zu num_literals=%zu\n", lit.block_size, lit.target_block_size, num_literals);
  SplitterFinishBlock(&lit, 1);
  SplitterFinishBlock(&cmd, 1);
  SplitterFinishBlock(&dist, 1);
  mb->literal_histograms = lit.histograms; mb->literal_histograms_size = lit.histograms_size;
  mb->command_histograms = cmd.histograms; mb->command_histograms_size = cmd.histograms_size;
  mb->distance_histograms = dist.histograms; mb->distance_histograms_size = dist.histograms_size;
  mb->literal_context_map = NULL;
  mb->literal_context_map_size = 0;
  if (num_contexts > 1) { /* MapStaticContexts metablock.rs:832-857 */
    mb->literal_context_map_size = mb->literal_split.num_types << 6;
    mb->literal_context_map = (uint32_t*)malloc(mb->literal_context_map_size * 4);
    for (size_t i = 0; i < mb->literal_split.num_types; ++i)
      for (size_t j = 0; j < 64; ++j)
        mb->literal_context_map[(i << 6) + j] = (uint32_t)(i * num_contexts) + static_context_map[j];
  }
}
static void MetaBlockSplitFree(MetaBlockSplit* mb) {
  free(mb->literal_split.types); free(mb->literal_split.lengths);
  free(mb->command_split.types); free(mb->command_split.lengths);
  free(mb->distance_split.types); free(mb->distance_split.lengths);
  free(mb->literal_context_map);
  free(mb->literal_histograms); free(mb->command_histograms); free(mb->distance_histograms);
}

/* ------------------------------------------------------------------------------------------------
 * Literal context decisions: encode.rs:1717-1927.
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t kStaticContextMapContinuation[64] = {1, 1, 2, 2};
static const uint32_t kStaticContextMapSimpleUTF8[64] = {0, 0, 1, 1};
static const uint32_t kStaticContextMapComplexUTF8[64] = {
    11, 11, 12, 12, 0, 0, 0, 0, 1, 1, 9, 9, 2, 2, 2, 2, 1, 1, 1, 1, 8, 3, 3, 3, 1, 1, 1, 1, 2, 2, 2, 2,
    8,  4,  4,  4,  8, 7, 4, 4, 8, 0, 0, 0, 3, 3, 3, 3, 5, 5, 10, 5, 5, 5, 10, 5, 6, 6, 6, 6, 6, 6, 6, 6};

static float ShannonOnly(const uint32_t* p, size_t n) { size_t t; return ShannonEntropy(p, n, &t); }

static void ChooseContextMap(int quality, uint32_t* bigram_histo, size_t* num_literal_contexts, const uint32_t** map) {
  uint32_t monogram_histo[3] = {0}, two_prefix_histo[6] = {0};
  float entropy[4];
  for (size_t i = 0; i < 9; ++i) {
    monogram_histo[i % 3] += bigram_histo[i];
    two_prefix_histo[i % 6] += bigram_histo[i];
  }
  entropy[1] = ShannonOnly(monogram_histo, 3);
  entropy[2] = ShannonOnly(two_prefix_histo, 3) + ShannonOnly(two_prefix_histo + 3, 3);
  entropy[3] = 0.0f;
  for (size_t i = 0; i < 3; ++i) entropy[3] += ShannonOnly(bigram_histo + 3 * i, 3);
  size_t total = monogram_histo[0] + monogram_histo[1] + monogram_histo[2];
  entropy[0] = 1.0f / (float)total;
  entropy[1] *= entropy[0];
  entropy[2] *= entropy[0];
  entropy[3] *= entropy[0];
  if (quality < 7) entropy[3] = entropy[1] * 10.0f;
  if (entropy[1] - entropy[2] < 0.2f && entropy[1] - entropy[3] < 0.2f) {
    *num_literal_contexts = 1;
  } else if (entropy[2] - entropy[3] < 0.02f) {
    *num_literal_contexts = 2;
    *map = kStaticContextMapSimpleUTF8;
  } else {
    *num_literal_contexts = 3;
    *map = kStaticContextMapContinuation;
  }
}
static int ShouldUseComplexStaticContextMap(const uint8_t* input, size_t start_pos, size_t length, size_t size_hint,
                                            size_t* num_literal_contexts, const uint32_t** map) {
  if (size_hint < (1u << 20)) return 0;
  const size_t end_pos = start_pos + length;
  uint32_t combined_histo[32] = {0};
  uint32_t context_histo[13][32];
  uint32_t total = 0;
  float entropy[3];
  memset(context_histo, 0, sizeof(context_histo));
  for (; start_pos + 64 <= end_pos; start_pos += 4096) {
    const size_t stride_end_pos = start_pos + 64;
    uint8_t prev2 = input[start_pos], prev1 = input[start_pos + 1];
    for (size_t pos = start_pos + 2; pos < stride_end_pos; ++pos) {
      const uint8_t literal = input[pos];
      const uint8_t context = (uint8_t)kStaticContextMapComplexUTF8[ContextUTF8(prev1, prev2)];
      ++total;
      ++combined_histo[literal >> 3];
      ++context_histo[context][literal >> 3];
      prev2 = prev1;
      prev1 = literal;
    }
  }
  entropy[1] = ShannonOnly(combined_histo, 32);
  entropy[2] = 0.0f;
  for (size_t i = 0; i < 13; ++i) entropy[2] += ShannonOnly(context_histo[i], 32);
  entropy[0] = 1.0f / (float)total;
  entropy[1] *= entropy[0];
  entropy[2] *= entropy[0];
  if (entropy[2] > 3.0f || entropy[1] - entropy[2] < 0.2f) return 0;
  *num_literal_contexts = 13;
  *map = kStaticContextMapComplexUTF8;
  return 1;
}
static void DecideOverLiteralContextModeling(const uint8_t* input, size_t start_pos, size_t length, int quality,
                                             size_t size_hint, size_t* num_literal_contexts, const uint32_t** map) {
  if (quality < 5 || length < 64) return;
  if (ShouldUseComplexStaticContextMap(input, start_pos, length, size_hint, num_literal_contexts, map)) return;
  const size_t end_pos = start_pos + length;
  uint32_t bigram_prefix_histo[9] = {0};
  static const int lut[4] = {0, 0, 1, 2};
  for (; start_pos + 64 <= end_pos; start_pos += 4096) {
    const size_t stride_end_pos = start_pos + 64;
    int prev = lut[input[start_pos] >> 6] * 3;
    for (size_t pos = start_pos + 1; pos < stride_end_pos; ++pos) {
      const uint8_t literal = input[pos];
      ++bigram_prefix_histo[prev + lut[literal >> 6]];
      prev = lut[literal >> 6] * 3;
    }
  }
  ChooseContextMap(quality, bigram_prefix_histo, num_literal_contexts, map);
}

/* ------------------------------------------------------------------------------------------------
 * Metablock serialisation: brotli_bit_stream.rs:1272-1311, :1357-1399, :1506-1591, :1613-1858, :2035-2261.
 * ---------------------------------------------------------------------------------------------- */
static void StoreCompressedMetaBlockHeader(int is_final, size_t length, size_t* ix, uint8_t* st) {
  WriteBits(1, (uint64_t)is_final, ix, st);
  if (is_final) WriteBits(1, 0, ix, st);
  uint32_t lg = length == 1 ? 1 : Log2FloorNonZero((uint64_t)(length - 1)) + 1;
  uint32_t mnibbles = (lg < 16 ? 16 : lg + 3) / 4;
  WriteBits(2, mnibbles - 4, ix, st);
  WriteBits(mnibbles * 4, length - 1, ix, st);
  if (!is_final) WriteBits(1, 0, ix, st);
}
static void StoreUncompressedMetaBlockHeader(size_t length, size_t* ix, uint8_t* st) {
  WriteBits(1, 0, ix, st); /* brotli_bit_stream.rs:2743-2756 */
  uint32_t lg = length == 1 ? 1 : Log2FloorNonZero((uint64_t)(length - 1)) + 1;
  uint32_t mnibbles = (lg < 16 ? 16 : lg + 3) / 4;
  WriteBits(2, mnibbles - 4, ix, st);
  WriteBits(mnibbles * 4, length - 1, ix, st);
  WriteBits(1, 1, ix, st);
}
static void StoreVarLenUint8(size_t n, size_t* ix, uint8_t* st) {
  if (n == 0) WriteBits(1, 0, ix, st);
  else {
    uint32_t nbits = Log2FloorNonZero(n);
    WriteBits(1, 1, ix, st);
    WriteBits(3, nbits, ix, st);
    WriteBits(nbits, n - ((size_t)1 << nbits), ix, st);
  }
}
static uint32_t BlockLengthPrefixCode(uint32_t len) {
  uint32_t code = (len >= 177) ? (len >= 753 ? 20 : 14) : (len >= 41 ? 7 : 0);
  while (code < 25 && len >= kBlockLenOffset[code + 1]) ++code;
  return code;
}
typedef struct { size_t last_type, second_last_type; } BlockTypeCodeCalculator;
static size_t NextBlockTypeCode(BlockTypeCodeCalculator* c, uint8_t type) {
  size_t type_code = (type == c->last_type + 1) ? 1u : (type == c->second_last_type) ? 0u : (size_t)type + 2u;
  c->second_last_type = c->last_type;
  c->last_type = type;
  return type_code;
}
typedef struct {
  BlockTypeCodeCalculator calc;
  uint8_t type_depths[258];
  uint16_t type_bits[258];
  uint8_t length_depths[26];
  uint16_t length_bits[26];
} BlockSplitCode;
typedef struct {
  size_t histogram_length, num_block_types;
  const uint8_t* block_types;
  const uint32_t* block_lengths;
  size_t num_blocks;
  BlockSplitCode code;
  size_t block_ix, block_len, entropy_ix;
  uint8_t* depths;
  uint16_t* bits;
} BlockEncoder;

static void StoreBlockSwitch(BlockSplitCode* code, uint32_t block_len, uint8_t block_type, int is_first, size_t* ix,
                             uint8_t* st) {
  size_t typecode = NextBlockTypeCode(&code->calc, block_type);
  if (!is_first) WriteBits(code->type_depths[typecode], code->type_bits[typecode], ix, st);
  uint32_t lencode = BlockLengthPrefixCode(block_len);
  WriteBits(code->length_depths[lencode], code->length_bits[lencode], ix, st);
  WriteBits(kBlockLenNBits[lencode], block_len - kBlockLenOffset[lencode], ix, st);
}
static void BuildAndStoreBlockSplitCode(const uint8_t* types, const uint32_t* lengths, size_t num_blocks,
                                        size_t num_types, BlockSplitCode* code, size_t* ix, uint8_t* st) {
  uint32_t type_histo[258] = {0}, length_histo[26] = {0};
  BlockTypeCodeCalculator calc = {1, 0};
  for (size_t i = 0; i < num_blocks; ++i) {
    size_t type_code = NextBlockTypeCode(&calc, types[i]);
    if (i != 0) ++type_histo[type_code];
    ++length_histo[BlockLengthPrefixCode(lengths[i])];
  }
  StoreVarLenUint8(num_types - 1, ix, st);
  if (num_types > 1) {
    BuildAndStoreHuffmanTree(type_histo, num_types + 2, num_types + 2, code->type_depths, code->type_bits, ix, st);
    BuildAndStoreHuffmanTree(length_histo, 26, 26, code->length_depths, code->length_bits, ix, st);
    StoreBlockSwitch(code, lengths[0], types[0], 1, ix, st);
  }
}
static void BlockEncoderInit(BlockEncoder* e, size_t histogram_length, const BlockSplit* split) {
  memset(e, 0, sizeof(*e));
  e->histogram_length = histogram_length;
  e->num_block_types = split->num_types;
  e->block_types = split->types;
  e->block_lengths = split->lengths;
  e->num_blocks = split->num_blocks;
  e->code.calc.last_type = 1;
  e->code.calc.second_last_type = 0;
  e->block_len = split->num_blocks ? split->lengths[0] : 0;
}
static void BlockEncoderBuildCodes(BlockEncoder* e, const uint32_t* histograms, size_t histograms_size,
                                   size_t alphabet_size, size_t* ix, uint8_t* st) {
  size_t table_size = histograms_size * e->histogram_length;
  e->depths = (uint8_t*)calloc(table_size + 1, 1);
  e->bits = (uint16_t*)calloc(table_size + 1, 2);
  for (size_t i = 0; i < histograms_size; ++i) {
    size_t o = i * e->histogram_length;
    BuildAndStoreHuffmanTree(histograms + o, e->histogram_length, alphabet_size, e->depths + o, e->bits + o, ix, st);
  }
}
static inline void BlockEncoderStoreSymbol(BlockEncoder* e, size_t symbol, size_t* ix, uint8_t* st) {
  if (e->block_len == 0) {
    size_t b = ++e->block_ix;
    e->block_len = e->block_lengths[b];
    e->entropy_ix = (size_t)e->block_types[b] * e->histogram_length;
    if (getenv("ORACLE_DEBUG")) fprintf(stderr, "switch hl=%zu block %zu at bit %zu\n", e->histogram_length, b, *ix);
    StoreBlockSwitch(&e->code, e->block_lengths[b], e->block_types[b], 0, ix, st);
  }
  --e->block_len;
  size_t i = e->entropy_ix + symbol;
  WriteBits(e->depths[i], e->bits[i], ix, st);
}
static inline void BlockEncoderStoreSymbolWithContext(BlockEncoder* e, size_t symbol, size_t context,
                                                      const uint32_t* context_map, size_t* ix, uint8_t* st,
                                                      size_t context_bits) {
  if (e->block_len == 0) {
    size_t b = ++e->block_ix;
    e->block_len = e->block_lengths[b];
    e->entropy_ix = (size_t)e->block_types[b] << context_bits;
    StoreBlockSwitch(&e->code, e->block_lengths[b], e->block_types[b], 0, ix, st);
  }
  --e->block_len;
  size_t histo_ix = context_map[e->entropy_ix + context];
  size_t i = histo_ix * e->histogram_length + symbol;
  WriteBits(e->depths[i], e->bits[i], ix, st);
}
static void StoreTrivialContextMap(size_t num_types, size_t context_bits, size_t* ix, uint8_t* st) {
  StoreVarLenUint8(num_types - 1, ix, st);
  if (num_types > 1) {
    size_t repeat_code = context_bits - 1;
    size_t repeat_bits = ((size_t)1 << repeat_code) - 1;
    size_t alphabet_size = num_types + repeat_code;
    uint32_t histogram[272] = {0};
    uint8_t depths[272] = {0};
    uint16_t bits[272] = {0};
    WriteBits(1, 1, ix, st);
    WriteBits(4, repeat_code - 1, ix, st);
    histogram[repeat_code] = (uint32_t)num_types;
    histogram[0] = 1;
    for (size_t i = context_bits; i < alphabet_size; ++i) histogram[i] = 1;
    BuildAndStoreHuffmanTree(histogram, alphabet_size, alphabet_size, depths, bits, ix, st);
    for (size_t i = 0; i < num_types; ++i) {
      size_t code = i == 0 ? 0 : i + context_bits - 1;
      WriteBits(depths[code], bits[code], ix, st);
      WriteBits(depths[repeat_code], bits[repeat_code], ix, st);
      WriteBits((unsigned)repeat_code, repeat_bits, ix, st);
    }
    WriteBits(1, 1, ix, st);
  }
}
static void MoveToFrontTransform(const uint32_t* v_in, size_t v_size, uint32_t* v_out) {
  uint8_t mtf[256];
  if (v_size == 0) return;
  uint32_t max_value = v_in[0];
  for (size_t i = 1; i < v_size; ++i) if (v_in[i] > max_value) max_value = v_in[i];
  for (size_t i = 0; i <= max_value; ++i) mtf[i] = (uint8_t)i;
  size_t mtf_size = max_value + 1;
  for (size_t i = 0; i < v_size; ++i) {
    size_t index = 0;
    while (index < mtf_size && mtf[index] != (uint8_t)v_in[i]) ++index;
    v_out[i] = (uint32_t)index;
    uint8_t value = mtf[index];
    for (size_t k = index; k != 0; --k) mtf[k] = mtf[k - 1];
    mtf[0] = value;
  }
}
static void RunLengthCodeZeros(size_t in_size, uint32_t* v, size_t* out_size, uint32_t* max_run_length_prefix) {
  uint32_t max_reps = 0;
  for (size_t i = 0; i < in_size;) {
    uint32_t reps = 0;
    for (; i < in_size && v[i] != 0; ++i) {}
    for (; i < in_size && v[i] == 0; ++i) ++reps;
    max_reps = MAXZ(reps, max_reps);
  }
  uint32_t max_prefix = max_reps > 0 ? Log2FloorNonZero(max_reps) : 0;
  max_prefix = MINZ(max_prefix, *max_run_length_prefix);
  *max_run_length_prefix = max_prefix;
  *out_size = 0;
  for (size_t i = 0; i < in_size;) {
    if (v[i] != 0) {
      v[*out_size] = v[i] + *max_run_length_prefix;
      ++i;
      ++*out_size;
    } else {
      uint32_t reps = 1;
      for (size_t k = i + 1; k < in_size && v[k] == 0; ++k) ++reps;
      i += reps;
      while (reps != 0) {
        if (reps < (2u << max_prefix)) {
          uint32_t run_length_prefix = Log2FloorNonZero(reps);
          uint32_t extra_bits = reps - (1u << run_length_prefix);
          v[*out_size] = run_length_prefix + (extra_bits << 9);
          ++*out_size;
          break;
        } else {
          uint32_t extra_bits = (1u << max_prefix) - 1u;
          v[*out_size] = max_prefix + (extra_bits << 9);
          reps -= (2u << max_prefix) - 1u;
          ++*out_size;
        }
      }
    }
  }
}
static void EncodeContextMap(const uint32_t* context_map, size_t context_map_size, size_t num_clusters, size_t* ix,
                             uint8_t* st) {
  uint32_t max_run_length_prefix = 6;
  size_t num_rle_symbols = 0;
  uint32_t histogram[272] = {0};
  uint8_t depths[272] = {0};
  uint16_t bits[272] = {0};
  StoreVarLenUint8(num_clusters - 1, ix, st);
  if (num_clusters == 1) return;
  uint32_t* rle_symbols = (uint32_t*)malloc(context_map_size * 4);
  MoveToFrontTransform(context_map, context_map_size, rle_symbols);
  RunLengthCodeZeros(context_map_size, rle_symbols, &num_rle_symbols, &max_run_length_prefix);
  for (size_t i = 0; i < num_rle_symbols; ++i) ++histogram[rle_symbols[i] & 0x1ff];
  {
    int use_rle = max_run_length_prefix > 0;
    WriteBits(1, (uint64_t)use_rle, ix, st);
    if (use_rle) WriteBits(4, max_run_length_prefix - 1, ix, st);
  }
  BuildAndStoreHuffmanTree(histogram, num_clusters + max_run_length_prefix, num_clusters + max_run_length_prefix,
                           depths, bits, ix, st);
  for (size_t i = 0; i < num_rle_symbols; ++i) {
    uint32_t rle_symbol = rle_symbols[i] & 0x1ff;
    uint32_t extra_bits_val = rle_symbols[i] >> 9;
    WriteBits(depths[rle_symbol], bits[rle_symbol], ix, st);
    if (rle_symbol > 0 && rle_symbol <= max_run_length_prefix) WriteBits(rle_symbol, extra_bits_val, ix, st);
  }
  WriteBits(1, 1, ix, st);
  free(rle_symbols);
}
static void StoreCommandExtra(const Command* cmd, size_t* ix, uint8_t* st) {
  uint32_t copylen_code = CommandCopyLenCode(cmd);
  uint16_t inscode = oracle_insert_length_code(cmd->insert_len_);
  uint16_t copycode = oracle_copy_length_code(copylen_code);
  uint32_t insnumextra = kInsExtra[inscode];
  uint64_t insextraval = cmd->insert_len_ - kInsBase[inscode];
  uint64_t copyextraval = copylen_code - kCopyBase[copycode];
  WriteBits(insnumextra + kCopyExtra[copycode], (copyextraval << insnumextra) | insextraval, ix, st);
}

static void StoreMetaBlock(const uint8_t* input, size_t start_pos, size_t length, uint8_t prev_byte,
                           uint8_t prev_byte2, int is_last, const Command* commands, size_t n_commands,
                           MetaBlockSplit* mb, size_t* ix, uint8_t* st) {
  size_t pos = start_pos;
  BlockEncoder lit, cmd, dist;
  StoreCompressedMetaBlockHeader(is_last, length, ix, st);
  BlockEncoderInit(&lit, 256, &mb->literal_split);
  BlockEncoderInit(&cmd, 704, &mb->command_split);
  BlockEncoderInit(&dist, 64, &mb->distance_split);
  BuildAndStoreBlockSplitCode(lit.block_types, lit.block_lengths, lit.num_blocks, lit.num_block_types, &lit.code, ix, st);
  BuildAndStoreBlockSplitCode(cmd.block_types, cmd.block_lengths, cmd.num_blocks, cmd.num_block_types, &cmd.code, ix, st);
  BuildAndStoreBlockSplitCode(dist.block_types, dist.block_lengths, dist.num_blocks, dist.num_block_types, &dist.code, ix, st);
  WriteBits(2, 0, ix, st); /* NPOSTFIX */
  WriteBits(4, 0, ix, st); /* NDIRECT >> NPOSTFIX */
  for (size_t i = 0; i < mb->literal_split.num_types; ++i) WriteBits(2, 2 /* CONTEXT_UTF8 */, ix, st);
  if (mb->literal_context_map_size == 0) StoreTrivialContextMap(mb->literal_histograms_size, 6, ix, st);
  else EncodeContextMap(mb->literal_context_map, mb->literal_context_map_size, mb->literal_histograms_size, ix, st);
  StoreTrivialContextMap(mb->distance_histograms_size, 2, ix, st);
  BlockEncoderBuildCodes(&lit, mb->literal_histograms, mb->literal_histograms_size, 256, ix, st);
  BlockEncoderBuildCodes(&cmd, mb->command_histograms, mb->command_histograms_size, 704, ix, st);
  BlockEncoderBuildCodes(&dist, mb->distance_histograms, mb->distance_histograms_size, 64, ix, st);
  for (size_t i = 0; i < n_commands; ++i) {
    const Command c = commands[i];
    BlockEncoderStoreSymbol(&cmd, c.cmd_prefix_, ix, st);
    StoreCommandExtra(&c, ix, st);
    if (mb->literal_context_map_size == 0) {
      for (size_t j = c.insert_len_; j != 0; --j) BlockEncoderStoreSymbol(&lit, input[pos++], ix, st);
    } else {
      for (size_t j = c.insert_len_; j != 0; --j) {
        size_t context = ContextUTF8(prev_byte, prev_byte2);
        uint8_t literal = input[pos];
        BlockEncoderStoreSymbolWithContext(&lit, literal, context, mb->literal_context_map, ix, st, 6);
        prev_byte2 = prev_byte;
        prev_byte = literal;
        ++pos;
      }
    }
    pos += CommandCopyLen(&c);
    if (CommandCopyLen(&c)) {
      prev_byte2 = input[pos - 2];
      prev_byte = input[pos - 1];
      if (c.cmd_prefix_ >= 128) {
        BlockEncoderStoreSymbol(&dist, c.dist_prefix_ & 0x3ff, ix, st);
        WriteBits(c.dist_prefix_ >> 10, c.dist_extra_, ix, st);
      }
    }
  }
  if (getenv("ORACLE_DEBUG")) {
    size_t tl = 0, tc = 0; for (size_t i = 0; i < n_commands; ++i) { tl += commands[i].insert_len_; tc += CommandCopyLen(&commands[i]); }
    fprintf(stderr, "emitted: lit block_ix=%zu remaining block_len=%zu total lits=%zu copies=%zu\n", lit.block_ix, lit.block_len, tl, tc);
  }
  if (getenv("ORACLE_DEBUG") && pos != start_pos + length)
    fprintf(stderr, "COVERAGE MISMATCH pos=%zu expected=%zu\n", pos, start_pos + length);
  free(lit.depths); free(lit.bits); free(cmd.depths); free(cmd.bits); free(dist.depths); free(dist.bits);
  if (is_last) JumpToByteBoundary(ix, st);
}
static void StoreUncompressedMetaBlock(int is_final, const uint8_t* input, size_t position, size_t len, size_t* ix,
                                       uint8_t* st) {
  StoreUncompressedMetaBlockHeader(len, ix, st); /* brotli_bit_stream.rs:2775-2833 */
  JumpToByteBoundary(ix, st);
  memcpy(st + (*ix >> 3), input + position, len);
  *ix += len << 3;
  st[*ix >> 3] = 0;
  if (is_final) {
    WriteBits(1, 1, ix, st);
    WriteBits(1, 1, ix, st);
    JumpToByteBoundary(ix, st);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Encoder driver: encode.rs:546-625 (params), :834-893 (hasher choice), :1325-1354 (should_compress),
 * :1941-2167 (WriteMetaBlockInternal), :2214-2543 (encode_data), :360-400 (extend_last_command).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  /* out-parameters for tests */
  size_t num_metablocks, num_commands_total, num_literals_total;
  int hasher_type, bucket_bits, block_bits, hash_len, n_last;
} OracleStats;

static int ShouldCompress(const uint8_t* data, size_t last_flush_pos, size_t bytes, size_t num_literals,
                          size_t num_commands) {
  if (num_commands < (bytes >> 8) + 2 && (float)num_literals > 0.99f * (float)bytes) {
    uint32_t literal_histo[256] = {0};
    const uint32_t kSampleRate = 13;
    const float kMinEntropy = 7.92f;
    const float bit_cost_threshold = (float)bytes * kMinEntropy / (float)kSampleRate;
    size_t t = (bytes + kSampleRate - 1) / kSampleRate;
    size_t pos = last_flush_pos;
    for (size_t i = 0; i < t; ++i) {
      ++literal_histo[data[pos]];
      pos += kSampleRate;
    }
    if (BitsEntropy(literal_histo, 256) > bit_cost_threshold) return 0;
  }
  return 1;
}

/* flags */
#define ORACLE_FLAG_NO_CONTEXT_MODELING 1
#define ORACLE_FLAG_NO_DICTIONARY 2

size_t oracle_brotli_compress(int quality, int lgwin, const uint8_t* input_in, size_t input_size, uint8_t* out,
                              size_t out_cap, size_t size_hint, int flags, OracleStats* stats) {
  init_tables();
  init_context_luts();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (out_cap < input_size + (input_size >> 3) + 1024) return 0;
  if (input_size == 0) { out[0] = 6; return 1; } /* encode.rs:1463-1467 */
  if (quality < 4) quality = 4;
  if (quality > 9) quality = 9;
  if (lgwin < 10) lgwin = 10;
  if (lgwin > 24) lgwin = 24;
  if (size_hint == 0) size_hint = input_size;
  /* flat copy with slack so that 8-byte hash loads near the end stay in bounds */
  uint8_t* data = (uint8_t*)calloc(input_size + 64, 1);
  memcpy(data, input_in, input_size);

  int lgblock = 16; /* ComputeLgBlock encode.rs:570-585 */
  if (quality >= 9 && lgwin > lgblock) lgblock = MINZ(18, lgwin);
  const size_t block_size = (size_t)1 << lgblock;

  Hasher h; /* ChooseHasher encode.rs:834-893 (H40-42 are unsupported there and fall back to H6, :1096-1114) */
  memset(&h, 0, sizeof(h));
  h.use_dictionary = !(flags & ORACLE_FLAG_NO_DICTIONARY);
  if (quality == 9) {
    h.type = 9; h.n_last = 16; h.block_bits = 8; h.bucket_bits = 15; h.hash_len = 4;
  } else if (lgwin <= 16) {
    /* type 40/41/42 requested -> BrotliMakeHasher falls back to InitializeH6 with the *default* hasher params
       (encode.rs:318-357: bucket_bits 15, block_bits 8, hash_len 5, 16 last distances). */
    h.type = 6; h.block_bits = 8; h.bucket_bits = 15; h.hash_len = 5; h.n_last = 16;
  } else if (size_hint > (1u << 22) && lgwin >= 19) {
    h.type = 6; h.block_bits = MINZ(quality - 1, 9); h.bucket_bits = 15; h.hash_len = 5;
    h.n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  } else {
    h.type = 5; h.block_bits = MINZ(quality - 1, 9);
    h.bucket_bits = (quality < 7 && size_hint <= (1u << 20)) ? 14 : 15;
    h.hash_len = 4;
    h.n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  }
  h.block_mask = (1u << h.block_bits) - 1u;
  h.hash_mask = ~0ull >> (64 - 8 * h.hash_len);
  h.num = (uint16_t*)calloc((size_t)1 << h.bucket_bits, 2);
  h.buckets = (uint32_t*)calloc((size_t)1 << (h.bucket_bits + h.block_bits), 4);
  if (stats) { stats->hasher_type = h.type; stats->bucket_bits = h.bucket_bits; stats->block_bits = h.block_bits;
               stats->hash_len = h.hash_len; stats->n_last = h.n_last; }

  size_t cmd_cap = input_size / 2 + block_size + 16;
  Command* commands = (Command*)malloc(cmd_cap * sizeof(Command));
  size_t num_commands = 0, num_literals = 0, last_insert_len = 0;
  int dist_cache[16] = {4, 11, 15, 16}; /* encode.rs:693-703 default */
  int saved_dist_cache[4] = {4, 11, 15, 16};
  uint8_t prev_byte = 0, prev_byte2 = 0;
  size_t last_flush_pos = 0;

  size_t ix = 0;
  memset(out, 0, 16);
  { /* EncodeWindowBits encode.rs:603-625 */
    if (lgwin == 16) WriteBits(1, 0, &ix, out);
    else if (lgwin == 17) WriteBits(7, 1, &ix, out);
    else if (lgwin > 17) WriteBits(4, (uint64_t)(((lgwin - 17) << 1) | 1), &ix, out);
    else WriteBits(7, (uint64_t)(((lgwin - 8) << 4) | 1), &ix, out);
  }
  const size_t max_metablock = (size_t)1 << MINZ(1 + MAXZ(lgwin, lgblock), 24); /* encode.rs:1713 */

  for (size_t block_start = 0; block_start < input_size; block_start += block_size) {
    size_t bytes = MINZ(block_size, input_size - block_start);
    const int is_last = block_start + bytes == input_size;
    const size_t input_pos = block_start + bytes;
    size_t position = block_start;
    /* StitchToPreviousBlock mod.rs:210-222 */
    if (bytes >= HashTypeLength(&h) - 1 && position >= 3) {
      HasherStore(&h, data, position - 3);
      HasherStore(&h, data, position - 2);
      HasherStore(&h, data, position - 1);
    }
    if (num_commands && last_insert_len == 0) { /* extend_last_command encode.rs:360-400 */
      Command* last = &commands[num_commands - 1];
      const size_t max_backward_distance = ((size_t)1 << lgwin) - 16;
      const size_t last_copy_len = last->copy_len_ & 0x01ffffff;
      const size_t last_processed = position - last_copy_len;
      const size_t max_distance = MINZ(last_processed, max_backward_distance);
      const size_t cmd_dist = (size_t)dist_cache[0];
      const uint32_t distance_code = CommandRestoreDistanceCode(last);
      if (distance_code < 16 || distance_code - 15 == cmd_dist) {
        if (cmd_dist <= max_distance) {
          while (bytes != 0 && data[position] == data[position - cmd_dist]) {
            last->copy_len_++;
            bytes--;
            position++;
          }
        }
        last->cmd_prefix_ = GetLengthCode(last->insert_len_,
                                          (size_t)((int)(last->copy_len_ & 0x01ffffff) + (int)(last->copy_len_ >> 25)),
                                          (last->dist_prefix_ & 0x3ff) == 0);
      }
    }
    CreateBackwardReferences(bytes, position, data, quality, lgwin, &h, dist_cache, &last_insert_len, commands,
                             &num_commands, &num_literals);
    {
      const size_t max_literals = max_metablock / 8, max_commands = max_metablock / 8;
      const size_t processed_bytes = input_pos - last_flush_pos;
      const int next_input_fits = processed_bytes + block_size <= max_metablock;
      if (!is_last && next_input_fits && num_literals < max_literals && num_commands < max_commands) continue;
    }
    if (last_insert_len > 0) {
      CommandInitInsert(&commands[num_commands++], last_insert_len);
      num_literals += last_insert_len;
      last_insert_len = 0;
    }
    /* ---- WriteMetaBlockInternal encode.rs:1941-2167 (non-appendable stream) ---- */
    {
      const size_t mb_bytes = input_pos - last_flush_pos;
      if (stats) { stats->num_metablocks++; stats->num_commands_total += num_commands; stats->num_literals_total += num_literals; }
      if (!ShouldCompress(data, last_flush_pos, mb_bytes, num_literals, num_commands)) {
        memcpy(dist_cache, saved_dist_cache, sizeof(saved_dist_cache));
        StoreUncompressedMetaBlock(is_last, data, last_flush_pos, mb_bytes, &ix, out);
      } else {
        const size_t saved_ix = ix;
        const uint8_t saved0 = out[ix >> 3], saved1 = out[(ix >> 3) + 1];
        MetaBlockSplit mb;
        memset(&mb, 0, sizeof(mb));
        size_t num_literal_contexts = 1;
        const uint32_t* literal_context_map = NULL;
        if (!(flags & ORACLE_FLAG_NO_CONTEXT_MODELING))
          DecideOverLiteralContextModeling(data, last_flush_pos, mb_bytes, quality, size_hint, &num_literal_contexts,
                                           &literal_context_map);
        BuildMetaBlockGreedy(data, last_flush_pos, prev_byte, prev_byte2, num_literal_contexts, literal_context_map,
                             commands, num_commands, &mb);
        /* BrotliOptimizeHistograms metablock.rs:1076-1108 */
        for (size_t i = 0; i < mb.literal_histograms_size; ++i)
          oracle_optimize_huffman_counts_for_rle(256, mb.literal_histograms + i * 256);
        for (size_t i = 0; i < mb.command_histograms_size; ++i)
          oracle_optimize_huffman_counts_for_rle(704, mb.command_histograms + i * 704);
        for (size_t i = 0; i < mb.distance_histograms_size; ++i)
          oracle_optimize_huffman_counts_for_rle(64, mb.distance_histograms + i * 64);
        if (getenv("ORACLE_DEBUG"))
          fprintf(stderr, "mb bytes=%zu cmds=%zu lit types=%zu blocks=%zu | cmd types=%zu blocks=%zu | dist types=%zu blocks=%zu ctx=%zu\n",
                  mb_bytes, num_commands, mb.literal_split.num_types, mb.literal_split.num_blocks,
                  mb.command_split.num_types, mb.command_split.num_blocks, mb.distance_split.num_types,
                  mb.distance_split.num_blocks, num_literal_contexts);
        if (getenv("ORACLE_DEBUG")) {
          for (size_t b = 0; b < mb.literal_split.num_blocks; ++b)
            fprintf(stderr, " lit block %zu type %u len %u\n", b, mb.literal_split.types[b], mb.literal_split.lengths[b]);
          for (size_t b = 0; b < mb.distance_split.num_blocks; ++b)
            fprintf(stderr, " dist block %zu type %u len %u\n", b, mb.distance_split.types[b], mb.distance_split.lengths[b]);
        }
        StoreMetaBlock(data, last_flush_pos, mb_bytes, prev_byte, prev_byte2, is_last, commands, num_commands, &mb, &ix,
                       out);
        MetaBlockSplitFree(&mb);
        if (mb_bytes + 4 + (saved_ix >> 3) < (ix >> 3)) { /* encode.rs:2141-2163 */
          memcpy(dist_cache, saved_dist_cache, sizeof(saved_dist_cache));
          memset(out + (saved_ix >> 3), 0, (ix >> 3) - (saved_ix >> 3) + 9);
          out[saved_ix >> 3] = saved0 & (uint8_t)((1u << (saved_ix & 7)) - 1u);
          (void)saved1;
          ix = saved_ix;
          StoreUncompressedMetaBlock(is_last, data, last_flush_pos, mb_bytes, &ix, out);
        }
      }
      last_flush_pos = input_pos;
      if (last_flush_pos > 0) prev_byte = data[last_flush_pos - 1];
      if (last_flush_pos > 1) prev_byte2 = data[last_flush_pos - 2];
      num_commands = 0;
      num_literals = 0;
      memcpy(saved_dist_cache, dist_cache, sizeof(saved_dist_cache));
    }
  }
  free(commands);
  free(h.num);
  free(h.buckets);
  free(data);
  return (ix + 7) >> 3;
}

/* test hook: run only the LZ77 stage over a whole buffer (one call per 1<<lgblock block, as encode_data does) and
 * return the commands; used to diff GPU command streams against the reference's greedy parse. */
size_t oracle_backward_references(int quality, int lgwin, const uint8_t* input_in, size_t input_size, size_t size_hint,
                                  uint32_t* out_cmds /* 5 u32 per command */, size_t cap_cmds, size_t* last_insert) {
  init_tables();
  init_context_luts();
  /* compress into a scratch buffer with a stats hook would not expose commands; re-run the LZ77 loop directly */
  if (size_hint == 0) size_hint = input_size;
  uint8_t* data = (uint8_t*)calloc(input_size + 64, 1);
  memcpy(data, input_in, input_size);
  int lgblock = 16;
  if (quality >= 9 && lgwin > lgblock) lgblock = MINZ(18, lgwin);
  const size_t block_size = (size_t)1 << lgblock;
  Hasher h;
  memset(&h, 0, sizeof(h));
  if (quality == 9) { h.type = 9; h.n_last = 16; h.block_bits = 8; h.bucket_bits = 15; h.hash_len = 4; }
  else if (lgwin <= 16) { h.type = 6; h.block_bits = 8; h.bucket_bits = 15; h.hash_len = 5; h.n_last = 16; }
  else if (size_hint > (1u << 22) && lgwin >= 19) {
    h.type = 6; h.block_bits = MINZ(quality - 1, 9); h.bucket_bits = 15; h.hash_len = 5;
    h.n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  } else {
    h.type = 5; h.block_bits = MINZ(quality - 1, 9);
    h.bucket_bits = (quality < 7 && size_hint <= (1u << 20)) ? 14 : 15; h.hash_len = 4;
    h.n_last = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  }
  h.block_mask = (1u << h.block_bits) - 1u;
  h.hash_mask = ~0ull >> (64 - 8 * h.hash_len);
  h.num = (uint16_t*)calloc((size_t)1 << h.bucket_bits, 2);
  h.buckets = (uint32_t*)calloc((size_t)1 << (h.bucket_bits + h.block_bits), 4);
  Command* commands = (Command*)malloc((input_size / 2 + block_size + 16) * sizeof(Command));
  size_t num_commands = 0, num_literals = 0, last_insert_len = 0;
  int dist_cache[16] = {4, 11, 15, 16};
  for (size_t block_start = 0; block_start < input_size; block_start += block_size) {
    size_t bytes = MINZ(block_size, input_size - block_start);
    size_t position = block_start;
    if (bytes >= HashTypeLength(&h) - 1 && position >= 3) {
      HasherStore(&h, data, position - 3); HasherStore(&h, data, position - 2); HasherStore(&h, data, position - 1);
    }
    CreateBackwardReferences(bytes, position, data, quality, lgwin, &h, dist_cache, &last_insert_len, commands,
                             &num_commands, &num_literals);
  }
  size_t n = MINZ(num_commands, cap_cmds);
  for (size_t i = 0; i < n; ++i) {
    out_cmds[5 * i + 0] = commands[i].insert_len_; out_cmds[5 * i + 1] = commands[i].copy_len_;
    out_cmds[5 * i + 2] = commands[i].dist_extra_; out_cmds[5 * i + 3] = commands[i].cmd_prefix_;
    out_cmds[5 * i + 4] = commands[i].dist_prefix_;
  }
  *last_insert = last_insert_len;
  free(commands); free(h.num); free(h.buckets); free(data);
  return num_commands;
}
