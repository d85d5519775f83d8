// bro_dict.cuh -- static-dictionary matches (RFC 7932 section 8 / appendix A) for the parse stage.
//
// Reference behaviour (backward_references/mod.rs:1896-1988, called from FindLongestMatch :1797 when neither the
// distance cache nor the bucket gave a match): hash the 4 bytes at the position, look at two table slots, accept a word
// if at least len - 9 of its bytes match (identity or one of the nine "omit last N" transforms), encode it as a distance
// beyond the window: max_backward + 1 + word_index + (transform_id << size_bits[len]).
//
// Differences: the hash table is this library's own (gen_dict.py), and the reference's running lookups/matches counter (a
// sequential, stream-global heuristic that switches the search off on input without dictionary words) has no parallel
// equivalent and is dropped -- the lookup is done by the position-parallel match kernel where it costs little.
//
// Included by bro_parse.cuh (after Match / score_regular, before find_match); do not include directly.
#pragma once

namespace bro {

struct DictView {
  const uint8_t* words;   // 122784 bytes
  const uint16_t* hash;   // 32768 entries: word_index << 5 | length, 0 = empty
};

// packed copy length of a raw / final command: bits 0..24 bytes covered, 25..30 (word length - bytes covered), 31 dictionary
#define BRO_LEN_MASK 0x1FFFFFFu
BRO_HD uint32_t pack_dict_len(uint32_t matchlen, uint32_t wordlen) { return matchlen | ((wordlen - matchlen) << 25) | 0x80000000u; }
BRO_HD uint32_t len_bytes(uint32_t packed) { return packed & BRO_LEN_MASK; }
BRO_HD uint32_t len_coded(uint32_t packed) { return (packed & BRO_LEN_MASK) + ((packed >> 25) & 0x3Fu); }
BRO_HD bool len_is_dict(uint32_t packed) { return (packed >> 31) != 0; }

BRO_HD uint32_t dict_size_bits(uint32_t len) {  // NDBITS, RFC 7932 section 8
  // lengths 4..24: 10,10,11,11,10,10,10,10,10,9,9,8,7,7,8,7,7,6,6,5,5 packed 4 bits each
  const uint64_t lo = 0x899AAAAABBAA0000ull;   // lengths 0..15
  const uint64_t hi = 0x0000000556677877ull;   // lengths 16..31
  return (uint32_t)(((len < 16 ? lo : hi) >> ((len & 15u) * 4u)) & 0xFu);
}
BRO_HD uint32_t dict_offset(uint32_t len) {  // offset of the first word of this length = sum over l < len of l << NDBITS[l]
  static constexpr uint32_t kOff[32] = {0, 0, 0, 0, 0, 4096, 9216, 21504, 35840, 44032, 53248, 63488, 74752, 87040, 93696, 100864,
                                        104704, 106752, 108928, 113536, 115968, 118528, 119872, 121280, 122016, 122784, 122784,
                                        122784, 122784, 122784, 122784, 122784};
  return kOff[len & 31u];
}
BRO_HD uint32_t dict_omit_last_transform(uint32_t cut) {  // RFC 7932 appendix B: identity, OmitLast1..9
  switch (cut) {
    case 0: return 0; case 1: return 12; case 2: return 27; case 3: return 23; case 4: return 42;
    case 5: return 63; case 6: return 56; case 7: return 48; case 8: return 59; default: return 64;
  }
}
BRO_HD uint32_t dict_hash14(uint32_t w) { return (w * 0x1e35a7bdu) >> 18; }

// The dictionary candidate of a position is found by the match stage (only where the bucket search found nothing) and
// travels to the parse in best[p]:
//   bucket match : dist << 8 | len            (len 4..64, bit 7 clear)
//   dictionary   : cut << 26 | word_id << 8 | 0x80 | bytes matched      word_id = word_index + (transform << NDBITS[len])
// The parse uses it only if the distance cache gave nothing either (mod.rs:1797 "if !is_match_found").
#define BRO_BEST_DICT 0x80u
BRO_HD uint32_t best_pack_dict(uint32_t ml, uint32_t wl, uint32_t idx) {
  const uint32_t word_id = idx + (dict_omit_last_transform(wl - ml) << dict_size_bits(wl));
  return ((wl - ml) << 26) | (word_id << 8) | BRO_BEST_DICT | ml;
}
// Tries the two slots of the position's bucket (later slot wins ties, as in SearchInStaticDictionary); returns the packed
// candidate or 0.  max_len = bytes left in the range, max_backward_here = min(absolute position, window - 16).
BRO_HD uint32_t dict_candidate(const DictView& D, int hash_type, const uint8_t* cur, uint32_t max_len, uint32_t max_backward_here) {
  uint32_t best = 0, best_score = BRO_MIN_SCORE;
  const uint32_t key = dict_hash14(load32(cur)) << 1;
  for (uint32_t s = 0; s < 2; ++s) {
    const uint32_t item = D.hash[key + s];
    if (item == 0) continue;
    const uint32_t len = item & 31u, idx = item >> 5;
    if (len > max_len) continue;
    const uint8_t* w = D.words + dict_offset(len) + len * idx;
    uint32_t ml = 0;
    while (ml < len && cur[ml] == w[ml]) ++ml;
    if (ml + 10u <= len || ml < 4) continue;  // ml < 4: the slot was reached through a hash collision
    const uint32_t word_id = idx + (dict_omit_last_transform(len - ml) << dict_size_bits(len));
    const uint32_t score = score_regular(hash_type, ml, max_backward_here + 1u + word_id);
    if (score < best_score) continue;
    best = best_pack_dict(ml, len, idx);
    best_score = score;
  }
  return best;
}
// Parse side: decodes a dictionary candidate of best[] for a position with max_len bytes left in its unit.
BRO_HD bool dict_decode(uint32_t b, int hash_type, uint32_t max_len, uint32_t max_backward_here, Match* m) {
  const uint32_t ml = b & 0x7Fu, wl = ml + ((b >> 26) & 0xFu);
  if (wl > max_len) return false;
  m->len = pack_dict_len(ml, wl);
  m->dist = max_backward_here + 1u + ((b >> 8) & 0x3FFFFu);
  m->score = score_regular(hash_type, ml, m->dist);
  return true;
}

}  // namespace bro
