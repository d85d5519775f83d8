// bro_dict.cuh -- static-dictionary matches (RFC 7932 section 8 / appendix A) for the parse stage.
//
// Reference behaviour (backward_references/mod.rs:1896-1988, called from FindLongestMatch :1797 when neither the
// distance cache nor the bucket gave a match): hash the 4 bytes at the position, look at two table slots, accept a word
// if at least len - 9 of its bytes match (identity or one of the nine "omit last N" transforms), encode it as a distance
// beyond the window: max_backward + 1 + word_index + (transform_id << size_bits[len]).
//
// Differences, both forced by parallelism: the hash table is this library's own (gen_dict.py), and the reference's
// running lookups/matches counter (a sequential, stream-global heuristic that switches the search off on input without
// dictionary words) is replaced by a per-unit gate that is a pure function of the unit's bytes: 64 sample positions
// are probed, the dictionary is searched in this unit iff one of them hits.
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

// Tries the two slots of the position's bucket.  m->score is the score to reach (BRO_MIN_SCORE when nothing was found);
// on success m->len = pack_dict_len(bytes matched, word length), m->dist = the dictionary distance.
BRO_HD bool dict_search(const DictView& D, int hash_type, const uint8_t* cur, uint32_t max_len, uint32_t max_backward_here,
                        Match* m) {
  bool found = false;
  const uint32_t key = dict_hash14(load32(cur)) << 1;
  for (uint32_t s = 0; s < 2; ++s) {
    const uint32_t item = D.hash[key + s];
    if (item == 0) continue;
    const uint32_t len = item & 31u, idx = item >> 5;
    if (len > max_len) continue;
    const uint8_t* w = D.words + dict_offset(len) + len * idx;
    uint32_t ml = 0;
    while (ml < len && cur[ml] == w[ml]) ++ml;
    if (ml + 10u <= len || ml < 4) continue;  // ml < 4: the slot was reached through a hash collision (the reference would accept
                                              // such a 1..3 byte OmitLast match when the window is tiny; not worth a byte-wise path)
    const uint32_t backward = max_backward_here + 1u + idx + (dict_omit_last_transform(len - ml) << dict_size_bits(len));
    if (backward > 0x3FFFFFCu) continue;
    const uint32_t score = score_regular(hash_type, ml, backward);
    if (score < m->score) continue;
    m->len = pack_dict_len(ml, len);
    m->dist = backward;
    m->score = score;
    found = true;
  }
  return found;
}

// sample k of the per-unit gate: does position ustart + 64 k start a usable dictionary word ?
BRO_HD bool dict_gate_sample(const DictView& D, int hash_type, const uint8_t* data, uint32_t ustart, uint32_t uend, uint32_t k) {
  const uint32_t p = ustart + 64u * k;
  if (p + 8u > uend) return false;
  Match m;
  m.len = m.dist = 0;
  m.score = BRO_MIN_SCORE;
  return dict_search(D, hash_type, data + p, uend - p, 0x3FFFF0u, &m);
}
BRO_HD bool dict_unit_gate(const DictView& D, int hash_type, const uint8_t* data, uint32_t ustart, uint32_t uend) {
  for (uint32_t k = 0; k < 64; ++k)
    if (dict_gate_sample(D, hash_type, data, ustart, uend, k)) return true;
  return false;
}

}  // namespace bro
