// bro_finalize.cuh -- turn the per-unit raw matches into the metablock's final command records.
//
// One thread per parse unit.  Work that the reference does in one sequential sweep (Command::init
// command.rs:273, ComputeDistanceCode :48, the dist-cache pushes in CreateBackwardReferences
// mod.rs:2495-2503, extend_last_command encode.rs:360-400, the trailing insert-only command encode.rs:2478)
// is made unit-parallel:
//   * literal carry: the trailing literals of a unit become part of the next unit's first command (or of a final
//     insert-only command at the metablock end);
//   * copy continuation: if a unit's first match continues the previous unit's last copy (no literals in between,
//     same distance) it is absorbed into that command -- this is what heals the truncation of matches at unit
//     boundaries (the analogue of extend_last_command);
//   * distance cache: the cache is a pure function of the distance sequence (cache[0] is always the previous
//     command's distance; a distance is pushed iff it differs from it), so each unit reconstructs its incoming
//     cache by looking back over earlier commands of the same metablock.  A metablock starts with an unknown
//     cache (short codes are only used against distances seen inside the metablock), which makes metablocks
//     independent of each other and lets any of them be stored raw.
#pragma once
#include "bro_common.cuh"
#include "bro_parse.cuh"

namespace bro {

struct GCmd {  // final command record in device memory (32 bytes)
  uint32_t insert_len;
  uint32_t copy_len;    // packed like RawCmd::copy_len (bro_dict.cuh): bytes covered | word-length delta | dictionary flag
  uint32_t dist_extra;
  uint16_t cmd_prefix;
  uint16_t dist_prefix;
  uint32_t lit_idx;   // rank of its first literal among the metablock's literals
  uint32_t dist_idx;  // rank of its distance symbol among the metablock's distance symbols
  uint32_t pos;       // input position of its first literal
  uint32_t pad;       // owning parse unit (device: dist_idx is unit-relative until unit_dist_off[pad] is added)
  BRO_HD Cmd as_cmd() const {
    Cmd c;
    c.insert_len = insert_len; c.copy_len = len_coded(copy_len); c.dist_extra = dist_extra;  // Cmd carries the coded length
    c.cmd_prefix = cmd_prefix; c.dist_prefix = dist_prefix;
    return c;
  }
};

struct UnitView {
  const RawCmd* raw;         // [num_units][cu]
  const uint32_t* ncmd;      // raw commands per unit
  const uint32_t* tail;      // trailing literals per unit
  uint32_t cu;               // raw command slots per unit
  uint32_t unit;             // unit size
  uint32_t n;                // input size
};

BRO_HD uint32_t unit_carry_in(const UnitView& V, uint32_t u0, uint32_t u) {
  uint32_t carry = 0;
  while (u > u0) {
    --u;
    carry += V.tail[u];
    if (V.ncmd[u] != 0) break;
  }
  return carry;
}
// does the first raw command of unit u continue the last copy of unit u-1 ?
BRO_HD bool unit_absorbed(const UnitView& V, uint32_t u0, uint32_t u) {
  if (u == u0 || V.ncmd[u] == 0) return false;
  const RawCmd& f = V.raw[(size_t)u * V.cu];
  if (f.insert_len != 0 || len_is_dict(f.copy_len)) return false;
  if (V.tail[u - 1] != 0 || V.ncmd[u - 1] == 0) return false;
  const RawCmd& l = V.raw[(size_t)(u - 1) * V.cu + V.ncmd[u - 1] - 1];
  return l.distance == f.distance && !len_is_dict(l.copy_len);
}
// number of final commands a unit contributes to its metablock (u1 = one past the metablock's last unit)
BRO_HD uint32_t unit_final_ncmd(const UnitView& V, uint32_t u0, uint32_t u1, uint32_t u) {
  uint32_t n = V.ncmd[u] - (unit_absorbed(V, u0, u) ? 1u : 0u);
  if (u + 1 == u1) {
    uint32_t carry_out = V.tail[u] + (V.ncmd[u] == 0 ? unit_carry_in(V, u0, u) : 0u);
    if (carry_out) ++n;
  }
  return n;
}

// Writes the final commands of unit u at out[0 .. unit_final_ncmd) ; lit_base = rank of the unit's first literal
// (sum over earlier units of unit_len - copied bytes).  Returns the number of distance symbols of the unit;
// GCmd::dist_idx is written unit-relative (the caller adds the unit's prefix afterwards).
BRO_HD_NOINLINE uint32_t finalize_unit(const UnitView& V, uint32_t u0, uint32_t u1, uint32_t u, uint32_t lit_base,
                                       GCmd* out) {
  const uint32_t ustart = u * V.unit;
  const uint32_t nraw = V.ncmd[u];
  const RawCmd* rc = V.raw + (size_t)u * V.cu;
  const bool absorbed = unit_absorbed(V, u0, u);
  uint32_t carry = unit_carry_in(V, u0, u);
  // incoming distance cache: collapse runs of equal distances walking backwards over earlier commands
  int32_t dc[4] = {0x3fffffff, 0x3fffffff, 0x3fffffff, 0x3fffffff};
  {
    int k = 0;
    uint32_t last = 0;  // 0 is never a valid distance
    for (uint32_t v = u; v > u0 && k < 4;) {
      --v;
      for (uint32_t i = V.ncmd[v]; i > 0 && k < 4;) {
        --i;
        if (len_is_dict(V.raw[(size_t)v * V.cu + i].copy_len)) continue;  // not part of the distance sequence
        uint32_t d = V.raw[(size_t)v * V.cu + i].distance;
        if (d != last) { dc[k++] = (int32_t)d; last = d; }
      }
    }
  }
  uint32_t nout = 0, ndist = 0, lit_idx = lit_base, pos = ustart;
  for (uint32_t i = 0; i < nraw; ++i) {
    uint32_t ins = rc[i].insert_len, len = rc[i].copy_len, dist = rc[i].distance;
    const bool is_dict = len_is_dict(len);
    if (i == 0) {
      if (absorbed) {  // emitted by the owner in an earlier unit
        pos += len;
        continue;
      }
      ins += carry;
    }
    if (i + 1 == nraw) {  // last command of the unit: absorb continuations from following units
      for (uint32_t v = u + 1; v < u1 && unit_absorbed(V, u0, v); ++v) {
        len += V.raw[(size_t)v * V.cu].copy_len;
        if (!(V.ncmd[v] == 1 && V.tail[v] == 0)) break;
      }
    }
    uint32_t code = is_dict ? dist + 15u : compute_distance_code(dist, dc);
    if (code != 0 && !is_dict) { dc[3] = dc[2]; dc[2] = dc[1]; dc[1] = dc[0]; dc[0] = (int32_t)dist; }
    uint32_t sym_nbits, extra;
    prefix_encode_copy_distance(code, &sym_nbits, &extra);
    GCmd g;
    g.insert_len = ins;
    g.copy_len = len;
    g.dist_extra = extra;
    g.dist_prefix = (uint16_t)sym_nbits;
    g.cmd_prefix = (uint16_t)combine_length_codes(insert_length_code(ins), copy_length_code(len_coded(len)), code == 0);
    g.lit_idx = lit_idx - ((i == 0) ? carry : 0u);
    g.dist_idx = ndist;
    g.pos = pos - ((i == 0) ? carry : 0u);
    g.pad = u;  // owning unit: consumers add the unit's distance-symbol prefix to dist_idx
    if (g.cmd_prefix >= 128) ++ndist;
    out[nout++] = g;
    lit_idx += rc[i].insert_len;
    pos += rc[i].insert_len + len_bytes(rc[i].copy_len);
  }
  if (u + 1 == u1) {
    uint32_t carry_out = V.tail[u] + (nraw == 0 ? carry : 0u);
    if (carry_out) {
      const uint32_t uend = bmin(V.n, ustart + V.unit);
      GCmd g;
      g.insert_len = carry_out;
      g.copy_len = 0;
      g.dist_extra = 0;
      g.dist_prefix = 0;
      g.cmd_prefix = (uint16_t)combine_length_codes(insert_length_code(carry_out), copy_length_code(4), false);
      g.lit_idx = lit_idx + V.tail[u] - carry_out;
      g.dist_idx = ndist;
      g.pos = uend - carry_out;
      g.pad = u;
      out[nout++] = g;
    }
  }
  return ndist;
}

}  // namespace bro
