// Host half of the "calls only" packer, see hostpack.cpp.  Internal to the library.
#pragma once
#include <cstdint>

namespace eahp {

struct Job {
  const uint8_t* blob;      // the strings, one after the other (file order == row order)
  const int64_t* off;       // n + 1 offsets
  const int32_t* strand;    // 1 '+', 2 '-'
  const int32_t* templid;   // nullptr or the identity
  int64_t n;
  // scattered form (ea_pack_calls: the reference's std::vector<std::string>): row r = string t = templid[r] (or r).  Rows
  // need NOT be the strings in file order here (paired-end data: templates are read in name order and the table is sorted
  // by coordinate afterwards) - every row is fetched through its pointer anyway.  Two ways to say where string t lies:
  //   tables  : str[t] / len[t]
  //   records : an array of fixed-size records (the std::string objects themselves), string t's data pointer at
  //             rec + t * stride + ptr_off and its length (size_t) at rec + t * stride + len_off - no tables to build
  const uint8_t* const* str = nullptr;
  const int64_t* len = nullptr;
  const uint8_t* rec = nullptr;
  int64_t stride = 0, ptr_off = 0, len_off = 0;
  int64_t nstr = 0;
  bool scattered() const { return str != nullptr || rec != nullptr; }
  int64_t sidx(int64_t r) const { return templid ? (int64_t)templid[r] : r; }
  const uint8_t* ptr_at(int64_t t) const { return rec ? *reinterpret_cast<const uint8_t* const*>(rec + t * stride + ptr_off) : str[t]; }
  int64_t len_at(int64_t t) const { return rec ? (int64_t)*reinterpret_cast<const uint64_t*>(rec + t * stride + len_off) : len[t]; }
  const uint8_t* row_ptr(int64_t r) const { return scattered() ? ptr_at(sidx(r)) : blob + off[r]; }
  int64_t row_len(int64_t r) const { return scattered() ? len_at(sidx(r)) : off[r + 1] - off[r]; }
};

enum : unsigned { kNotIdentity = 1u, kBadLength = 2u, kBadStrand = 4u, kBadCode = 8u, kBadTemplid = 16u, kNullString = 32u };

// pass 1 over the rows of tiles [t0, t1) (tiles of 32 rows): ls[r] = length | (strand == '-') << 31, tile_nu[t] = unit rows
// of tile t (max over its rows of ceil(length / 32)); returns the OR of the flags above
unsigned scan_rows(const Job& job, int64_t t0, int64_t t1, uint32_t* ls, int32_t* tile_nu);
// planar units of tiles [t0, t1) in the tile-interleaved layout of planar.cuh; dst = where tile t0 starts
unsigned pack_tiles(const Job& job, const int64_t* toff, int64_t t0, int64_t t1, uint8_t* dst);
// long-read layout (units of a row contiguous): rows [r0, r1), uoff = per-row unit offsets; dst = where row r0 starts
unsigned pack_rows(const Job& job, const int64_t* uoff, int64_t r0, int64_t r1, uint8_t* dst);

}  // namespace eahp
