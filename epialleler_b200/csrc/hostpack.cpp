// Host half of the "calls only" packer (ea_pack_blob_calls): turns the caller's byte-per-position strings into the planar
// call planes of planar.cuh ON THE HOST, with all cores, so that 4 bits per position cross PCIe instead of 8.  The base
// nibble (high nibble) is dropped - the cytosine-report path never reads it (`& 15`, src/rcpp_cx_report.cpp:117,
// src/rcpp_threshold_reads.cpp:34).  One AVX2 instruction pair per bit plane: shift the wanted bit of every byte into
// the byte's top bit, VPMOVMSKB gathers the 32 top bits = one plane word of a unit.
// Plain C++ (g++), no CUDA: the library's C ABI functions in lib.cu drive it.
#include "hostpack.h"

#include <immintrin.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace eahp {

namespace {

inline bool valid_code(unsigned c) { return (0xFCE4u >> (c & 15u)) & 1u; }  // {2,5,6,7,10..15}

// scalar: 32 bytes (n valid) -> one unit; returns false if a valid position holds a code outside the domain
bool unit_scalar(const uint8_t* p, int n, uint32_t out[4]) {
  uint32_t a = 0, b = 0, c = 0, d = 0;
  bool ok = true;
  for (int k = 0; k < n; k++) {
    const unsigned v = p[k] & 15u;
    ok &= valid_code(v);
    a |= (v & 1u) << k; b |= ((v >> 1) & 1u) << k; c |= ((v >> 2) & 1u) << k; d |= ((v >> 3) & 1u) << k;
  }
  out[0] = a; out[1] = b; out[2] = c; out[3] = d;
  return ok;
}

__attribute__((target("avx2"))) bool unit_avx2(const uint8_t* p, int n, uint32_t out[4]) {
  __m256i v;
  if (n >= 32) {
    v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
  } else {  // the row's last unit: never read past the row
    alignas(32) uint8_t tmp[32] = {0};
    memcpy(tmp, p, (size_t)n);
    v = _mm256_load_si256(reinterpret_cast<const __m256i*>(tmp));
  }
  const uint32_t keep = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
  out[0] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 7)) & keep;
  out[1] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6)) & keep;
  out[2] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5)) & keep;
  out[3] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 4)) & keep;
  // domain check: table lookup on the low nibbles, 0x80 for the codes the reference's readers emit
  const __m256i tbl = _mm256_setr_epi8(0, 0, (char)0x80, 0, 0, (char)0x80, (char)0x80, (char)0x80, 0, 0, (char)0x80, (char)0x80, (char)0x80, (char)0x80,
                                       (char)0x80, (char)0x80, 0, 0, (char)0x80, 0, 0, (char)0x80, (char)0x80, (char)0x80, 0, 0, (char)0x80, (char)0x80,
                                       (char)0x80, (char)0x80, (char)0x80, (char)0x80);
  const uint32_t okm = (uint32_t)_mm256_movemask_epi8(_mm256_shuffle_epi8(tbl, _mm256_and_si256(v, _mm256_set1_epi8(0x0F))));
  return (okm & keep) == keep;
}

bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

}  // namespace

unsigned scan_rows(const Job& job, int64_t t0, int64_t t1, uint32_t* ls, int32_t* tile_nu) {
  unsigned flags = 0;
  for (int64_t t = t0; t < t1; t++) {
    int mx = 0;
    const int64_t r1 = (t * 32 + 32 < job.n) ? t * 32 + 32 : job.n;
    for (int64_t r = t * 32; r < r1; r++) {
      if (job.scattered()) {   // scattered strings: any in-range templid is fine
        const int64_t t = job.sidx(r);
        if (t < 0 || t >= job.nstr) { flags |= kBadTemplid; ls[r] = 0; continue; }
        if (job.len_at(t) > 0 && !job.ptr_at(t)) { flags |= kNullString; ls[r] = 0; continue; }
      } else if (job.templid && job.templid[r] != (int32_t)r) {
        flags |= kNotIdentity;
      }
      const int64_t l = job.row_len(r);
      if (l < 0 || l > 0x7FFFFFF0ll) { flags |= kBadLength; ls[r] = 0; continue; }
      const int32_t s = job.strand[r];
      if (s != 1 && s != 2) flags |= kBadStrand;
      ls[r] = (uint32_t)l | (s == 2 ? 0x80000000u : 0u);
      const int nu = (int)((l + 31) >> 5);
      if (nu > mx) mx = nu;
    }
    tile_nu[t] = mx;
  }
  return flags;
}

// unit without the domain check (the device validates the planes it receives): full 32-byte load, garbage beyond n masked
__attribute__((target("avx2"))) static inline void unit_avx2_fast(const uint8_t* p, int n, __m128i* dst) {
  const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
  const uint32_t keep = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
  const uint32_t a = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 7)) & keep;
  const uint32_t b = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6)) & keep;
  const uint32_t c = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5)) & keep;
  const uint32_t d = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 4)) & keep;
  _mm_stream_si128(dst, _mm_set_epi32((int)d, (int)c, (int)b, (int)a));   // written once, read by the DMA engine only
}

// the row's LAST, partial unit (n < 32 valid positions) of a row of at least 32 bytes: load the 32 bytes that END at the
// row's end - always inside the string - and shift the plane words down, instead of copying the tail into a scratch block
__attribute__((target("avx2"))) static inline void unit_avx2_tail(const uint8_t* row_end, int n, __m128i* dst) {
  const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(row_end - 32));
  const int sh = 32 - n;
  const uint32_t a = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 7)) >> sh;
  const uint32_t b = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6)) >> sh;
  const uint32_t c = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5)) >> sh;
  const uint32_t d = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 4)) >> sh;
  _mm_stream_si128(dst, _mm_set_epi32((int)d, (int)c, (int)b, (int)a));
}

__attribute__((target("avx2"))) static void pack_tiles_avx2(const Job& job, const int64_t* toff, int64_t t0, int64_t t1, uint8_t* dst) {
  // a full 32-byte load at or below safe_end stays inside the blob; scattered strings: only full units are loaded directly
  const uint8_t* const safe_end = job.scattered() ? nullptr : job.blob + job.off[job.n] - 32;
  static const uint8_t kNone[1] = {0};
  for (int64_t t = t0; t < t1; t++) {
    __m128i* tile = reinterpret_cast<__m128i*>(dst + (toff[t] - toff[t0]) * 512);
    const int nu_t = (int)(toff[t + 1] - toff[t]);
    const uint8_t* row[32];
    int len[32];
    for (int l = 0; l < 32; l++) {
      const int64_t r = t * 32 + l;
      if (r < job.n) { row[l] = job.row_ptr(r); len[l] = (int)job.row_len(r); }
      else { row[l] = kNone; len[l] = 0; }
    }
    for (int j = 0; j < nu_t; j++) {   // unit-major: the tile's 512-byte unit rows are written front to back
      for (int l = 0; l < 32; l++) {
        const int n = len[l] - 32 * j;
        const uint8_t* p = row[l] + 32 * j;
        if (n <= 0) { _mm_stream_si128(tile + j * 32 + l, _mm_setzero_si128()); continue; }
        if (safe_end ? p <= safe_end : n >= 32) { unit_avx2_fast(p, n, tile + j * 32 + l); continue; }
        if (len[l] >= 32) { unit_avx2_tail(row[l] + len[l], n, tile + j * 32 + l); continue; }
        alignas(32) uint8_t tmp[32] = {0};
        memcpy(tmp, p, (size_t)(n < 32 ? n : 32));
        unit_avx2_fast(tmp, n, tile + j * 32 + l);
      }
    }
  }
  _mm_sfence();
}

// AVX-512 + VBMI + GFNI: 64 positions (two units) per step.  A masked load takes the bytes (nothing beyond the row is
// touched, masked-out bytes read as code 0), GF2P8AFFINE with the data as the MATRIX operand transposes every 8 x 8 bit block
// (byte m of a qword = bit m of its eight bytes, after a byte reversal inside the qwords that puts position k at bit k), and
// one VPERMB gathers the four low planes of both units: bytes 0-15 = unit 0 {A,B,C,D}, bytes 16-31 = unit 1.
#define EA_HP_512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi,gfni")))
EA_HP_512 static inline __m256i planes64(const uint8_t* p, int n /* valid bytes, 1..64 */) {
  alignas(64) static const uint8_t kRev[64] = {7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8,
                                               7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8};
  alignas(64) static const uint8_t kIdx[64] = {0, 8, 16, 24, 1, 9, 17, 25, 2, 10, 18, 26, 3, 11, 19, 27, 32, 40, 48, 56, 33, 41, 49, 57, 34, 42, 50, 58, 35, 43, 51, 59};
  const __mmask64 k = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
  __m512i v = _mm512_maskz_loadu_epi8(k, p);
  v = _mm512_shuffle_epi8(v, _mm512_load_si512(kRev));
  const __m512i t = _mm512_gf2p8affine_epi64_epi8(_mm512_set1_epi64((long long)0x8040201008040201ull), v, 0);
  return _mm512_castsi512_si256(_mm512_permutexvar_epi8(_mm512_load_si512(kIdx), t));
}

EA_HP_512 static void pack_tiles_avx512(const Job& job, const int64_t* toff, int64_t t0, int64_t t1, uint8_t* dst) {
  static const uint8_t kNone[1] = {0};
  for (int64_t t = t0; t < t1; t++) {
    __m128i* tile = reinterpret_cast<__m128i*>(dst + (toff[t] - toff[t0]) * 512);
    const int nu_t = (int)(toff[t + 1] - toff[t]);
    const uint8_t* row[32];
    int len[32];
    for (int l = 0; l < 32; l++) {
      const int64_t r = t * 32 + l;
      if (r < job.n) { row[l] = job.row_ptr(r); len[l] = (int)job.row_len(r); }
      else { row[l] = kNone; len[l] = 0; }
    }
    for (int j = 0; j < nu_t; j += 2) {   // two unit rows of the tile per step (1 KB written front to back)
      const bool two = j + 1 < nu_t;
      for (int l = 0; l < 32; l++) {
        const int n = len[l] - 32 * j;
        __m256i w = _mm256_setzero_si256();
        if (n > 0) w = planes64(row[l] + 32 * j, n < 64 ? n : 64);
        _mm_stream_si128(tile + j * 32 + l, _mm256_castsi256_si128(w));
        if (two) _mm_stream_si128(tile + (j + 1) * 32 + l, _mm256_extracti128_si256(w, 1));
      }
    }
  }
  _mm_sfence();
}

EA_HP_512 static void pack_rows_avx512(const Job& job, const int64_t* uoff, int64_t r0, int64_t r1, uint8_t* dst) {
  for (int64_t r = r0; r < r1; r++) {
    const int64_t len = job.row_len(r);
    const uint8_t* row = job.row_ptr(r);
    uint8_t* d = dst + (uoff[r] - uoff[r0]) * 16;
    const int64_t nu = (len + 31) >> 5;
    for (int64_t j = 0; j < nu; j += 2) {
      const int64_t n = len - 32 * j;
      const __m256i w = planes64(row + 32 * j, (int)(n < 64 ? n : 64));
      if (j + 1 < nu) _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + j * 16), w);
      else _mm_storeu_si128(reinterpret_cast<__m128i*>(d + j * 16), _mm256_castsi256_si128(w));
    }
  }
}

static bool have_avx512() {
  static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                        __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("gfni") && !getenv("EA_NO_AVX512");
  return v;
}

unsigned pack_tiles(const Job& job, const int64_t* toff, int64_t t0, int64_t t1, uint8_t* dst) {
  if (have_avx512() && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) { pack_tiles_avx512(job, toff, t0, t1, dst); return 0u; }
  if (have_avx2() && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) { pack_tiles_avx2(job, toff, t0, t1, dst); return 0u; }
  bool ok = true;
  for (int64_t t = t0; t < t1; t++) {
    uint8_t* tile = dst + (toff[t] - toff[t0]) * 512;
    const int nu_t = (int)(toff[t + 1] - toff[t]);
    for (int l = 0; l < 32; l++) {
      const int64_t r = t * 32 + l;
      int64_t len = 0;
      const uint8_t* row = nullptr;
      if (r < job.n) { len = job.row_len(r); row = job.row_ptr(r); }
      const int nu = (int)((len + 31) >> 5);
      for (int j = 0; j < nu_t; j++) {
        uint32_t u[4] = {0, 0, 0, 0};
        if (j < nu) {
          const int n = (int)(len - 32 * (int64_t)j < 32 ? len - 32 * (int64_t)j : 32);
          ok &= unit_scalar(row + 32 * j, n, u);
        }
        memcpy(tile + ((size_t)j * 32 + l) * 16, u, 16);
      }
    }
  }
  return ok ? 0u : kBadCode;
}

unsigned pack_rows(const Job& job, const int64_t* uoff, int64_t r0, int64_t r1, uint8_t* dst) {
  if (have_avx512()) { pack_rows_avx512(job, uoff, r0, r1, dst); return 0u; }   // the device validates the planes it receives
  const bool avx2 = have_avx2();
  bool ok = true;
  for (int64_t r = r0; r < r1; r++) {
    const int64_t len = job.row_len(r);
    const uint8_t* row = job.row_ptr(r);
    uint8_t* d = dst + (uoff[r] - uoff[r0]) * 16;
    const int64_t nu = (len + 31) >> 5;
    for (int64_t j = 0; j < nu; j++) {
      uint32_t u[4];
      const int n = (int)(len - 32 * j < 32 ? len - 32 * j : 32);
      ok &= avx2 ? unit_avx2(row + 32 * j, n, u) : unit_scalar(row + 32 * j, n, u);
      memcpy(d + j * 16, u, 16);
    }
  }
  return ok ? 0u : kBadCode;
}

}  // namespace eahp
