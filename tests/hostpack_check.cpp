// CPU check of the host half of the packers (csrc/hostpack.cpp): the scattered-string form (ea_pack_calls) must produce the
// very planes the contiguous-blob form (ea_pack_blob_calls) produces, for short rows, partial last units and ragged tiles.
// Built and run by tests/test_hostpack.py.
#include "hostpack.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
int main() {
  srand(5);
  const int64_t n = 5000;
  std::vector<std::string> strs; std::vector<int64_t> off{0}; std::vector<uint8_t> blob;
  const uint8_t codes[] = {2,5,6,7,10,11,12,13,14,15};
  for (int64_t i = 0; i < n; i++) { int len = i % 7 == 0 ? rand() % 40 : 100 + rand() % 300; std::string s; for (int k = 0; k < len; k++) s.push_back((char)(((rand()%4+1) << 4) | codes[rand() % 10])); strs.push_back(s); blob.insert(blob.end(), s.begin(), s.end()); off.push_back(off.back() + len); }
  blob.resize(blob.size() + 64);
  std::vector<int32_t> strand(n, 1);
  std::vector<const uint8_t*> sp(n); std::vector<int64_t> sl(n);
  for (int64_t i = 0; i < n; i++) { sp[i] = (const uint8_t*)strs[i].data(); sl[i] = strs[i].size(); }
  eahp::Job jb{blob.data(), off.data(), strand.data(), nullptr, n};
  eahp::Job js{nullptr, nullptr, strand.data(), nullptr, n}; js.str = sp.data(); js.len = sl.data(); js.nstr = n;
  const int64_t nt = (n + 31) / 32;
  std::vector<uint32_t> ls(n), ls2(n); std::vector<int32_t> nu(nt), nu2(nt);
  eahp::scan_rows(jb, 0, nt, ls.data(), nu.data()); eahp::scan_rows(js, 0, nt, ls2.data(), nu2.data());
  if (ls != ls2 || nu != nu2) { printf("scan mismatch\n"); return 1; }
  std::vector<int64_t> toff(nt + 1, 0); for (int64_t t = 0; t < nt; t++) toff[t + 1] = toff[t] + nu[t];
  void *a, *b; posix_memalign(&a, 64, toff[nt] * 512 + 64); posix_memalign(&b, 64, toff[nt] * 512 + 64);
  memset(a, 0xAA, toff[nt] * 512); memset(b, 0x55, toff[nt] * 512);
  eahp::pack_tiles(jb, toff.data(), 0, nt, (uint8_t*)a); eahp::pack_tiles(js, toff.data(), 0, nt, (uint8_t*)b);
  if (memcmp(a, b, toff[nt] * 512) != 0) { printf("MISMATCH\n"); return 1; }
  // whatever SIMD path the library picked (AVX-512 + GFNI, AVX2, scalar) against a plain restatement of the layout:
  // unit j of row l of tile t = 16 bytes at ((toff[t] + j) * 32 + l) * 16, plane b bit k = bit b of the code at position 32 j + k
  {
    const uint32_t* units = (const uint32_t*)a;
    for (int64_t r = 0; r < n; r++) {
      const int64_t t = r / 32, l = r % 32;
      for (int64_t j = 0; j < toff[t + 1] - toff[t]; j++) {
        uint32_t want[4] = {0, 0, 0, 0};
        for (int k = 0; k < 32; k++) { const int64_t pos = 32 * j + k; if (pos < (int64_t)strs[r].size()) for (int pb = 0; pb < 4; pb++) want[pb] |= (uint32_t)(((uint8_t)strs[r][pos] >> pb) & 1) << k; }
        if (memcmp(units + ((toff[t] + j) * 32 + l) * 4, want, 16) != 0) { printf("plane mismatch at row %lld unit %lld\n", (long long)r, (long long)j); return 1; }
      }
    }
    // the long-read layout (units of a row contiguous)
    std::vector<int64_t> uoff(n + 1, 0); for (int64_t r = 0; r < n; r++) uoff[r + 1] = uoff[r] + ((int64_t)strs[r].size() + 31) / 32;
    std::vector<uint32_t> rows_out((size_t)uoff[n] * 4 + 16, 0xDEADBEEFu);
    eahp::pack_rows(js, uoff.data(), 0, n, (uint8_t*)rows_out.data());
    for (int64_t r = 0; r < n; r++)
      for (int64_t j = 0; j < uoff[r + 1] - uoff[r]; j++) {
        uint32_t want[4] = {0, 0, 0, 0};
        for (int k = 0; k < 32; k++) { const int64_t pos = 32 * j + k; if (pos < (int64_t)strs[r].size()) for (int pb = 0; pb < 4; pb++) want[pb] |= (uint32_t)(((uint8_t)strs[r][pos] >> pb) & 1) << k; }
        if (memcmp(rows_out.data() + (uoff[r] + j) * 4, want, 16) != 0) { printf("row-layout mismatch at row %lld unit %lld\n", (long long)r, (long long)j); return 1; }
      }
    if (rows_out[(size_t)uoff[n] * 4] != 0xDEADBEEFu) { printf("pack_rows wrote past its end\n"); return 1; }
  }
  // rows that reach their strings through templid (paired-end data): strings stored in a shuffled order
  std::vector<int32_t> perm(n); for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
  for (int64_t i = n - 1; i > 0; i--) std::swap(perm[i], perm[rand() % (i + 1)]);
  std::vector<const uint8_t*> sp2(n); std::vector<int64_t> sl2(n);
  for (int64_t r = 0; r < n; r++) { sp2[perm[r]] = sp[r]; sl2[perm[r]] = sl[r]; }   // string perm[r] holds row r's bytes
  eahp::Job jt{nullptr, nullptr, strand.data(), perm.data(), n}; jt.str = sp2.data(); jt.len = sl2.data(); jt.nstr = n;
  std::vector<uint32_t> ls3(n); std::vector<int32_t> nu3(nt);
  const unsigned f3 = eahp::scan_rows(jt, 0, nt, ls3.data(), nu3.data());
  if (f3 || ls3 != ls || nu3 != nu) { printf("templid scan mismatch (flags %u)\n", f3); return 1; }
  memset(b, 0x33, toff[nt] * 512);
  eahp::pack_tiles(jt, toff.data(), 0, nt, (uint8_t*)b);
  if (memcmp(a, b, toff[nt] * 512) != 0) { printf("templid MISMATCH\n"); return 1; }
  // record-array form: the std::string objects themselves (libstdc++: data pointer at 0, length at 8)
  if (sizeof(std::string) == 32 && *reinterpret_cast<const char* const*>(&strs[5]) == strs[5].data()) {
    eahp::Job jr{nullptr, nullptr, strand.data(), nullptr, n};
    jr.rec = reinterpret_cast<const uint8_t*>(strs.data()); jr.stride = sizeof(std::string); jr.ptr_off = 0; jr.len_off = 8; jr.nstr = n;
    std::vector<uint32_t> ls4(n); std::vector<int32_t> nu4(nt);
    if (eahp::scan_rows(jr, 0, nt, ls4.data(), nu4.data()) || ls4 != ls || nu4 != nu) { printf("records scan mismatch\n"); return 1; }
    memset(b, 0x77, toff[nt] * 512);
    eahp::pack_tiles(jr, toff.data(), 0, nt, (uint8_t*)b);
    if (memcmp(a, b, toff[nt] * 512) != 0) { printf("records MISMATCH\n"); return 1; }
  }
  perm[7] = (int32_t)n;   // out of range
  if (!(eahp::scan_rows(jt, 0, 1, ls3.data(), nu3.data()) & eahp::kBadTemplid)) { printf("bad templid not flagged\n"); return 1; }
  printf("scattered == blob, templid followed: OK\n");
  return 0;
}
