// CPU check of the host half of the packers (csrc/hostpack.cpp): the scattered-string form (ea_pack_calls) must produce the
// very planes the contiguous-blob form (ea_pack_blob_calls) produces, for short rows, partial last units and ragged tiles.
// Built and run by tests/test_hostpack.py.
#include "hostpack.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
  eahp::Job js{nullptr, nullptr, strand.data(), nullptr, n}; js.str = sp.data(); js.len = sl.data();
  const int64_t nt = (n + 31) / 32;
  std::vector<uint32_t> ls(n), ls2(n); std::vector<int32_t> nu(nt), nu2(nt);
  eahp::scan_rows(jb, 0, nt, ls.data(), nu.data()); eahp::scan_rows(js, 0, nt, ls2.data(), nu2.data());
  if (ls != ls2 || nu != nu2) { printf("scan mismatch\n"); return 1; }
  std::vector<int64_t> toff(nt + 1, 0); for (int64_t t = 0; t < nt; t++) toff[t + 1] = toff[t] + nu[t];
  void *a, *b; posix_memalign(&a, 64, toff[nt] * 512 + 64); posix_memalign(&b, 64, toff[nt] * 512 + 64);
  memset(a, 0xAA, toff[nt] * 512); memset(b, 0x55, toff[nt] * 512);
  eahp::pack_tiles(jb, toff.data(), 0, nt, (uint8_t*)a); eahp::pack_tiles(js, toff.data(), 0, nt, (uint8_t*)b);
  printf("%s\n", memcmp(a, b, toff[nt] * 512) == 0 ? "scattered == blob: OK" : "MISMATCH");
  return memcmp(a, b, toff[nt] * 512) != 0;
}
