// Stand-in for the R process around integration/r_shim/rcpp_b200_shim.cpp (R and Rcpp are absent from the build image):
// it holds the packed templates the way the reference's reader leaves them - a heap std::vector<std::string>, one
// separately allocated, PAGEABLE string per template (src/rcpp_read_bam.cpp:61-69,185: seqxm->emplace_back) - and makes
// exactly the calls handle_of() in the shim makes: build the pointer/length tables, then ea_pack_calls (or ea_pack).
// bench.py times the end-to-end path through it; tests/ check that the handle equals one packed from a blob.
// Plain C++ (g++), linked against the product's C ABI only.
#include <cstdint>
#include <algorithm>
#include <initializer_list>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/epialleler_b200.h"

extern "C" {

// what rcpp_read_bam leaves behind: one emplace_back per template
void* harness_strings_new(const uint8_t* blob, const int64_t* off, int64_t n) {
  auto* v = new std::vector<std::string>();
  v->reserve((size_t)n);
  for (int64_t i = 0; i < n; i++) v->emplace_back(reinterpret_cast<const char*>(blob + off[i]), (size_t)(off[i + 1] - off[i]));
  return v;
}
void harness_strings_free(void* p) { delete static_cast<std::vector<std::string>*>(p); }
int64_t harness_strings_count(void* p) { return (int64_t)static_cast<std::vector<std::string>*>(p)->size(); }

// Where a std::string keeps its data pointer and its length, found with probe strings (libstdc++: 0 and 8 in a 32-byte
// object; the probe makes no assumption).  false: layout not understood - use the pointer tables.
static bool string_layout(int64_t* ptr_off, int64_t* len_off) {
  static_assert(sizeof(std::string) % 8 == 0, "unexpected std::string size");
  int64_t po = -1, lo = -1;
  for (size_t n : {(size_t)100, (size_t)3, (size_t)7777}) {   // heap, short-string and large probes must agree
    std::string probe(n, 'x');
    const unsigned char* raw = reinterpret_cast<const unsigned char*>(&probe);
    int64_t p = -1, l = -1;
    for (size_t o = 0; o + 8 <= sizeof(std::string); o += 8) {
      if (p < 0 && *reinterpret_cast<const char* const*>(raw + o) == probe.data()) p = (int64_t)o;
      else if (l < 0 && *reinterpret_cast<const size_t*>(raw + o) == n) l = (int64_t)o;
    }
    if (p < 0 || l < 0 || (po >= 0 && (p != po || l != lo))) return false;
    po = p; lo = l;
  }
  *ptr_off = po; *len_off = lo;
  return true;
}

// handle_of() of the shim, table-free form: the vector's std::string objects ARE the record array
int harness_pack_records(void* p, const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid, int64_t nrow,
                         int device, ea_handle** out) {
  auto* seqxm = static_cast<std::vector<std::string>*>(p);
  int64_t po, lo;
  if (!string_layout(&po, &lo)) return -1000;
  return ea_pack_calls_records(rname, strand, start, templid, nrow, seqxm->data(), (int64_t)seqxm->size(), (int64_t)sizeof(std::string), po, lo, device, out);
}

// the shim's multi-GPU pack: the same record array, cut into range shards, one device each
int harness_pack_multi_records(void* p, const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid, int64_t nrow,
                               const int* devices, int ndev, ea_multi** out) {
  auto* seqxm = static_cast<std::vector<std::string>*>(p);
  int64_t po, lo;
  if (!string_layout(&po, &lo)) return -1000;
  return ea_pack_multi_records(rname, strand, start, templid, nrow, seqxm->data(), (int64_t)seqxm->size(), (int64_t)sizeof(std::string), po, lo, devices, ndev, out);
}

// handle_of() of the shim: pointer and length tables over the vector, then the packer
int harness_pack(void* p, const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid, int64_t nrow,
                 int device, int calls_only, ea_handle** out) {
  auto* seqxm = static_cast<std::vector<std::string>*>(p);
  const size_t n = seqxm->size();
  // plain arrays filled by a few threads (no R API in them): a std::vector would zero-fill 2 x 400 MB on one thread first
  std::unique_ptr<const uint8_t*[]> str(new const uint8_t*[n ? n : 1]);
  std::unique_ptr<int64_t[]> len(new int64_t[n ? n : 1]);
  const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      for (size_t i = n * t / nt; i < n * (t + 1) / nt; i++) {
        str[i] = reinterpret_cast<const uint8_t*>((*seqxm)[i].data());
        len[i] = static_cast<int64_t>((*seqxm)[i].size());
      }
    });
  for (auto& x : th) x.join();
  return (calls_only ? ea_pack_calls : ea_pack)(rname, strand, start, templid, nrow, str.get(), len.get(), (int64_t)n, device, out);
}

}  // extern "C"
