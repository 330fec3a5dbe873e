// TEST INFRASTRUCTURE ONLY -- the counter-based synthetic read generator (epialleler_b200/csrc/synth.h)
// compiled for the HOST with g++, so that bench.py's reference arm and cpu_baseline leg can regenerate the workload's
// rows without loading the product library.  synth.h is plain C++ when __CUDACC__ is not defined; nothing of the CUDA
// library is linked here.
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "../epialleler_b200/csrc/synth.h"

extern "C" int oracle_synth_host(const ea_synth_cfg* cfg, int64_t row_begin, int64_t row_end, int32_t* rname, int32_t* strand,
                                 int32_t* start, int64_t* off, uint8_t* blob) {
  easynth::Plan plan;
  if (easynth::make_plan(cfg, &plan)) return -1;
  if (row_begin < 0 || row_end < row_begin || row_end > plan.n_reads || !off) return -1;
  const int64_t n = row_end - row_begin;
  off[0] = 0;
  for (int64_t i = 0; i < n; i++) {
    const easynth::RowInfo r = easynth::row_info(plan, row_begin + i);
    if (rname) rname[i] = r.rname;
    if (strand) strand[i] = r.strand;
    if (start) start[i] = r.start;
    off[i + 1] = off[i] + r.len;
  }
  if (!blob) return 0;
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::thread::hardware_concurrency(), n / 4096 + 1));
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      for (int64_t i = n * t / nt; i < n * (t + 1) / nt; i++) {
        const easynth::RowInfo r = easynth::row_info(plan, row_begin + i);
        uint8_t* d = blob + off[i];
        for (int j = 0; j < r.len; j++) d[j] = easynth::read_byte(plan, row_begin + i, r, j);
      }
    });
  for (auto& x : th) x.join();
  return 0;
}
