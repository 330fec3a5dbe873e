// Linearised MHL report on the PLANAR call planes (planar.cuh): replaces rcpp_mhl_report
// (/root/reference/src/rcpp_mhl_report.cpp:46-228) for haplotype contexts made of whole letter pairs (Hh / Xx / Zz - what
// generateMhlReport passes, R/generateMhlReport.R:178-180) on rows of at most 1024 positions.
//
// Same window model and finalize as window_mhl.cuh (difference arrays in shared memory, one CTA per 1024-position window);
// what changes is how a read is walked: a WARP takes a tile of 32 rows, ONE LANE PER ROW, 512 contiguous bytes per load
// (the byte-row kernel: a thread per read, 16 bytes per load, two byte-serial passes).  On bit planes
//   * the per-read counts (haplotype-context calls h, out-of-context methylated / unmethylated) are POPCs of LOP3 masks;
//   * coverage transitions are the set bits of nd ^ (nd << 1 | carry), nd = "not '-'" - a few per read;
//   * the calls of the haplotype context are the set bits of one mask, visited in position order to track methylated
//     stretches (:163-176); when a stretch closes, its span's units are re-read (L1/L2 hits) and the transitions of nd
//     inside the span push +-S(min(size, hmax)) into the numerator's difference array.
// The call codes are all the report needs, so it also runs on handles that hold the call planes only.
#pragma once
#include "planar.cuh"
#include "window_mhl.cuh"

namespace ea {

// ACC: int when rows-per-window x S(min(longest row, hmax)) < 2^31 (every sum of a window fits, the dispatcher checks):
// the difference arrays take native 32-bit shared atomics and 40 KB instead of 64 KB per CTA; else long long.
template <int CLS /* bit 0 Hh, bit 1 Xx, bit 2 Zz */, typename ACC>
__global__ void __launch_bounds__(kMhlThreads, (sizeof(ACC) == 4 ? 4 : 3))
k_window_mhl_pl(Planar pl, WinTable wt, MhlParams prm, MhlOut out) {
  constexpr int S = kMhlS;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ long long s_scan[4 * kMhlWarps];
  __shared__ uint32_t s_scan32[kMhlWarps + 1];
  __shared__ long long s_base;
  const int NCLS = prm.n_cls;
  ACC* const L64 = reinterpret_cast<ACC*>(smem_raw);                           // [2][3][S]
  int* const dWb = reinterpret_cast<int*>(L64 + 2 * 3 * S);                    // [2][S]
  uint32_t* const Tb = reinterpret_cast<uint32_t*>(dWb + 2 * S);               // [2][NCLS][S]
  const int kZero4 = (2 * 3 * S * (int)sizeof(ACC) + 2 * S * 4 + 2 * NCLS * S * 4) / 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr uint32_t sel2 = (CLS & 1) ? FULL : 0u, sel6 = (CLS & 2) ? FULL : 0u, sel7 = (CLS & 4) ? FULL : 0u;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < kZero4; i += kMhlThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];

    for (int t = (lo >> 5) + warp; t <= ((hi - 1) >> 5); t += kMhlWarps) {
      const long long u0 = pl.toff[t];
      const int nu_t = (int)(pl.toff[t + 1] - u0);
      const int r = t * 32 + lane;
      const bool mine = r >= lo && r < hi;
      const uint32_t lsv = mine ? pl.ls[r] : 0u;
      const int len = (int)(lsv & 0x7FFFFFFFu);
      const int s = (int)(lsv >> 31);
      const int rel0 = mine ? pl.start[r] - w0 : 0;
      const uint4* src = pl.units + u0 * 32 + lane;
      // ---- 1. whole-row counts ----
      uint32_t h = 0, om = 0, ou = 0;
#pragma unroll 4
      for (int j = 0; j < nu_t; j++) {
        const uint4 v = ldg_plunit(src + j * 32);
        const uint32_t in = ((~v.x & v.y & ~v.z) & sel2) | ((~v.x & v.y & v.z) & sel6) | ((v.x & v.y & v.z) & sel7);
        const uint32_t ooc = ((v.y & ~v.x) | (v.x & v.z)) & ~in;   // a call letter (low three bits 2, 6 | 5, 7) outside the context
        h += __popc(in);
        om += __popc(ooc & ~v.w);
        ou += __popc(ooc & v.w);
      }
      const bool keep = mine && len > 0 && rel0 + len > 0 && !((int)h < prm.hmin || (double)om / (double)(om + ou) > prm.max_ooctx);
      const long long hv = (long long)h;
      const long long dv = (long long)mhl_S(h < prm.cap ? h : prm.cap);
      int* const dW = dWb + s * S;
      ACC* const dH = L64 + (s * 3 + 0) * S;
      ACC* const dDe = L64 + (s * 3 + 1) * S;
      ACC* const dNu = L64 + (s * 3 + 2) * S;
      uint32_t* const T = Tb + s * NCLS * S;
      const int nu = keep ? (len + 31) >> 5 : 0;
      uint32_t covb = 0, msz = 0;
      int ms = 0, me = 0;
      auto close_stretch = [&]() {  // +-S(min(size, hmax)) around the non-'-' pieces of [ms, me] (both ends are calls)
        const long long nv = (long long)mhl_S(msz < prm.cap ? msz : prm.cap);
        uint32_t carry = 0;
        for (int ju = ms >> 5; ju <= (me >> 5); ju++) {
          const uint4 v = __ldg(src + ju * 32);
          const int b0 = ju == (ms >> 5) ? (ms & 31) : 0, b1 = ju == (me >> 5) ? (me & 31) : 31;
          const uint32_t span = (FULL << b0) & (FULL >> (31 - b1));
          const uint32_t sp = ~(v.x & v.y & ~v.z & v.w) & span;
          for (uint32_t tr = (sp ^ ((sp << 1) | carry)) & span; tr; tr &= tr - 1u) {
            const int k = __ffs(tr) - 1;
            const int q = rel0 + 32 * ju + k;
            if (q < S) acc_add(&dNu[max(q, 0)], ((sp >> k) & 1u) ? nv : -nv);
          }
          carry = sp >> 31;
        }
        if (rel0 + me + 1 < S) acc_add(&dNu[max(rel0 + me + 1, 0)], -nv);
        msz = 0;
      };
      // ---- 2. events, unit by unit (the second read of the row's units hits L2).  The unit loop is WARP-UNIFORM (the tile's
      //      unit rows) and every per-lane loop is followed by __syncwarp(): lanes that leave a data-dependent loop early do
      //      not reconverge on their own, and the rest of the sweep would run on a third of the lanes ----
      for (int j = 0; j < nu_t; j++) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < nu) v = ldg_plunit(src + j * 32);
        const int kmax = min(32, len - 32 * j);
        const uint32_t valid = j < nu ? (kmax >= 32 ? FULL : ((1u << kmax) - 1u)) : 0u;
        const uint32_t t2 = (~v.x & v.y & ~v.z) & sel2, t6 = (~v.x & v.y & v.z) & sel6, t7 = (v.x & v.y & v.z) & sel7;
        const uint32_t call = (t2 | t6 | t7) & valid;
        const uint32_t nd = ~(v.x & v.y & ~v.z & v.w) & valid;
        const int base = rel0 + 32 * j;
        for (uint32_t tr = (nd ^ ((nd << 1) | covb)) & valid; tr; tr &= tr - 1u) {  // coverage transitions
          const int k = __ffs(tr) - 1;
          const bool on = (nd >> k) & 1u;
          diff_event(base + k, S, dW, dH, dDe, on ? 1 : -1, on ? hv : -hv, on ? dv : -dv);
        }
        __syncwarp();
        if (j < nu) covb = (nd >> (kmax - 1)) & 1u;
        if (CLS & 1) { for (uint32_t m = t2 & valid; m; m &= m - 1u) { const int q = base + __ffs(m) - 1; if ((unsigned)q < (unsigned)S) atomicAdd(&T[prm.slot[0] * S + q], 1u); } __syncwarp(); }
        if (CLS & 2) { for (uint32_t m = t6 & valid; m; m &= m - 1u) { const int q = base + __ffs(m) - 1; if ((unsigned)q < (unsigned)S) atomicAdd(&T[prm.slot[1] * S + q], 1u); } __syncwarp(); }
        if (CLS & 4) { for (uint32_t m = t7 & valid; m; m &= m - 1u) { const int q = base + __ffs(m) - 1; if ((unsigned)q < (unsigned)S) atomicAdd(&T[prm.slot[2] * S + q], 1u); } __syncwarp(); }
        for (uint32_t m = call; m; m &= m - 1u) {   // methylated stretches over the context's calls in position order
          const int k = __ffs(m) - 1;
          if (!((v.w >> k) & 1u)) {
            if (!msz) ms = 32 * j + k;
            me = 32 * j + k; msz++;
          } else if (msz) close_stretch();
        }
        __syncwarp();
      }
      if (covb) diff_event(rel0 + len, S, dW, dH, dDe, -1, -hv, -dv);
      if (msz) close_stretch();
      __syncwarp();
    }
    __syncthreads();

    mhl_window_finalize<ACC>(smem_raw, prm, out, win, w0, s_scan, s_scan32, &s_base);
  }
}

}  // namespace ea
