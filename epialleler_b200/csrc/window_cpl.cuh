// Short-read window kernel, "chunk per lane" formulation (the default for reads up to 1 KB).
//
// One CTA per genomic window of S positions; each of its 8 warps takes tiles of up to 32 consecutive
// rows (one contiguous, 16-byte aligned byte range of the padded CSR).  A tile goes through three
// phases that alternate between the two natural decompositions of the work:
//
//   rows    one LANE PER ROW: offsets/start/strand are read coalesced, the row's coverage start/end
//           deltas are pushed, and a 32-bit descriptor (shared-memory index of the chunk's first
//           position, or 0 when the chunk lies outside the window) is written for every 16-byte
//           chunk of the row;
//   chunks  one LANE PER CHUNK: coalesced LDG.128 straight from HBM (several loads in flight per
//           lane, no staging buffer), nibble-plane classification, per-chunk class counts for the
//           thresholds, and everything that does not depend on the pass flag - {total calls} of the
//           requested contexts and '-' bytes - as shared-memory reductions.  Count word and the
//           methylated-call masks of the chunk replace its descriptor in shared memory;
//   rows    one LANE PER ROW again: the row's chunk counts are summed, the reference's pass/fail
//           chain is evaluated for 32 rows at once (src/rcpp_threshold_reads.cpp:39-70; through exact
//           integer tables built from the same IEEE double divisions, see k_thr_tables), flags are
//           stored, and the few passing rows push their methylated calls (SURVEY.md 8a invariant 4:
//           only the methylated count depends on `pass`).
//
// No lane ever walks a row serially over data bytes, every global load is coalesced, and the only
// warp-divergent code is the event loop over the handful of calls in a chunk.
// The finalize (coverage prefix sum, row test total_ctx > floor(cov/2), ordered compaction;
// src/rcpp_cx_report.cpp:58-85) touches shared memory only.
//
// Window state layout, events and finalize: win_common.cuh.
// Preconditions (checked by the dispatcher): padded rows <= 1024 bytes, rpt * max chunks per row <= tc,
// at most 32767 rows overlap a window.
#pragma once
#include "win_common.cuh"

namespace ea {

constexpr int kCplThreads = 256;
constexpr int kCplWarps = kCplThreads / 32;
constexpr int kCplBatch = 5;      // chunk loads in flight per lane

template <int TCTX, int WANT, bool FUSED, int S>
__global__ void __launch_bounds__(kCplThreads, 4)
k_window_cpl(Csr csr, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int rpt, int tc, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int SG = kGuardLo + S + kGuardHi;
  constexpr int MW = FUSED ? (NC == 3 ? 2 : 1) : 0;  // mask words per chunk
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[kCplWarps + 1];
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ __align__(4) uint8_t s_bp4[32];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  constexpr int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const aux_d = T + NC * 2 * SG + warp * (1 + MW) * tc;  // descriptor, then count word
  uint32_t* const aux_m = aux_d + tc;                              // methylated-call masks (slots 0,1)
  uint32_t* const aux_m2 = aux_m + tc;                             // slot 2 (CX reports)
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T), bp_tbl = smem_u32(s_bp4);
  const uint32_t w_floor = w_base + 4u * kGuardLo;

  if (threadIdx.x < 32) s_bp4[threadIdx.x] = (uint8_t)(4 * bit_pos(threadIdx.x));

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < kZero4; i += kCplThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_npass = 0;
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    int npass_w = 0;
    // the window's rows are split evenly over the warps, each warp's share into equal tiles of at most rpt rows
    const int per_warp = (hi - lo + kCplWarps - 1) / kCplWarps;
    const int my_lo = lo + warp * per_warp, my_hi = min(hi, my_lo + per_warp);
    const int n_t = (per_warp + rpt - 1) / rpt;
    const int per_tile = n_t ? (per_warp + n_t - 1) / n_t : 1;

    for (int base = my_lo; base < my_hi; base += per_tile) {
      // ------------------------------ rows: metadata, coverage deltas, chunk descriptors ------------------------------
      const int tend = min(base + per_tile, my_hi);
      const int r = base + lane;
      const bool valid = r < tend;
      const long long o0 = csr.off[valid ? r : tend];
      long long o1 = __shfl_down_sync(FULL, o0, 1);
      if (lane == 31) o1 = csr.off[tend];
      const long long tile_o0 = __shfl_sync(FULL, o0, 0);
      const int n_chunks = (int)((__shfl_sync(FULL, o1, 31) - tile_o0) >> 4);
      int start = 0, spv = 0;
      if (valid) { start = csr.start[r]; spv = csr.sp[r]; }
      const int nch = (int)((o1 - o0) >> 4);
      const int cbeg = (int)((o0 - tile_o0) >> 4);
      const int strand0 = spv & 1;
      const int len = nch * 16 - (spv >> 4);
      const int rel0 = start - w0;
      const bool overlaps = valid && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
      {
        // descriptor of chunk c: T-relative word index of its byte 0 (strand0*SG + rel + kGuardLo) when the chunk
        // overlaps the window, else 0; bit 31 = pass flag in the two-pass kernel
        int rel16 = rel0 + kGuardLo;
        uint32_t flag = 0u;
        if (overlaps) {
          const uint32_t sinc = strand0 ? 0x10000u : 1u;
          red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
          const int e = rel0 + len;
          if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
          if (!FUSED && (pass_in == nullptr || pass_in[r] != 0)) flag = 0x80000000u;
        } else {
          rel16 = -0x40000000;  // never in the window
        }
        uint32_t* dst = aux_d + cbeg;
        const uint32_t sbase = (uint32_t)(strand0 * SG) | flag;
        for (int c = 0; c < nch; c++) {
          dst[c] = ((unsigned)(rel16 - 1) < (unsigned)(S + 15)) ? sbase + (uint32_t)rel16 : 0u;
          rel16 += 16;
        }
      }
      __syncwarp();

      // ------------------------------ chunks: classification, counts, pass-independent events ------------------------------
      const uint8_t* ptr = csr.data + tile_o0 + (size_t)lane * 16;
      for (int j0 = lane; j0 < n_chunks; j0 += 32 * kCplBatch, ptr += 512 * kCplBatch) {
        uint4 v[kCplBatch];
#pragma unroll
        for (int u = 0; u < kCplBatch; u++)
          if (j0 + 32 * u < n_chunks) v[u] = ldg_chunk(ptr + 512 * u);
#pragma unroll
        for (int u = 0; u < kCplBatch; u++) {
          const int j = j0 + 32 * u;
          if (j >= n_chunks) break;
          const uint32_t d = aux_d[j];
          Planes p;
          make_planes(v[u], p);
          uint32_t m[3], meth[3];
          want_masks<WANT>(p, m, meth);
          const uint32_t ix = d & 0x7FFFFFFFu;  // strand*SG + rel16, 0 = not in the window
          if (ix) {
            const uint32_t s = ix >= (uint32_t)SG ? 1u : 0u;
            const uint32_t rel4 = 4u * (ix - s * (uint32_t)SG);
            const uint32_t methpass = FUSED ? 0u : ((d >> 31) ? (meth[0] | meth[1] | meth[2]) : 0u);
            chunk_events<NC, SG>(p, m, methpass, w_base + rel4, d_base + rel4, t_base + 4u * ix, s ? 0x10000u : 1u, w_floor, bp_tbl);
          }
          if (FUSED) {
            aux_d[j] = count4<TCTX>(p);
            aux_m[j] = ix ? (meth[0] | (meth[1] << 2)) : 0u;
            if (NC == 3) aux_m2[j] = ix ? meth[2] : 0u;
          }
        }
      }

      // ------------------------------ rows: class counts -> pass flag -> methylated calls of passing rows ------------------------------
      if (FUSED) {
        __syncwarp();
        bool pass = false;
        if (overlaps) {
          uint32_t acc_in = 0, acc_out = 0;
          const uint32_t* src = aux_d + cbeg;
          for (int c0 = 0; c0 < nch; c0 += 8) {  // byte fields hold up to 8 chunks (8 * 16 < 256)
            uint32_t part = 0;
            const int c1 = min(c0 + 8, nch);
            for (int c = c0; c < c1; c++) part += src[c];
            acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
            acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
          }
          pass = eval_pass_tables(acc_in >> 16, acc_in & 0xFFFFu, acc_out >> 16, acc_out & 0xFFFFu, thr, tbl);
          if (rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
        }
        // methylated calls of the passing rows, one row at a time with a lane per chunk
        const uint32_t ta = t_base + 4u * (uint32_t)(strand0 * SG + rel0 + kGuardLo);
        for (uint32_t pm = __ballot_sync(FULL, pass); pm; pm &= pm - 1u) {
          const int src_lane = __ffs(pm) - 1;
          const int cb = __shfl_sync(FULL, cbeg, src_lane), n = __shfl_sync(FULL, nch, src_lane);
          const uint32_t ta_s = __shfl_sync(FULL, ta, src_lane);
          for (int c = lane; c < n; c += 32) {
            const uint32_t mm = aux_m[cb + c];
            if (mm) {
              push_meth(mm & 0x33333333u, ta_s + 64u * c, bp_tbl);
              if (NC > 1) push_meth((mm >> 2) & 0x33333333u, ta_s + 64u * c + 2 * SG * 4, bp_tbl);
            }
            if (NC > 2) {
              const uint32_t m2 = aux_m2[cb + c];
              if (m2) push_meth(m2, ta_s + 64u * c + 4 * SG * 4, bp_tbl);
            }
          }
        }
      }
      __syncwarp();  // the next tile's descriptors overwrite aux
    }
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(FULL, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    __syncthreads();

    window_finalize<WANT, S, kCplThreads, FUSED>(W, D, T, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

}  // namespace ea
