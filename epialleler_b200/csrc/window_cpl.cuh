// Short-read window kernel, "chunk per lane" formulation (the default for reads up to 1 KB).
//
// One CTA per genomic window of S positions; each of its 8 warps takes tiles of up to 32 consecutive
// rows (one contiguous, 16-byte aligned byte range of the padded CSR).  A tile goes through three
// phases that alternate between the two natural decompositions of the work:
//
//   rows    one LANE PER ROW: offsets/start/strand are read coalesced, the row's coverage start/end
//           deltas are pushed, and a 32-bit descriptor (window-relative position, strand) is written
//           to shared memory for every 16-byte chunk of the row;
//   chunks  one LANE PER CHUNK: coalesced LDG.128 straight from HBM (several loads in flight per
//           lane, no staging buffer), nibble-plane classification, per-chunk class counts for the
//           thresholds, and everything that does not depend on the pass flag - {total calls} of the
//           requested contexts and '-' bytes - as shared-memory atomics.  Count word and the
//           methylated-call masks of the chunk replace its descriptor in shared memory;
//   rows    one LANE PER ROW again: the row's chunk counts are summed, the reference's pass/fail
//           chain is evaluated for 32 rows at once (src/rcpp_threshold_reads.cpp:39-70), flags are
//           stored, and the few passing rows push their methylated calls (SURVEY.md 8a invariant 4:
//           only the methylated count depends on `pass`).
//
// No lane ever walks a row serially over data bytes, every global load is coalesced, and the only
// warp-divergent code is the event loop over the handful of calls in a chunk.
// The finalize (coverage prefix sum, row test total_ctx > floor(cov/2), ordered compaction;
// src/rcpp_cx_report.cpp:58-85) touches shared memory only.
//
// Shared-memory window state, per strand, SG = 16 + S + 32 words each (guard bands absorb the
// out-of-window bytes of chunks that straddle a window edge, so events need no bounds checks):
//   W0     low 16 bits: running-coverage DELTA (+1 row start, -1 row end, -1/+1 around long '-'
//          runs; two's complement), high 16 bits: number of isolated '-' bytes at the position;
//          cov(p) = prefix_sum(low)(p) - high(p)
//   T[c]   per requested context: total calls | methylated calls of passing reads << 16
// Preconditions (checked by the dispatcher): padded rows <= 1024 bytes, rpt * max chunks per row <= tc,
// at most 32767 rows overlap a window.
#pragma once
#include "kernels.cuh"
#include "window_tpr.cuh"  // block_exscan_n, block_exscan64_n, bit_pos

namespace ea {

constexpr int kCplThreads = 256;
constexpr int kCplWarps = kCplThreads / 32;
constexpr int kGuardLo = 16;
constexpr int kGuardHi = 32;
constexpr uint32_t kDescBias = 2048u;
constexpr uint32_t kDescSkip = 0xFFFFFFFFu;
constexpr int kCplBatch = 5;  // chunk loads in flight per lane

// class counts of one chunk, one byte each: in-context calls | methylated << 8 | out-of-context << 16 | methylated << 24
template <int TCTX>
__device__ __forceinline__ uint32_t count4(const Planes& p) {
  const uint32_t t0 = cls_in<TCTX>(p.A0, p.B0, p.C0), t1 = cls_in<TCTX>(p.A1, p.B1, p.C1);
  uint32_t c = __popc(t0 | t1) + (__popc((t0 & ~p.D0) | (t1 & ~p.D1)) << 8);
  if (TCTX != TC_CX) {
    const uint32_t o0 = cls_out<TCTX>(p.A0, p.B0, p.C0), o1 = cls_out<TCTX>(p.A1, p.B1, p.C1);
    c += (__popc(o0 | o1) << 16) + (__popc((o0 & ~p.D0) | (o1 & ~p.D1)) << 24);
  }
  return c;
}

// calls of the requested contexts (slot order CHH, CHG, CG among the requested) and their methylated subsets
template <int WANT>
__device__ __forceinline__ void want_masks(const Planes& p, uint32_t (&m)[3], uint32_t (&meth)[3]) {
  int slot = 0;
  m[0] = m[1] = m[2] = 0u;
  meth[0] = meth[1] = meth[2] = 0u;
  if (WANT & W_CHH) { const uint32_t a = cls_chh(p.A0, p.B0, p.C0), b = cls_chh(p.A1, p.B1, p.C1); m[slot] = a | b; meth[slot] = (a & ~p.D0) | (b & ~p.D1); slot++; }
  if (WANT & W_CHG) { const uint32_t a = cls_chg(p.A0, p.B0, p.C0), b = cls_chg(p.A1, p.B1, p.C1); m[slot] = a | b; meth[slot] = (a & ~p.D0) | (b & ~p.D1); slot++; }
  if (WANT & W_CG) { const uint32_t a = cls_cg(p.A0, p.B0, p.C0), b = cls_cg(p.A1, p.B1, p.C1); m[slot] = a | b; meth[slot] = (a & ~p.D0) | (b & ~p.D1); slot++; }
}

// Shared-memory events of one in-window chunk.  w0q / tq point at the chunk's byte 0 in the strand's W0 / T[0]
// arrays (guard offset included); methpass = methylated calls to be counted now (0 in the fused kernel).
template <int NC, int SG>
__device__ __forceinline__ void chunk_events(const Planes& p, const uint32_t (&m)[3], uint32_t methpass, uint32_t* __restrict__ w0q,
                                             uint32_t* __restrict__ tq, int rel16) {
  const uint32_t e0 = p.A0 & p.B0 & ~p.C0, e1 = p.A1 & p.B1 & ~p.C1;
  uint32_t E = (e0 & p.D0) | (e1 & p.D1);  // code 11 bytes
  if (__popc(E) > 2) {  // runs (mate gaps): -1 where a run starts, +1 right after it ends; clamped to the window start
    const uint32_t prev = (E << 8) | ((E >> 20) & 0x30u) | ((E >> 27) & 0x2u);  // "byte before me is 11"
    const uint32_t starts = E & ~prev;
    uint32_t* const w0 = w0q - rel16;  // index 0 = guard start
    for (uint32_t mm = starts | (prev & ~E); mm;) {
      const uint32_t bit = mm & (0u - mm);
      const int q = rel16 + bit_pos(__ffs(mm) - 1);
      mm ^= bit;
      atomicAdd(&w0[max(q, kGuardLo)], (starts & bit) ? 0xFFFFFFFFu : 1u);
    }
    if ((E >> 29) & 1u) atomicAdd(&w0[max(rel16 + 16, kGuardLo)], 1u);  // the run reaches the chunk end
    E = 0u;
  }
  uint32_t ev = m[0] | m[1] | m[2] | E;
  while (ev) {
    const uint32_t bit = ev & (0u - ev);
    const int bp = bit_pos(__ffs(ev) - 1);
    ev ^= bit;
    uint32_t* ptr = tq;
    uint32_t val = (methpass & bit) ? 0x10001u : 1u;
    if (NC > 1 && (m[1] & bit)) ptr = tq + 2 * SG;
    if (NC > 2 && (m[2] & bit)) ptr = tq + 4 * SG;
    if (E & bit) { ptr = w0q; val = 0x10000u; }
    atomicAdd(ptr + bp, val);
  }
}

// methylated calls of a passing row's chunk (fused kernel, after the flag is known)
__device__ __forceinline__ void push_meth(uint32_t mm, uint32_t* __restrict__ tq) {
  while (mm) {
    const int bp = bit_pos(__ffs(mm) - 1);
    mm &= mm - 1u;
    atomicAdd(tq + bp, 0x10000u);
  }
}

template <int TCTX, int WANT, bool FUSED, int S>
__global__ void __launch_bounds__(kCplThreads, 4)
k_window_cpl(Csr csr, WinTable wt, ThrParams thr, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int rpt, int tc, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int SG = kGuardLo + S + kGuardHi;
  constexpr int MW = FUSED ? (NC == 3 ? 2 : 1) : 0;  // mask words per chunk
  constexpr int PER = S / kCplThreads;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ long long s_scan64[kCplWarps + 1];
  __shared__ int s_scan[kCplWarps + 1];
  __shared__ long long s_base;
  __shared__ int s_npass;

  uint32_t* const W0 = reinterpret_cast<uint32_t*>(smem_raw);  // [2][SG]
  uint32_t* const T = W0 + 2 * SG;                             // [NC][2][SG]
  constexpr int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const aux_d = T + NC * 2 * SG + warp * (1 + MW) * tc;  // descriptor, then count word
  uint32_t* const aux_m = aux_d + tc;                              // methylated-call masks (slots 0,1)
  uint32_t* const aux_m2 = aux_m + tc;                             // slot 2 (CX reports)

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < kZero4; i += kCplThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_npass = 0;
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    int npass_w = 0;

    for (int base = lo + warp * rpt; base < hi; base += kCplWarps * rpt) {
      // ------------------------------ rows: metadata, coverage deltas, chunk descriptors ------------------------------
      const int tend = min(base + rpt, hi);
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
      uint32_t* const w0s = W0 + strand0 * SG + kGuardLo;
      {
        uint32_t d = kDescSkip, step = 0u;
        if (overlaps) {
          atomicAdd(&w0s[max(rel0, 0)], 1u);
          const int e = rel0 + len;
          if (e < S) atomicAdd(&w0s[e], 0xFFFFFFFFu);
          d = (uint32_t)(rel0 + kGuardLo + (int)kDescBias) | ((uint32_t)strand0 << 31);
          if (!FUSED && (pass_in == nullptr || pass_in[r] != 0)) d |= 1u << 30;
          step = 16u;
        }
        for (int c = 0; c < nch; c++) aux_d[cbeg + c] = d + step * (uint32_t)c;
      }
      __syncwarp();

      // ------------------------------ chunks: classification, counts, pass-independent events ------------------------------
      const uint8_t* const tile = csr.data + tile_o0;
      for (int j0 = lane; j0 < n_chunks; j0 += 32 * kCplBatch) {
        uint4 v[kCplBatch];
#pragma unroll
        for (int u = 0; u < kCplBatch; u++) {
          const int j = j0 + 32 * u;
          if (j < n_chunks) v[u] = ldg_chunk(tile + (size_t)j * 16);
        }
#pragma unroll
        for (int u = 0; u < kCplBatch; u++) {
          const int j = j0 + 32 * u;
          if (j >= n_chunks) continue;
          const uint32_t d = aux_d[j];
          if (d == kDescSkip) continue;
          Planes p;
          make_planes(v[u], p);
          const int rel16 = (int)(d & 0x7FFFu) - (int)kDescBias;  // window-relative position of byte 0, + kGuardLo
          const bool inwin = (unsigned)(rel16 - 1) < (unsigned)(S + 15);
          const int so = (int)(d >> 31) * SG + rel16;
          uint32_t m[3], meth[3];
          want_masks<WANT>(p, m, meth);
          if (FUSED) {
            const uint32_t cnt = count4<TCTX>(p);
            if (inwin) chunk_events<NC, SG>(p, m, 0u, W0 + so, T + so, rel16);
            aux_d[j] = cnt;
            aux_m[j] = inwin ? (meth[0] | (meth[1] << 2)) : 0u;
            if (NC == 3) aux_m2[j] = inwin ? meth[2] : 0u;
          } else {
            if (inwin) chunk_events<NC, SG>(p, m, (d & (1u << 30)) ? (meth[0] | meth[1] | meth[2]) : 0u, W0 + so, T + so, rel16);
          }
        }
      }

      // ------------------------------ rows: class counts -> pass flag -> methylated calls of passing rows ------------------------------
      if (FUSED) {
        __syncwarp();
        if (overlaps) {
          uint32_t part = 0, acc_in = 0, acc_out = 0;
          unsigned long long mb = 0ull;
          for (int c = 0; c < nch; c++) {
            part += aux_d[cbeg + c];
            uint32_t any = aux_m[cbeg + c];
            if (NC == 3) any |= aux_m2[cbeg + c];
            if (any) mb |= 1ull << c;
            if ((c & 7) == 7) {
              acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
              acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
              part = 0;
            }
          }
          acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
          acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
          const bool pass = eval_pass(acc_in >> 16, acc_in & 0xFFFFu, acc_out >> 16, acc_out & 0xFFFFu, thr);
          if (rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
          if (pass) {
            uint32_t* const tq0 = T + strand0 * SG + rel0 + kGuardLo;
            while (mb) {
              const int c = __ffsll((long long)mb) - 1;
              mb &= mb - 1ull;
              const uint32_t mm = aux_m[cbeg + c];
              push_meth(mm & 0x33333333u, tq0 + 16 * c);
              if (NC > 1) push_meth((mm >> 2) & 0x33333333u, tq0 + 16 * c + 2 * SG);
              if (NC > 2) push_meth(aux_m2[cbeg + c], tq0 + 16 * c + 4 * SG);
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

    // ------------------------------ finalize: coverage prefix sum, row detection, ordered compaction ------------------------------
    const int q0 = threadIdx.x * PER;
    int cov[2][PER];
    long long packed = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      int run = 0;
#pragma unroll
      for (int v4 = 0; v4 < PER / 4; v4++) {
        const int4 d = *reinterpret_cast<const int4*>(&W0[s * SG + kGuardLo + q0 + v4 * 4]);
        const int w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int lo16 = (int)(short)w[j];  // coverage delta
          run += lo16;
          cov[s][v4 * 4 + j] = run - ((w[j] - lo16) >> 16);  // minus the '-' bytes at this position (prefix added later)
        }
      }
      packed += (long long)(run + (1 << 20)) << (32 * s);  // biased so both halves stay non-negative
    }
    long long tot64;
    const long long ex = block_exscan64_n<kCplWarps>(packed, &tot64, s_scan64);
    const int bias = (int)threadIdx.x << 20;
    const int pre0 = (int)(ex & 0xFFFFFFFFll) - bias, pre1 = (int)(ex >> 32) - bias;
    uint32_t flags = 0;  // bit 2k+s: (q0+k, strand s) is a report row
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int v4 = 0; v4 < PER / 4; v4++) {
          const uint4 v = *reinterpret_cast<const uint4*>(&T[(c * 2 + s) * SG + kGuardLo + q0 + v4 * 4]);
          if (v.x | v.y | v.z | v.w) {
            const uint32_t t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const uint32_t half = (uint32_t)(cov[s][v4 * 4 + j] + (s ? pre1 : pre0)) >> 1;
              if ((t[j] & 0xFFFFu) > half) flags |= 1u << (2 * (v4 * 4 + j) + s);
            }
          }
        }
    int nrows_total;
    const int my_off = block_exscan_n<kCplWarps>(__popc(flags), &nrows_total, s_scan);
    if (threadIdx.x == 0) {
      const unsigned long long bb = nrows_total ? atomicAdd(out.cursor, (unsigned long long)nrows_total) : 0ull;
      s_base = (long long)bb;
      out.w_base[win] = (long long)bb;
      out.w_cnt[win] = nrows_total;
      if (FUSED && s_npass) atomicAdd(out.n_pass, (unsigned long long)s_npass);
    }
    __syncthreads();
    const long long wbase = s_base;
    if (flags && wbase + nrows_total <= out.cap) {
      long long dst = wbase + my_off;
#pragma unroll
      for (int k = 0; k < PER; k++)
#pragma unroll
        for (int s = 0; s < 2; s++) {
          if (!((flags >> (2 * k + s)) & 1u)) continue;
          const uint32_t half = (uint32_t)(cov[s][k] + (s ? pre1 : pre0)) >> 1;
          int code = 0; uint32_t val = 0; int slot = 0;
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            if (!((WANT >> cc) & 1)) continue;
            const uint32_t v = T[(slot * 2 + s) * SG + kGuardLo + q0 + k];
            if ((v & 0xFFFFu) > half) { code = cc == 0 ? 2 : (cc == 1 ? 6 : 7); val = v; }
            slot++;
          }
          out.scratch[dst++] = make_int4(w0 + q0 + k, (s + 1) | (code << 8), (int)(val >> 16), (int)((val & 0xFFFFu) - (val >> 16)));
        }
    }
  }
}

}  // namespace ea
