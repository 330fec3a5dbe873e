// Short-read window kernel: one THREAD per read over TMA-staged tiles.
//
// Why a second formulation next to k_window (kernels.cuh): with ~150-byte reads the lane-group
// kernel spends most of its issue slots on per-row overheads that are paid per warp instruction
// (row metadata, cross-lane reductions, pass evaluation) and on a divergent event loop.  Here
//   * each warp bulk-copies a tile of up to 32 consecutive padded rows (one contiguous, 16-byte
//     aligned byte range of the CSR) into shared memory with ONE cp.async.bulk (TMA) completed on a
//     warp-private mbarrier, then every lane walks its own row with LDS.128: no shuffles, row
//     metadata loads are coalesced over 32 rows, the IEEE-double pass chain runs for 32 rows at once;
//   * ONE sweep over the row does everything that does not depend on the pass flag: class counts for
//     the thresholds, {total calls} of the requested contexts and coverage deltas (row start/end,
//     '-' runs) as shared-memory atomics.  By SURVEY.md 8a invariant 4 only the methylated count
//     depends on `pass`, so after the flag is known the few passing rows (~10%) get a
//     warp-cooperative fix-up: the 32 lanes take one chunk each of such a row and add its
//     methylated calls;
//   * the finalize touches shared memory only: vectorised prefix sum of the coverage deltas, row
//     detection (total_ctx > floor(cov/2), src/rcpp_cx_report.cpp:62-72), ordered compaction.
// Preconditions (checked by the host dispatcher): every padded row fits a tile, at most 32767 rows
// overlap a window (16-bit packed counters).
#pragma once
#include "kernels.cuh"

namespace ea {

constexpr int kTprThreads = 256;
constexpr int kTprWarps = kTprThreads / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
// 1-D TMA: global -> shared bulk copy, bytes and both addresses multiples of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NWARPS>
__device__ __forceinline__ int block_exscan_n(int v, int* total, int* s_warp /* NWARPS+1 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int x = lane < NWARPS ? s_warp[lane] : 0;
    int xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < NWARPS) s_warp[lane] = xi - x;
    if (lane == NWARPS - 1) s_warp[NWARPS] = xi;
  }
  __syncthreads();
  *total = s_warp[NWARPS];
  return s_warp[warp] + inc - v;
}


// position (byte index within the chunk) of merged-indicator bit i
__device__ __forceinline__ int bit_pos(int i) { return ((i & 1) << 3) + (i & 4) + (i >> 3); }

// One shared-memory atomic per set bit of m: t[q] += 1 (+ 0x10000 when the bit is also in `meth`).
// The first bit is straight-line code (almost every warp step has some lane with an event), the rest loop.
__device__ __forceinline__ void push_mask(uint32_t m, uint32_t meth, uint32_t* __restrict__ t, int rel, int S) {
  if (m) {
    uint32_t bit = m & (0u - m);
    int q = rel + bit_pos(__ffs(m) - 1);
    if ((unsigned)q < (unsigned)S) atomicAdd(&t[q], (meth & bit) ? 0x10001u : 1u);
    m ^= bit;
    while (m) {
      bit = m & (0u - m);
      q = rel + bit_pos(__ffs(m) - 1);
      if ((unsigned)q < (unsigned)S) atomicAdd(&t[q], (meth & bit) ? 0x10001u : 1u);
      m ^= bit;
    }
  }
}

// Coverage events of one chunk (independent of the pass flag).  cv = coverage words of the row's strand.
// A coverage word packs two 16-bit fields that are decoded at the end: low = running-coverage DELTA
// (+1 row start, -1 true row end, -1/+1 around long '-' runs; two's-complement arithmetic keeps the
// packed sum exact), high = number of isolated '-' bytes at the position.
// cov(p) = prefix_sum(low)(p) - high(p).
__device__ __forceinline__ void push_cov(const Planes& p, int32_t* __restrict__ cv, int S, int rel, bool first_in_win, bool is_last) {
  const uint32_t E = (p.A0 & p.B0 & ~p.C0 & p.D0) | (p.A1 & p.B1 & ~p.C1 & p.D1);  // code 11 bytes
  if (first_in_win) atomicAdd(&cv[max(rel, 0)], 1);
  if (E) {
    if (__popc(E) <= 2) {  // isolated '-' bytes: one count each
      uint32_t m = E;
      int q = rel + bit_pos(__ffs(m) - 1);
      if ((unsigned)q < (unsigned)S) atomicAdd(&cv[q], 0x10000);
      m &= m - 1u;
      if (m) {
        q = rel + bit_pos(__ffs(m) - 1);
        if ((unsigned)q < (unsigned)S) atomicAdd(&cv[q], 0x10000);
      }
    } else {               // runs (mate gaps): -1 where a run starts, +1 right after it ends
      const uint32_t prev = (E << 8) | ((E >> 20) & 0x30u) | ((E >> 27) & 0x2u);  // "byte before me is 11"
      const uint32_t starts = E & ~prev;
      for (uint32_t m = starts | (prev & ~E); m;) {
        const uint32_t bit = m & (0u - m);
        const int q = rel + bit_pos(__ffs(m) - 1);
        m ^= bit;
        if (q < S) atomicAdd(&cv[max(q, 0)], (starts & bit) ? -1 : 1);
      }
      if (((E >> 29) & 1u) && rel + 16 < S) atomicAdd(&cv[max(rel + 16, 0)], 1);  // run reaches the chunk end
    }
  }
  if (is_last) {
    const int qe = rel + 16 - pad_count(p);
    if (qe < S) atomicAdd(&cv[max(qe, 0)], -1);
  }
}

// requested-context call masks of a chunk, one per context slot (order CHH, CHG, CG among the requested),
// plus the methylated ones; merged indicator layout
template <int WANT>
__device__ __forceinline__ uint4 call_masks(const Planes& p) {
  uint32_t m[3] = {0u, 0u, 0u};
  int slot = 0;
  uint32_t meth = 0;
  if (WANT & W_CHH) { const uint32_t a = cls_chh(p.A0, p.B0, p.C0), b = cls_chh(p.A1, p.B1, p.C1); m[slot++] = a | b; meth |= (a & ~p.D0) | (b & ~p.D1); }
  if (WANT & W_CHG) { const uint32_t a = cls_chg(p.A0, p.B0, p.C0), b = cls_chg(p.A1, p.B1, p.C1); m[slot++] = a | b; meth |= (a & ~p.D0) | (b & ~p.D1); }
  if (WANT & W_CG) { const uint32_t a = cls_cg(p.A0, p.B0, p.C0), b = cls_cg(p.A1, p.B1, p.C1); m[slot++] = a | b; meth |= (a & ~p.D0) | (b & ~p.D1); }
  return make_uint4(m[0], m[1], m[2], meth);
}

template <int WANT>
__device__ __forceinline__ void push_calls(const uint4& cm, uint32_t meth_if_pass, uint32_t* __restrict__ tms, int S, int rel) {
  constexpr int NC = n_ctx<WANT>();
  push_mask(cm.x, meth_if_pass, tms, rel, S);
  if (NC > 1) push_mask(cm.y, meth_if_pass, tms + 2 * S, rel, S);
  if (NC > 2) push_mask(cm.z, meth_if_pass, tms + 4 * S, rel, S);
}

template <int NWARPS>
__device__ __forceinline__ long long block_exscan64_n(long long v, long long* total, long long* s_warp /* NWARPS+1 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const long long x = lane < NWARPS ? s_warp[lane] : 0;
    long long xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < NWARPS) s_warp[lane] = xi - x;
    if (lane == NWARPS - 1) s_warp[NWARPS] = xi;
  }
  __syncthreads();
  *total = s_warp[NWARPS];
  return s_warp[warp] + inc - v;
}

template <int TCTX, int WANT, bool FUSED, int S>
__global__ void __launch_bounds__(kTprThreads, 3)
k_window_tpr(Csr csr, WinTable wt, ThrParams thr, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int tile_bytes, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int Sp = S + 4;
  constexpr int PER = S / kTprThreads;  // positions per thread in the finalize (8 or 16)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTprWarps];
  __shared__ long long s_scan64[kTprWarps + 1];
  __shared__ int s_scan[kTprWarps + 1];
  __shared__ long long s_base;
  __shared__ int s_npass;

  uint32_t* tm = reinterpret_cast<uint32_t*>(smem_raw);                          // [NC][2][S]
  int32_t* covd = reinterpret_cast<int32_t*>(smem_raw + sizeof(uint32_t) * NC * 2 * S);  // [2][Sp]
  unsigned char* tiles = smem_raw + sizeof(uint32_t) * (NC * 2 * S + 2 * Sp);     // [kTprWarps][tile_bytes]
  constexpr int kZero4 = (NC * 2 * S + 2 * Sp) / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* my_tile = tiles + (size_t)warp * tile_bytes;
  uint64_t* my_bar = &s_bar[warp];
  uint32_t phase = 0;

  if (lane == 0) mbar_init(my_bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < kZero4; i += kTprThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_npass = 0;
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];

    // ---------------- rows: 32 per warp step, staged through shared memory by TMA ----------------
    for (int base = lo + warp * 32; base < hi; base += kTprWarps * 32) {
      const int r = base + lane;
      const bool valid = r < hi;
      const int rc = valid ? r : hi;  // off[hi] exists (off has n+1 entries)
      const long long o0 = csr.off[rc];
      long long o1 = __shfl_down_sync(0xFFFFFFFFu, o0, 1);
      if (lane == 31) o1 = valid ? csr.off[r + 1] : o0;
      int start = 0, strand0 = 0;
      if (valid) { start = csr.start[r]; strand0 = csr.strand[r] - 1; }
      const int nvalid = min(32, hi - base);
      int first = 0;
      while (first < nvalid) {
        // rows [first, first+k) of this step whose bytes fit one tile
        const long long of = __shfl_sync(0xFFFFFFFFu, o0, first);
        const uint32_t fit = __ballot_sync(0xFFFFFFFFu, valid && lane >= first && (o1 - of) <= (long long)tile_bytes) >> first;
        const int k = (~fit == 0u) ? 32 : max(1, __ffs(~fit) - 1);  // leading run of fitting rows (>=1 by the dispatcher's precondition)
        const long long oe = __shfl_sync(0xFFFFFFFFu, o1, first + k - 1);
        const uint32_t bytes = (uint32_t)(oe - of);
        if (bytes) {
          if (lane == 0) {
            mbar_expect_tx(my_bar, bytes);
            tma_load_1d(my_tile, csr.data + of, bytes, my_bar);
          }
          mbar_wait(my_bar, phase);
          phase ^= 1u;
        }
        const bool mine = valid && lane >= first && lane < first + k;
        const int nch = mine ? (int)((o1 - o0) >> 4) : 0;
        const int row_off = (int)(o0 - of);
        const int rel0 = start - w0;
        const bool overlaps = mine && rel0 + nch * 16 > 0;
        const int c_first = rel0 >= 0 ? 0 : ((-rel0) >> 4);
        const int c_end = rel0 + nch * 16 > S ? ((S - rel0 + 15) >> 4) : nch;
        if (overlaps) {
          unsigned char* row = my_tile + row_off;
          int32_t* cv = covd + strand0 * Sp;
          uint32_t* tms = tm + strand0 * S;
          if (FUSED) {
            // sweep 1: class counts, coverage events; the call masks replace the chunk in place
            uint32_t c0 = 0, c1 = 0;
            for (int c = 0; c < nch; c++) {
              Planes p;
              make_planes(*reinterpret_cast<const uint4*>(row + c * 16), p);
              count_chunk<TCTX>(p, c0, c1);
              if (c >= c_first && c < c_end) {
                push_cov(p, cv, S, rel0 + c * 16, c == c_first, c == nch - 1);
                const uint4 cm = call_masks<WANT>(p);
                if (NC == 1) *reinterpret_cast<uint2*>(row + c * 16) = make_uint2(cm.x, cm.w);
                else *reinterpret_cast<uint4*>(row + c * 16) = cm;
              }
            }
            const bool pass = eval_pass(c0 >> 16, c0 & 0xFFFFu, c1 >> 16, c1 & 0xFFFFu, thr);
            if (rel0 >= 0) {  // the window that holds the row's start owns its flag
              pass_out[r] = pass ? 1 : 0;
              if (pass) atomicAdd(&s_npass, 1);
            }
            // sweep 2: the calls, now that the flag is known ({total, methylated-of-passing} in one atomic)
            const uint32_t pm = pass ? 0xFFFFFFFFu : 0u;
            for (int c = c_first; c < c_end; c++) {
              uint4 cm;
              if (NC == 1) { const uint2 v = *reinterpret_cast<const uint2*>(row + c * 16); cm = make_uint4(v.x, 0u, 0u, v.y); }
              else cm = *reinterpret_cast<const uint4*>(row + c * 16);
              push_calls<WANT>(cm, cm.w & pm, tms, S, rel0 + c * 16);
            }
          } else {
            const bool pass = pass_in ? (pass_in[r] != 0) : true;
            const uint32_t pm = pass ? 0xFFFFFFFFu : 0u;
            for (int c = c_first; c < c_end; c++) {
              Planes p;
              make_planes(*reinterpret_cast<const uint4*>(row + c * 16), p);
              push_cov(p, cv, S, rel0 + c * 16, c == c_first, c == nch - 1);
              const uint4 cm = call_masks<WANT>(p);
              push_calls<WANT>(cm, cm.w & pm, tms, S, rel0 + c * 16);
            }
          }
        }
        // the rows were overwritten in place (generic proxy): order those writes before the next TMA into this tile
        if (FUSED) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        first += k;
      }
    }
    __syncthreads();

    // ---------------- finalize: coverage prefix sum, row detection, ordered compaction ----------------
    const int q0 = threadIdx.x * PER;
    int cov[2][PER];
    long long packed = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      int run = 0;
#pragma unroll
      for (int v4 = 0; v4 < PER / 4; v4++) {
        const int4 d = *reinterpret_cast<const int4*>(&covd[s * Sp + q0 + v4 * 4]);
        const int w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int lo16 = (int)(short)w[j];            // coverage delta
          run += lo16;
          cov[s][v4 * 4 + j] = run - ((w[j] - lo16) >> 16);  // minus the '-' bytes at this position (prefix added later)
        }
      }
      packed += (long long)(run + (1 << 20)) << (32 * s);  // biased so both halves stay non-negative
    }
    long long tot64;
    const long long ex = block_exscan64_n<kTprWarps>(packed, &tot64, s_scan64);
    const int bias = (int)threadIdx.x << 20;
    const int pre0 = (int)(ex & 0xFFFFFFFFll) - bias, pre1 = (int)(ex >> 32) - bias;
    uint32_t flags = 0;  // bit 2k+s: (q0+k, strand s) is a report row
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int v4 = 0; v4 < PER / 4; v4++) {
          const uint4 v = *reinterpret_cast<const uint4*>(&tm[(c * 2 + s) * S + q0 + v4 * 4]);
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
    const int my_off = block_exscan_n<kTprWarps>(__popc(flags), &nrows_total, s_scan);
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
            const uint32_t v = tm[(slot * 2 + s) * S + q0 + k];
            if ((v & 0xFFFFu) > half) { code = cc == 0 ? 2 : (cc == 1 ? 6 : 7); val = v; }
            slot++;
          }
          out.scratch[dst++] = make_int4(w0 + q0 + k, (s + 1) | (code << 8), (int)(val >> 16), (int)((val & 0xFFFFu) - (val >> 16)));
        }
    }
  }
}

}  // namespace ea
