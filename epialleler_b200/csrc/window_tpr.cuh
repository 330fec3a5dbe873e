// Short-read window kernel, "thread per row over TMA-staged tiles" formulation.
//
// One CTA per genomic window of S positions.  Each warp bulk-copies a tile of up to 32 consecutive padded
// rows (one contiguous, 16-byte aligned byte range of the CSR) into shared memory with ONE cp.async.bulk
// (TMA) completed on a warp-private mbarrier, then every lane walks ITS OWN row with LDS.128 in a single
// sweep:
//   * class counts for the thresholds stay in registers (no cross-lane reduction, no staging of counts);
//   * everything that does not depend on the pass flag - {total calls} of the requested contexts, '-'
//     bytes, row start/end coverage deltas - goes to the shared-memory window state as reductions
//     (win_common.cuh);
//   * chunks that hold methylated calls of a requested context are remembered in a per-row bit mask and
//     their call masks overwrite the chunk in place; once the row's flag is known (the reference's
//     pass/fail chain, src/rcpp_threshold_reads.cpp:39-70, through the exact integer tables of
//     k_thr_tables), a passing row revisits just those chunks and adds its methylated calls
//     (SURVEY.md 8a invariant 4: only the methylated count depends on `pass`).
// Row metadata loads are coalesced over 32 rows, the pass chain runs for 32 rows per instruction.
// Rows of very different lengths leave lanes idle at the end of a tile; the dispatcher prefers the
// chunk-per-lane kernel (window_cpl.cuh) for such data.
// Preconditions (checked by the dispatcher): padded rows <= 512 bytes and <= tile_bytes, at most 32767
// rows overlap a window.
#pragma once
#include "win_common.cuh"

namespace ea {

constexpr int kTprThreads = 256;
constexpr int kTprWarps = kTprThreads / 32;

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

template <int TCTX, int WANT, bool FUSED, int S>
__global__ void __launch_bounds__(kTprThreads, 3)
k_window_tpr(Csr csr, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int tile_bytes, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int SG = kGuardLo + S + kGuardHi;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTprWarps];
  __shared__ uint32_t s_scan[2 * kTprWarps];
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ __align__(4) uint8_t s_bp4[32];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  constexpr int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* const my_tile = smem_raw + sizeof(uint32_t) * (NC + 1) * 2 * SG + (size_t)warp * tile_bytes;
  uint64_t* const my_bar = &s_bar[warp];
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T), bp_tbl = smem_u32(s_bp4);
  const uint32_t w_floor = w_base + 4u * kGuardLo;
  const uint32_t tile_addr = smem_u32(my_tile);
  uint32_t phase = 0;

  if (threadIdx.x < 32) s_bp4[threadIdx.x] = (uint8_t)(4 * bit_pos32(threadIdx.x));
  if (lane == 0) mbar_init(my_bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

  // the window state is zeroed once; every finalize leaves it zeroed again (guard bands are never read)
  for (int i = threadIdx.x; i < kZero4; i += kTprThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    // the window's rows are split evenly over the warps, each warp's share into equal tiles of at most 32 rows
    const int per_warp = (hi - lo + kTprWarps - 1) / kTprWarps;
    const int my_lo = lo + warp * per_warp, my_hi = min(hi, my_lo + per_warp);
    const int n_t = (per_warp + 31) >> 5;
    const int per_tile = n_t ? (per_warp + n_t - 1) / n_t : 0;
    int npass_w = 0;

    for (int base = my_lo; base < my_hi; base += per_tile) {
      const int tend = min(base + per_tile, my_hi);
      const int r = base + lane;
      const bool valid = r < tend;
      const long long o0 = csr.off[valid ? r : tend];
      long long o1 = __shfl_down_sync(FULL, o0, 1);
      if (lane == 31) o1 = csr.off[tend];
      int start = 0, spv = 0;
      if (valid) { start = csr.start[r]; spv = csr.sp[r]; }
      const int nvalid = tend - base;
      int first = 0;
      while (first < nvalid) {
        // rows [first, first+k) of this step whose bytes fit one tile (normally all of them)
        const long long of = __shfl_sync(FULL, o0, first);
        const uint32_t fit = __ballot_sync(FULL, valid && lane >= first && (o1 - of) <= (long long)tile_bytes) >> first;
        const int k = (~fit == 0u) ? 32 : max(1, __ffs(~fit) - 1);  // leading run of fitting rows (>= 1 by the dispatcher's precondition)
        const long long oe = __shfl_sync(FULL, o1, first + k - 1);
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
        const int strand0 = spv & 1;
        const int len = nch * 16 - (spv >> 4);
        const int rel0 = start - w0;
        const bool overlaps = mine && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
        if (overlaps) {
          const uint32_t sinc = strand0 ? 0x10000u : 1u;
          red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
          const int e = rel0 + len;
          if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
          const uint32_t row_addr = tile_addr + (uint32_t)(o0 - of);
          // byte addresses of the row's position 0 in W/D (strand independent) and in T[0][strand0]
          const uint32_t rel4 = 4u * (uint32_t)(rel0 + kGuardLo);
          const uint32_t wq = w_base + rel4, dq = d_base + rel4, tq = t_base + 4u * (uint32_t)(strand0 * SG) + rel4;
          const bool pass_known = FUSED ? false : (pass_in == nullptr || pass_in[r] != 0);
          uint32_t part = 0, acc_in = 0, acc_out = 0, mb = 0;
          for (int c = 0; c < nch; c++) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(row_addr + 16u * c));
            Planes p;
            make_planes(v, p);
            uint32_t m[3], meth[3];
            want_masks<WANT>(p, m, meth);
            const int rel = rel0 + 16 * c;
            const bool inwin = (unsigned)(rel + 15) < (unsigned)(S + 15);  // -16 < rel < S
            if (FUSED) {
              part += count4<TCTX>(p);
              if ((c & 7) == 7) {
                acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
                acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
                part = 0;
              }
            }
            if (inwin) {
              chunk_events<NC, SG>(p, m, (!FUSED && pass_known) ? (meth[0] | meth[1] | meth[2]) : 0u, wq + 64u * c, dq + 64u * c, tq + 64u * c, sinc,
                                   w_floor, bp_tbl);
              if (FUSED && (meth[0] | meth[1] | meth[2])) {  // remember the chunk and leave its methylated-call masks in place
                mb |= 1u << c;
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(row_addr + 16u * c), "r"(meth[0] | (meth[1] << 2)), "r"(meth[2]) : "memory");
              }
            }
          }
          if (FUSED) {
            acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
            acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
            const bool pass = eval_pass_tables(acc_in >> 16, acc_in & 0xFFFFu, acc_out >> 16, acc_out & 0xFFFFu, thr, tbl);
            if (rel0 >= 0) {  // the window that holds the row's start owns its flag
              pass_out[r] = pass ? 1 : 0;
              npass_w += pass;
            }
            if (pass) {
              while (mb) {
                const int c = 31 - __clz(mb);
                mb ^= 1u << c;
                uint32_t m01, m2;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(m01), "=r"(m2) : "r"(row_addr + 16u * c));
                push_meth(m01 & 0x33333333u, tq + 64u * c, bp_tbl);
                if (NC > 1) push_meth((m01 >> 2) & 0x33333333u, tq + 64u * c + 2 * SG * 4, bp_tbl);
                if (NC > 2) push_meth(m2, tq + 64u * c + 4 * SG * 4, bp_tbl);
              }
            }
          }
        }
        // rows were overwritten in place (generic proxy): order those writes before the next TMA into this tile
        if (FUSED) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        first += k;
      }
    }
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(FULL, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    __syncthreads();

    window_finalize<WANT, S, kTprThreads, FUSED>(W, D, T, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

}  // namespace ea
