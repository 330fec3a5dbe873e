// Stand-alone per-read thresholding / beta values for short reads (rcpp_threshold_reads, rcpp_get_xm_beta;
// /root/reference/src/rcpp_threshold_reads.cpp:27-71, src/rcpp_get_xm_beta.cpp:18-40), row per lane: a lane walks its own
// row in 32-byte units straight from HBM (256-bit loads when rows are 32-byte aligned), the four class counts stay in
// one register of byte-wide fields, 32 rows are decided per instruction with the exact integer tables of win_common.cuh.
// No shared memory, no cross-lane traffic.  Rows of more than 1024 bytes (the tables' range) and long-read data sets
// keep the lane-group kernel k_threshold (kernels.cuh), which spreads one row over many lanes.
#pragma once
#include "win_common.cuh"
#include "window_long.cuh"  // load_unit

namespace ea {

#ifndef EA_THRROW_MINB
#define EA_THRROW_MINB 5
#endif

template <int TCTX, bool A32>
__global__ void __launch_bounds__(256, EA_THRROW_MINB)
k_threshold_row(Csr csr, ThrParams thr, ThrTables tbl, uint8_t* __restrict__ pass_out, double* __restrict__ beta_out,
                unsigned long long* n_pass) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  __shared__ uint16_t s_mmin[kThrTable], s_omax[kThrTable];  // pass-chain tables, staged (win_common.cuh)
  if (pass_out) {
    stage_thr_tables(tbl, s_mmin, s_omax, kThrTable - 1);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  uint32_t my_pass = 0;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < csr.n; base += stride) {
    const int64_t tend = min(base + 32, csr.n);
    const int64_t r = base + lane;
    const bool valid = r < tend;
    const long long o0 = csr.off[valid ? r : tend];
    long long o1 = __shfl_down_sync(FULL, o0, 1);
    if (lane == 31) o1 = csr.off[tend];
    if (!valid) continue;
    const int nch = (int)((o1 - o0) >> 4);
    const int nu = (nch + 1) >> 1;
    const uint8_t* row = csr.data + o0;
    uint32_t cnt = 0, acc_in = 0, acc_out = 0;
    uint4 va, vb, na, nb;
    if (nu > 0) load_unit<A32>(row, nch > 1, va, vb);
    for (int u = 0; u < nu; u++) {
      if (u + 1 < nu) load_unit<A32>(row + 32 * (u + 1), 2 * u + 3 < nch, na, nb);
      {
        Planes p;
        make_planes(va, p);
        cnt += count4<TCTX>(p);
      }
      {
        Planes p;
        make_planes(vb, p);
        cnt += count4<TCTX>(p);
      }
      if ((u & 3) == 3) {  // byte-wide fields hold the counts of four units
        acc_in += (cnt & 0xFFu) | ((cnt & 0xFF00u) << 8);
        acc_out += ((cnt >> 16) & 0xFFu) | ((cnt >> 24) << 16);
        cnt = 0;
      }
      va = na; vb = nb;
    }
    acc_in += (cnt & 0xFFu) | ((cnt & 0xFF00u) << 8);
    acc_out += ((cnt >> 16) & 0xFFu) | ((cnt >> 24) << 16);
    const uint32_t m = acc_in >> 16, all = acc_in & 0xFFFFu;
    if (pass_out) {
      const bool pass = eval_pass_staged(m, all, acc_out >> 16, acc_out & 0xFFFFu, thr, s_mmin, s_omax, kThrTable - 1);
      pass_out[r] = pass ? 1 : 0;
      my_pass += pass;
    }
    if (beta_out) beta_out[r] = (double)m / (double)(all ? all : 1u);
  }
  if (n_pass) {
    for (int d = 16; d >= 1; d >>= 1) my_pass += __shfl_down_sync(FULL, my_pass, d);
    if (lane == 0 && my_pass) atomicAdd(n_pass, (unsigned long long)my_pass);
  }
}

}  // namespace ea
