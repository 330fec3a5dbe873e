// C-ABI implementation (include/epialleler_b200.h): packer, window index, kernel dispatch.
// There is no CPU compute path in this file: without a CUDA device every entry point fails with
// EA_ERR_NODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cstdlib>
#include <ctime>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "../../include/epialleler_b200.h"
#include "kernels.cuh"
#include "window_tpr.cuh"
#include "window_cpl.cuh"
#include "window_row.cuh"
#include "window_long.cuh"
#include "threshold_row.cuh"
#include "bed.cuh"
#include "vcf.cuh"
#include "writer.cuh"
#include "window_mhl.cuh"
#include "planar.cuh"
#include "window_vs.cuh"
#include "window_mhl_pl.cuh"
#include "synth.h"
#include "hostpack.h"

namespace {

thread_local std::string g_err;
thread_local int g_pack_threads = 0;   // > 0: host threads a packer called from this thread may use (ea_pack_multi)
ea_poll_fn g_poll = nullptr;
int g_kernel_policy = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(EA_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define POLL()                                                                   \
  do {                                                                           \
    ea_poll_fn poll_ = h->poll ? h->poll : g_poll;   /* the handle's own callback wins over the process-wide one */ \
    if (poll_ && poll_()) return fail(EA_ERR_INTERRUPT, "interrupted by caller"); \
  } while (0)

// EA_TRACE=1: wall-clock stage marks on stderr (each mark synchronizes the device; diagnostics only)
struct Trace {
  bool on; double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit Trace(const char* what) : on(getenv("EA_TRACE") != nullptr), t0(0) { if (on) { cudaDeviceSynchronize(); t0 = now(); fprintf(stderr, "[ea] %s\n", what); } }
  void mark(const char* stage) { if (on) { cudaDeviceSynchronize(); const double t = now(); fprintf(stderr, "[ea]   %-28s %8.2f ms\n", stage, t - t0); t0 = t; } }
};

// Device memory comes from a PRIVATE stream-ordered pool per device (cudaMemPoolCreate, unlimited release threshold; the
// device's default pool, which other users of cudaMallocAsync in the process share, e.g. PyTorch, is left alone): a handle
// packs and frees tens of GB per call, and plain cudaMalloc/cudaFree of such blocks costs tens to hundreds of milliseconds
// (measured, EA_TRACE) while the pool hands the cached blocks back in microseconds.  ea_trim() returns the cached memory to
// the driver.  Allocation is made synchronous here (the callers' work runs on other streams).
// Result copy into CALLER memory.  Page-locked destinations (ea_host_alloc, cudaHostRegister) take the copy as it is.
// Pageable ones - R vectors, numpy arrays: what every binding passes - would be staged by the driver through small bounce
// buffers at a few GB/s with the fresh destination pages first touched one by one; instead the data crosses PCIe in 16 MB
// pieces into two page-locked blocks from the library's cache while a few host threads copy the previous piece out (and
// first-touch the destination's pages in parallel).  Returns after the data has landed.  Defined below ea_host_free.
cudaError_t d2h_result(void* dst, const void* src, size_t bytes, cudaStream_t s);
constexpr int kMaxDevices = 64;
cudaMemPool_t g_pools[kMaxDevices] = {};
std::mutex g_pool_mu;

cudaError_t ea_pool(int dev, cudaMemPool_t* out) {
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool;
    cudaError_t e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    g_pools[dev] = pool;
  }
  *out = g_pools[dev];
  return cudaSuccess;
}

cudaError_t pool_malloc(void** p, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaMemPool_t pool;
  if ((e = ea_pool(dev, &pool)) != cudaSuccess) return e;
  e = cudaMallocFromPoolAsync(p, bytes, pool, cudaStreamPerThread);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(cudaStreamPerThread);
}
template <class T>
cudaError_t pool_malloc(T** p, size_t bytes) { return pool_malloc(reinterpret_cast<void**>(p), bytes); }

// Growing the pool is what costs: a handle is some twenty blocks, and when they are all NEW memory the allocations of one
// call add up to 0.3-1 s - and the pool keeps growing over the following calls, a stall here and there, until it happens to
// hold a fitting block for every request (measured: 250-270 ms per pack+report+free with a warm pool, 270-1070 ms after
// ea_trim()).  The packers therefore say up front what the handle will need: if the pool's free bytes fall short, ONE block
// of the deficit is allocated and freed again - everything after it is carved out of cached memory.
void pool_reserve(size_t want_free) {
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) != cudaSuccess || ea_pool(dev, &pool) != cudaSuccess) { cudaGetLastError(); return; }
  uint64_t reserved = 0, used = 0;
  if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) != cudaSuccess ||
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) != cudaSuccess) { cudaGetLastError(); return; }
  const uint64_t have = reserved > used ? reserved - used : 0;
  if (have >= want_free) return;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return; }
  const size_t deficit = std::min<size_t>(want_free - have, free_b / 2);   // never more than half of what the device has left
  if (deficit < (64u << 20)) return;
  void* p = nullptr;
  if (cudaMallocFromPoolAsync(&p, deficit, pool, cudaStreamPerThread) == cudaSuccess) {
    cudaFreeAsync(p, cudaStreamPerThread);
    cudaStreamSynchronize(cudaStreamPerThread);
  } else {
    cudaGetLastError();
  }
}

// Freeing a block that any stream may still use needs a device-wide synchronisation first (like cudaFree).  Code that
// frees many blocks in a row opens a FreeScope: the first pool_free inside it synchronises, the others only free.
thread_local int g_free_scope_depth = 0;
thread_local bool g_free_scope_synced = false;
struct FreeScope {
  FreeScope() { if (g_free_scope_depth++ == 0) g_free_scope_synced = false; }
  ~FreeScope() { --g_free_scope_depth; }
};
void pool_free(void* p) {
  if (!p) return;
  if (!(g_free_scope_depth > 0 && g_free_scope_synced)) {
    cudaDeviceSynchronize();  // nothing on any stream may still use the block
    g_free_scope_synced = true;
  }
  cudaFreeAsync(p, cudaStreamPerThread);
}

// stream-ordered free: for blocks whose every use was on stream s (no device-wide synchronisation)
void pool_free_on(void* p, cudaStream_t s) {
  if (p) cudaFreeAsync(p, s);
}

// temporaries of one call, all used on one stream: freed stream-ordered when the scope ends, early returns included
struct StreamTmp {
  cudaStream_t s;
  std::vector<void*> v;
  explicit StreamTmp(cudaStream_t stream) : s(stream) {}
  ~StreamTmp() { for (void* p : v) pool_free_on(p, s); }
  template <class T> T* keep(T* p) { v.push_back(p); return p; }
};

template <class T>
int dmalloc(T** p, size_t n) {
  *p = nullptr;
  CU(pool_malloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)));
  return EA_OK;
}

struct WinIndex {
  int S = 0;
  int32_t n = 0;
  int32_t *rname = nullptr, *pos0 = nullptr, *lo = nullptr, *hi = nullptr;
  long long *w_base = nullptr, *w_dst = nullptr;
  int32_t* w_cnt = nullptr;
  int32_t max_rows = 0;  // largest hi-lo
  void release() {
    FreeScope fs;
    pool_free(rname); pool_free(pos0); pool_free(lo); pool_free(hi); pool_free(w_base); pool_free(w_dst); pool_free(w_cnt);
    *this = WinIndex();
  }
};

}  // namespace

struct ea_handle {
  int device = 0;
  int64_t n = 0;
  int64_t nbytes = 0;   // sum of true lengths
  int64_t npadded = 0;  // bytes of the padded data array
  int32_t max_len = 0;
  bool rows32 = false;  // every padded row is a multiple of 32 bytes (pairs of chunks can be handled as one unit)
  bool has_bytes = true;   // false: packed by ea_pack_blob_calls - call planes only, no byte rows (base codes dropped)
  int64_t h2d_bytes = 0;   // bytes the packer sent over PCIe
  bool owns = true;
  int32_t* d_rname = nullptr;
  uint8_t* d_strand = nullptr;
  uint8_t* d_sp = nullptr;       // strand '-' flag | pad count << 4 (window_cpl.cuh)
  int32_t* d_start = nullptr;
  int32_t* d_len = nullptr;
  int64_t* d_off = nullptr;
  uint8_t* d_data = nullptr;
  uint64_t* d_maxend = nullptr;  // per row: rname<<32 | prefix max of row ends
  // bit-sliced call planes in tiles of 32 rows (planar.cuh); built for data sets whose rows are at most 1024 positions
  uint4* d_pl_units = nullptr;
  int64_t* d_pl_toff = nullptr;
  uint32_t* d_pl_ls = nullptr;
  int64_t pl_tiles = 0, pl_unit_rows = 0;   // tiles; 512-byte unit rows (tiles of 32 rows) or 16-byte units (tiles of one row)
  int pl_th = 32;                           // rows per tile: 32 (rows of up to 1024 positions) or 1 (long reads)
  bool pl_ok = false;
  WinIndex idx[4];               // small cache of window indexes keyed by S
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[10] = {};
  // per-call state
  uint8_t* d_pass = nullptr;
  double* d_beta = nullptr;
  unsigned long long* d_counters = nullptr;  // [0] row cursor, [1] n_pass
  int32_t* d_match = nullptr;                // per read: 1-based BED row or NA (ea_match_target)
  // last lMHL report (ea_mhl_report): scratch rows + ordered columns
  ea::MhlRow* d_mhl_scratch = nullptr;
  int32_t* d_mhl_i[5] = {};
  double* d_mhl_d[2] = {};
  int64_t mhl_cap = 0, mhl_rows = -1;
  uint16_t* d_thr_tables = nullptr;          // 2 x kThrTable integer thresholds of the pass/fail chain (window_cpl.cuh)
  int4* d_scratch = nullptr;
  int32_t* d_cols[6] = {};
  int64_t cap_rows = 0;
  int64_t n_rows = -1;
  // range sharding (ea_set_owned_range): the report keeps the rows whose key rname<<32|pos lies in [own_key_lo, own_key_hi)
  // - found on the device by two binary searches over the ordered columns - and pass flags are reported from own_row_lo on
  bool own_set = false;
  int64_t own_key_lo = 0, own_key_hi = 0, own_row_lo = 0;
  int64_t row_first = 0;           // first owned row of the last report inside d_cols
  int32_t report_n_launches = 0;
  const char* report_kernel = "";  // which window kernel the last report ran
  char* d_text = nullptr;      // last formatted report (ea_cx_report_text)
  int64_t text_bytes = -1;
  void* d_flush = nullptr;
  int64_t flush_bytes = 0;
  int sm_count = 148;
  ea_poll_fn poll = nullptr;       // per-handle interrupt poll (ea_set_handle_poll); null = the process-wide one
};

namespace {

using namespace ea;

Csr csr_of(const ea_handle* h) {
  Csr c;
  c.data = h->d_data; c.off = h->d_off; c.len = h->d_len; c.start = h->d_start; c.strand = h->d_strand; c.sp = h->d_sp;
  c.rname = h->d_rname; c.n = h->n;
  return c;
}

int pick_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(EA_ERR_NODEVICE, "no CUDA device available: epialleler_b200 has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(EA_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  CU(cudaSetDevice(device));
  return EA_OK;
}

// ---------------------------------------------------------------------------------------------
// pack-time kernels
// ---------------------------------------------------------------------------------------------
enum { F_UNSORTED = 1, F_DOMAIN = 2, F_TEMPLID = 4, F_NEGSTART = 8, F_STRAND = 16, F_TOOLONG = 32,
       F_ODDCHUNKS = 64 /* informational: some padded row is not a multiple of 32 bytes */,
       F_NONIDENT = 128 /* informational: templid is not the identity (rows are not the strings in file order) */ };

__global__ void k_lens_from_templid(const int64_t* __restrict__ soff, const int32_t* __restrict__ templid, int64_t n,
                                    int64_t nstr, int32_t* __restrict__ len, int64_t* __restrict__ padded,
                                    unsigned* flags) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t t = templid ? templid[r] : r;
  int64_t l = 0;
  if (t != r) atomicOr(flags, F_NONIDENT);
  if (t < 0 || t >= nstr) atomicOr(flags, F_TEMPLID);
  else l = soff[t + 1] - soff[t];
  if (l < 0 || l > 0x7FFFFFF0ll) { atomicOr(flags, F_TOOLONG); l = 0; }
  len[r] = (int32_t)l;
  padded[r] = (l + 15) & ~15ll;
}

// Tight source strings -> 16-byte aligned destination rows padded with code 3 (kernels.cuh); the last pad byte carries the
// pad count in its high nibble.  One warp per batch of 32 consecutive rows [row0, row1): the lanes read the rows' metadata
// coalesced, then every lane moves 16-byte chunks of the batch (a warp scan of the chunk counts maps chunk -> row); the
// unaligned source is read as aligned 32-bit words and assembled with funnel shifts.
__global__ void k_repack(const uint8_t* __restrict__ blob, const int64_t* __restrict__ soff,
                         const int32_t* __restrict__ templid, int64_t row0, int64_t row1, const int64_t* __restrict__ doff,
                         const int32_t* __restrict__ len, uint8_t* __restrict__ dst) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_batches = (row1 - row0 + 31) >> 5;
  for (int64_t b = warp; b < n_batches; b += n_warps) {
    const int64_t r = row0 + b * 32 + lane;
    long long s = 0, d = 0;
    int l = 0;
    if (r < row1) {
      const int64_t t = templid ? templid[r] : r;
      s = soff[t]; l = len[r]; d = doff[r];
    }
    const int nch = (l + 15) >> 4;
    int inc = nch;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
      const int o = __shfl_up_sync(FULL, inc, k);
      if (lane >= k) inc += o;
    }
    const int ex = inc - nch;
    const int total = __shfl_sync(FULL, inc, 31);
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      int j = 0;  // the row of chunk c: the last lane whose exclusive chunk count is <= c
#pragma unroll
      for (int step = 16; step; step >>= 1) {
        const int e = __shfl_sync(FULL, ex, j + step);
        if (e <= c) j += step;
      }
      const long long sj = __shfl_sync(FULL, s, j), dj = __shfl_sync(FULL, d, j);
      const int lj = __shfl_sync(FULL, l, j), ej = __shfl_sync(FULL, ex, j);
      if (c < total) {
        const int b0 = (c - ej) * 16;
        const int m = min(16, lj - b0);
        const uintptr_t p = reinterpret_cast<uintptr_t>(blob + sj + b0);
        const uint32_t sh = (uint32_t)(p & 3) * 8u;
        const uint32_t* a = reinterpret_cast<const uint32_t*>(p & ~(uintptr_t)3);
        const int need = (int)(p & 3) + m;  // bytes of the aligned words that belong to this chunk
        const uint32_t w0 = a[0], w1 = need > 4 ? a[1] : 0u, w2 = need > 8 ? a[2] : 0u, w3 = need > 12 ? a[3] : 0u,
                       w4 = need > 16 ? a[4] : 0u;
        uint32_t o[4] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh)};
        if (m < 16) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int v = m - 4 * k;
            if (v < 4) {
              const uint32_t keep = v <= 0 ? 0u : (1u << (8 * v)) - 1u;
              o[k] = (o[k] & keep) | (0x03030303u & ~keep);
            }
          }
          o[3] |= (uint32_t)(16 - m) << 28;
        }
        *reinterpret_cast<uint4*>(dst + dj + b0) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// padded rows -> tight strings (ea_to_host; not on the hot path)
__global__ void k_unpad(const uint8_t* __restrict__ data, const int64_t* __restrict__ doff, const int32_t* __restrict__ len,
                        const int64_t* __restrict__ toff, int64_t n, uint8_t* __restrict__ out) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += n_warps) {
    const uint8_t* s = data + doff[r];
    uint8_t* d = out + toff[r];
    const int l = len[r];
    for (int i = lane; i < l; i += 32) d[i] = s[i];
  }
}

__global__ void k_strand_u8(const int32_t* __restrict__ s32, int64_t n, uint8_t* __restrict__ s8, unsigned* flags) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int v = s32[r];
  if (v != 1 && v != 2) atomicOr(flags, F_STRAND);
  s8[r] = (uint8_t)v;
}

__global__ void k_len_i64(const int32_t* __restrict__ len, int64_t n, int64_t* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) out[r] = len[r];
}

// validation + row-end keys.  key = rname<<32 | (start+len) (0 for empty rows, so they never extend coverage)
__global__ void k_row_keys(Csr csr, uint64_t* __restrict__ key, uint8_t* __restrict__ sp, int32_t* __restrict__ max_len, unsigned* flags) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= csr.n) return;
  const int32_t rn = csr.rname[r], st = csr.start[r], l = csr.len[r];
  unsigned f = 0;
  if (st < 0 || rn < 0) f |= F_NEGSTART;
  if ((int64_t)st + l > 0x7FFFFFF0ll) f |= F_TOOLONG;
  if (csr.strand[r] != 1 && csr.strand[r] != 2) f |= F_STRAND;
  if (r > 0) {
    const int32_t prn = csr.rname[r - 1], pst = csr.start[r - 1];
    if (prn > rn || (prn == rn && pst > st)) f |= F_UNSORTED;
  }
  if (((l + 15) >> 4) & 1) f |= F_ODDCHUNKS;
  if (f) atomicOr(flags, f);
  key[r] = ((uint64_t)(uint32_t)rn << 32) | (uint32_t)(l > 0 ? st + l : 0);
  sp[r] = (uint8_t)((csr.strand[r] == 2 ? 1 : 0) | (((16 - (l & 15)) & 15) << 4));
  atomicMax(max_len, l);
}

// every low nibble of the caller's strings must be one of {2,5,6,7,10,11,12,13,14,15}
// (runs on the tight source blob, 16 bytes per thread step; head/tail bytes one by one)
__global__ void k_validate_domain(const uint8_t* __restrict__ blob, int64_t n_bytes, unsigned* flags) {
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t head = min(n_bytes, (int64_t)((16 - (reinterpret_cast<uintptr_t>(blob) & 15)) & 15));
  const int64_t n_chunks = (n_bytes - head) / 16;
  bool bad = false;
  for (int64_t i = i0; i < n_chunks; i += stride) {
    const uint4 v = *reinterpret_cast<const uint4*>(blob + head + i * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int b = 0; b < 4; b++) bad |= !((0xFCE4u >> ((w[k] >> (8 * b)) & 15u)) & 1u);
  }
  if (i0 < head) bad |= !((0xFCE4u >> (blob[i0] & 15u)) & 1u);
  const int64_t tail0 = head + n_chunks * 16;
  if (tail0 + i0 < n_bytes && i0 < 16) bad |= !((0xFCE4u >> (blob[tail0 + i0] & 15u)) & 1u);
  if (bad) atomicOr(flags, F_DOMAIN);
}

// windows newly opened by row r (rows are sorted, so they are the ones beyond the furthest end so far)
__device__ __forceinline__ void new_windows(const Csr& csr, const uint64_t* maxend, int64_t r, int S, int& first, int& count) {
  const int32_t l = csr.len[r];
  first = 0; count = 0;
  if (l <= 0) return;
  const int32_t st = csr.start[r];
  const int fw = st / S, lw = (st + l - 1) / S;
  int prev_lw = -1;
  if (r > 0) {
    const uint64_t m = maxend[r - 1];
    const int32_t pe = (int32_t)(uint32_t)m;
    if ((int32_t)(m >> 32) == csr.rname[r] && pe > 0) prev_lw = (pe - 1) / S;
  }
  first = max(fw, prev_lw + 1);
  count = max(0, lw - first + 1);
}

__global__ void k_count_windows(Csr csr, const uint64_t* __restrict__ maxend, int S, int32_t* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= csr.n) return;
  int first, count;
  new_windows(csr, maxend, r, S, first, count);
  cnt[r] = count;
}

__global__ void k_fill_windows(Csr csr, const uint64_t* __restrict__ maxend, int S, const int32_t* __restrict__ woff,
                               int32_t* __restrict__ w_rname, int32_t* __restrict__ w_pos0) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= csr.n) return;
  int first, count;
  new_windows(csr, maxend, r, S, first, count);
  const int32_t base = woff[r], rn = csr.rname[r];
  for (int k = 0; k < count; k++) {
    w_rname[base + k] = rn;
    w_pos0[base + k] = (first + k) * S;
  }
}

__global__ void k_window_rows(Csr csr, const uint64_t* __restrict__ maxend, int S, int32_t n_win,
                              const int32_t* __restrict__ w_rname, const int32_t* __restrict__ w_pos0,
                              int32_t* __restrict__ w_lo, int32_t* __restrict__ w_hi, int32_t* max_rows) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_win) return;
  const uint64_t rn = (uint64_t)(uint32_t)w_rname[w] << 32;
  const uint32_t p0 = (uint32_t)w_pos0[w];
  // hi: first row with (rname,start) >= (rn, p0+S)
  const uint64_t khi = rn | (uint64_t)(p0 + (uint32_t)S);
  int64_t a = 0, b = csr.n;
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    const uint64_t k = ((uint64_t)(uint32_t)csr.rname[m] << 32) | (uint32_t)csr.start[m];
    if (k < khi) a = m + 1; else b = m;
  }
  const int64_t hi = a;
  // lo: first row whose running furthest end (same chromosome) reaches beyond p0
  const uint64_t klo = rn | (uint64_t)p0;
  a = 0; b = hi;
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if (maxend[m] <= klo) a = m + 1; else b = m;
  }
  w_lo[w] = (int32_t)a;
  w_hi[w] = (int32_t)hi;
  atomicMax(max_rows, (int32_t)(hi - a));
}

struct MaxU64 {
  __host__ __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; }
};

int grid_for(int64_t n, int threads) { return (int)std::min<int64_t>((n + threads - 1) / threads, 1 << 30); }

int flags_error(unsigned f) {
  if (f & F_TEMPLID) return fail(EA_ERR_ARG, "templid out of range (the reference's seqxm->at() would throw std::out_of_range)");
  if (f & F_TOOLONG) return fail(EA_ERR_ARG, "string length / end position does not fit 31 bits");
  if (f & F_NEGSTART) return fail(EA_ERR_ARG, "negative start or rname");
  if (f & F_STRAND) return fail(EA_ERR_ARG, "strand must be 1 ('+') or 2 ('-')");
  if (f & F_UNSORTED) return fail(EA_ERR_UNSORTED, "rows are not sorted by (rname, start): pre-sorted input is a requirement (rcpp_cx_report.cpp:19)");
  if (f & F_DOMAIN) return fail(EA_ERR_DOMAIN, "a string holds a methylation-call code the reference's readers never emit (valid low nibbles: 2,5,6,7,10..15)");
  return EA_OK;
}

// Common tail of every constructor: validation, row-end keys, stream/events.
int finish_handle(ea_handle* h) {
  CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& e : h->ev) CU(cudaEventCreate(&e));
  int rc;
  if ((rc = dmalloc(&h->d_counters, 4))) return rc;
  if ((rc = dmalloc(&h->d_maxend, (size_t)h->n))) return rc;
  if ((rc = dmalloc(&h->d_sp, (size_t)h->n))) return rc;
  unsigned* d_flags; int32_t* d_maxlen;
  StreamTmp tmps(h->stream);
  if ((rc = dmalloc(&d_flags, 1))) return rc;
  tmps.keep(d_flags);
  if ((rc = dmalloc(&d_maxlen, 1))) return rc;
  tmps.keep(d_maxlen);
  CU(cudaMemsetAsync(d_flags, 0, 4, h->stream));
  CU(cudaMemsetAsync(d_maxlen, 0, 4, h->stream));
  if (h->n > 0) {
    k_row_keys<<<grid_for(h->n, 256), 256, 0, h->stream>>>(csr_of(h), h->d_maxend, h->d_sp, d_maxlen, d_flags);
    size_t tb = 0;
    CU(cub::DeviceScan::InclusiveScan(nullptr, tb, h->d_maxend, h->d_maxend, MaxU64(), (int)h->n, h->stream));
    void* tmp; CU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
    tmps.keep(tmp);
    CU(cub::DeviceScan::InclusiveScan(tmp, tb, h->d_maxend, h->d_maxend, MaxU64(), (int)h->n, h->stream));
  }
  unsigned f = 0;
  CU(cudaMemcpyAsync(&f, d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(&h->max_len, d_maxlen, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  h->rows32 = !(f & F_ODDCHUNKS);
  return flags_error(f & ~(unsigned)F_ODDCHUNKS);
}

int build_index(ea_handle* h, int S, WinIndex** out) {
  WinIndex* ix = nullptr;
  for (auto& c : h->idx) if (c.S == S) { *out = &c; return EA_OK; }
  for (auto& c : h->idx) if (c.S == 0) { ix = &c; break; }
  if (!ix) ix = &h->idx[3];
  *out = ix;
  ix->release();
  ix->S = S;
  if (h->n == 0) return EA_OK;
  int rc;
  int32_t *d_cnt, *d_woff, *d_max;
  StreamTmp tmps(h->stream);
  if ((rc = dmalloc(&d_cnt, (size_t)h->n))) return rc;
  tmps.keep(d_cnt);
  if ((rc = dmalloc(&d_woff, (size_t)h->n))) return rc;
  tmps.keep(d_woff);
  if ((rc = dmalloc(&d_max, 1))) return rc;
  tmps.keep(d_max);
  k_count_windows<<<grid_for(h->n, 256), 256, 0, h->stream>>>(csr_of(h), h->d_maxend, S, d_cnt);
  size_t tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt, d_woff, (int)h->n, h->stream));
  void* tmp; CU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
  tmps.keep(tmp);
  CU(cub::DeviceScan::ExclusiveSum(tmp, tb, d_cnt, d_woff, (int)h->n, h->stream));
  int32_t last_off = 0, last_cnt = 0;
  CU(cudaMemcpyAsync(&last_off, d_woff + h->n - 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(&last_cnt, d_cnt + h->n - 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  ix->n = last_off + last_cnt;
  if ((rc = dmalloc(&ix->rname, (size_t)ix->n))) return rc;
  if ((rc = dmalloc(&ix->pos0, (size_t)ix->n))) return rc;
  if ((rc = dmalloc(&ix->lo, (size_t)ix->n))) return rc;
  if ((rc = dmalloc(&ix->hi, (size_t)ix->n))) return rc;
  if ((rc = dmalloc(&ix->w_base, (size_t)ix->n))) return rc;
  if ((rc = dmalloc(&ix->w_dst, (size_t)ix->n))) return rc;
  if ((rc = dmalloc(&ix->w_cnt, (size_t)ix->n))) return rc;
  CU(cudaMemsetAsync(d_max, 0, 4, h->stream));
  if (ix->n > 0) {
    k_fill_windows<<<grid_for(h->n, 256), 256, 0, h->stream>>>(csr_of(h), h->d_maxend, S, d_woff, ix->rname, ix->pos0);
    k_window_rows<<<grid_for(ix->n, 256), 256, 0, h->stream>>>(csr_of(h), h->d_maxend, S, ix->n, ix->rname, ix->pos0, ix->lo, ix->hi, d_max);
  }
  CU(cudaMemcpyAsync(&ix->max_rows, d_max, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  return EA_OK;
}

__global__ void k_unpack_ls(const uint32_t* __restrict__ ls, int64_t n, int32_t* __restrict__ len, uint8_t* __restrict__ strand) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t v = ls[r];
  len[r] = (int32_t)(v & 0x7FFFFFFFu);
  strand[r] = (uint8_t)(1u + (v >> 31));
}

// call planes that came from the host (ea_pack_blob_calls): every position inside a row must hold one of the codes the
// reference's readers emit {2,5,6,7,10..15}, every position beyond the row's end code 0
__global__ void k_validate_planes(Planar pl, int64_t n_tiles, int th, unsigned* flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  bool bad = false;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    const int64_t u0 = pl.toff[t];
    const int nu_t = (int)(pl.toff[t + 1] - u0);
    if (th == 32) {
      const int64_t r = t * 32 + lane;
      const int len = r < pl.n ? (int)(pl.ls[r] & 0x7FFFFFFFu) : 0;
      for (int j = 0; j < nu_t; j++) {
        const uint4 v = ldg_plunit(pl.units + (u0 + j) * 32 + lane);
        const int nv = len - 32 * j;
        const uint32_t in = nv >= 32 ? 0xFFFFFFFFu : (nv <= 0 ? 0u : (1u << nv) - 1u);
        const uint32_t valid = (v.w & (v.z | v.y)) | (~v.w & ((v.z & (v.x | v.y)) | (~v.z & v.y & ~v.x)));
        bad |= ((~valid & in) | ((v.x | v.y | v.z | v.w) & ~in)) != 0u;
      }
    } else {
      const int len = (int)(pl.ls[t] & 0x7FFFFFFFu);
      for (int j = lane; j < nu_t; j += 32) {
        const uint4 v = ldg_plunit(pl.units + u0 + j);
        const int nv = len - 32 * j;
        const uint32_t in = nv >= 32 ? 0xFFFFFFFFu : (nv <= 0 ? 0u : (1u << nv) - 1u);
        const uint32_t valid = (v.w & (v.z | v.y)) | (~v.w & ((v.z & (v.x | v.y)) | (~v.z & v.y & ~v.x)));
        bad |= ((~valid & in) | ((v.x | v.y | v.z | v.w) & ~in)) != 0u;
      }
    }
  }
  if (bad) atomicOr(flags, (unsigned)F_DOMAIN);
}

#define NEED_BYTES(h)                                                                                                   \
  do {                                                                                                                  \
    if (!(h)->has_bytes)                                                                                                \
      return fail(EA_ERR_NEEDS_BYTES, "this handle holds the call planes only (ea_pack_blob_calls): pack with ea_pack_blob for this call"); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// planar call planes (planar.cuh): built once per handle from the padded byte rows
// ---------------------------------------------------------------------------------------------
Planar planar_of(const ea_handle* h) {
  Planar p;
  p.units = h->d_pl_units; p.toff = h->d_pl_toff; p.ls = h->d_pl_ls; p.start = h->d_start; p.n = h->n;
  return p;
}

bool planar_wanted(const ea_handle* h) {
  static const bool off = getenv("EA_NO_PLANAR") != nullptr;
  return !off && h->n > 0 && h->max_len >= 1;
}

// Tile table (unit rows per tile, prefix sum) and ls[] on stream s; needs d_len / d_strand.  Synchronizes s.
int planar_tiles(ea_handle* h, cudaStream_t s) {
  int rc;
  h->pl_th = h->max_len <= 1024 ? 32 : 1;
  h->pl_tiles = h->pl_th == 32 ? (h->n + 31) >> 5 : h->n;
  int64_t* d_nu = nullptr;
  if ((rc = dmalloc(&d_nu, (size_t)h->pl_tiles + 1))) return rc;
  if ((rc = dmalloc(&h->d_pl_toff, (size_t)h->pl_tiles + 1))) { pool_free_on(d_nu, s); return rc; }
  if ((rc = dmalloc(&h->d_pl_ls, (size_t)h->n))) { pool_free_on(d_nu, s); return rc; }
  CU(cudaMemsetAsync(d_nu, 0, 8 * (size_t)(h->pl_tiles + 1), s));
  if (h->pl_th == 32) k_pl_tile_units<<<grid_for(h->n, 256), 256, 0, s>>>(h->d_len, h->d_strand, h->n, d_nu, h->d_pl_ls);
  else k_pl_row_units<<<grid_for(h->n, 256), 256, 0, s>>>(h->d_len, h->d_strand, h->n, d_nu, h->d_pl_ls);
  size_t tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_nu, h->d_pl_toff, (int)(h->pl_tiles + 1), s));
  void* tmp = nullptr;
  CU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
  CU(cub::DeviceScan::ExclusiveSum(tmp, tb, d_nu, h->d_pl_toff, (int)(h->pl_tiles + 1), s));
  CU(cudaMemcpyAsync(&h->pl_unit_rows, h->d_pl_toff + h->pl_tiles, 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  pool_free_on(tmp, s); pool_free_on(d_nu, s);
  return EA_OK;
}

// transpose the rows of tiles [t0, t1) on stream s (their padded byte rows must be complete in stream order)
int planar_fill(ea_handle* h, int64_t t0, int64_t t1, cudaStream_t s) {
  if (t1 <= t0) return EA_OK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((t1 - t0 + 7) / 8, (int64_t)h->sm_count * 16));
  if (h->pl_th == 32) k_planarize<<<grid, 256, 0, s>>>(h->d_data, h->d_off, h->d_len, h->n, h->d_pl_toff, t0, t1, h->d_pl_units);
  else k_planarize_rows<<<grid, 256, 0, s>>>(h->d_data, h->d_off, h->d_len, h->d_pl_toff, t0, t1, h->d_pl_units);
  CU(cudaGetLastError());
  return EA_OK;
}

// whole handle at once (device-resident constructors)
int build_planar(ea_handle* h) {
  if (h->pl_ok || !planar_wanted(h)) return EA_OK;
  int rc;
  if ((rc = planar_tiles(h, h->stream))) return rc;
  if ((rc = dmalloc(&h->d_pl_units, (size_t)h->pl_unit_rows * (h->pl_th == 32 ? 32 : 1)))) return rc;
  if ((rc = planar_fill(h, 0, h->pl_tiles, h->stream))) return rc;
  CU(cudaStreamSynchronize(h->stream));
  h->pl_ok = true;
  return EA_OK;
}

// ---------------------------------------------------------------------------------------------
// letter sets
// ---------------------------------------------------------------------------------------------
inline unsigned code_of(unsigned char c) { return ((c + 2u) >> 2) & 15u; }  // src/epialleleR.h:28

struct ThresholdSpec {
  int tctx = TC_NONE;  // named context, or -1 for the generic kernel
  LetterWeights lw;
};

void parse_threshold_letters(const char* m, const char* u, const char* om, const char* ou, ThresholdSpec* s) {
  const char* sets[4] = {m ? m : "", u ? u : "", om ? om : "", ou ? ou : ""};
  memset(&s->lw, 0, sizeof s->lw);
  unsigned mask[4] = {0, 0, 0, 0};
  bool multi = false;
  for (int k = 0; k < 4; k++)
    for (const unsigned char* p = (const unsigned char*)sets[k]; *p; p++) {
      const unsigned c = code_of(*p);
      if (s->lw.w[k][c] == 255) continue;
      if (s->lw.w[k][c]++) multi = true;
      mask[k] |= 1u << c;
    }
  s->tctx = -1;
  if (multi) return;
  const unsigned Z = 1u << 7, X = 1u << 6, H = 1u << 2, z = 1u << 15, x = 1u << 14, hh = 1u << 10;
  struct { int t; unsigned m, u, om, ou; } named[5] = {
      {TC_CG, Z, z, X | H, x | hh}, {TC_CHG, X, x, Z | H, z | hh}, {TC_CHH, H, hh, Z | X, z | x},
      {TC_CXG, Z | X, z | x, H, hh}, {TC_CX, Z | X | H, z | x | hh, 0, 0}};
  for (auto& nm : named)
    if (mask[0] == nm.m && mask[1] == nm.u && mask[2] == nm.om && mask[3] == nm.ou) s->tctx = nm.t;
}

// beta only looks at the in-context sets; out-of-context ones of the named context are harmless
void parse_beta_letters(const char* m, const char* u, ThresholdSpec* s) {
  parse_threshold_letters(m, u, "", "", s);
  if (s->tctx >= 0) return;
  static const char* oo[5][2] = {{"XH", "xh"}, {"ZH", "zh"}, {"ZX", "zx"}, {"H", "h"}, {"", ""}};
  for (int k = 0; k < 5; k++) {
    ThresholdSpec t;
    parse_threshold_letters(m, u, oo[k][0], oo[k][1], &t);
    if (t.tctx >= 0) { *s = t; return; }
  }
}

int want_of(const char* ctx) {
  int w = 0;
  for (const unsigned char* p = (const unsigned char*)(ctx ? ctx : ""); *p; p++) {
    const unsigned c = code_of(*p);
    if (c == 2) w |= W_CHH;
    if (c == 6) w |= W_CHG;
    if (c == 7) w |= W_CG;
  }
  return w;
}

int group_lanes(const ea_handle* h) {
  const int nch = (h->max_len + 15) / 16;
  return std::max(1, std::min(32, nch));
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
int thr_tables(ea_handle* h, const ThrParams& thr, ThrTables* tb);

template <int TCTX>
int launch_threshold_t(ea_handle* h, const ThrParams& tp, uint8_t* pass, double* beta, unsigned long long* n_pass) {
  int rc;
  // planar call planes (planar.cuh): a warp per tile of 32 rows, 512 contiguous bytes per load
  if ((g_kernel_policy == 0 || g_kernel_policy == 6) && h->pl_ok) {
    ThrTables tb{nullptr, nullptr};
    if (pass && (rc = thr_tables(h, tp, &tb))) return rc;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((h->pl_tiles + 7) / 8, (int64_t)h->sm_count * 32));
    if (h->pl_th == 32) k_threshold_pl<TCTX><<<blocks, 256, 0, h->stream>>>(planar_of(h), h->pl_tiles, tp, tb, pass, beta, n_pass);
    else k_threshold_plr<TCTX><<<blocks, 256, 0, h->stream>>>(planar_of(h), tp, tb, pass, beta, n_pass);
    return EA_OK;
  }
  // short reads: row per lane (threshold_row.cuh); the integer tables cover counts up to 1024
  if (g_kernel_policy != 1 && h->max_len <= 1024 && h->npadded / std::max<int64_t>(h->n, 1) <= 512) {
    ThrTables tb{nullptr, nullptr};
    if (pass && (rc = thr_tables(h, tp, &tb))) return rc;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((h->n + 255) / 256, (int64_t)h->sm_count * 48));
    if (h->rows32) k_threshold_row<TCTX, true><<<blocks, 256, 0, h->stream>>>(csr_of(h), tp, tb, pass, beta, n_pass);
    else k_threshold_row<TCTX, false><<<blocks, 256, 0, h->stream>>>(csr_of(h), tp, tb, pass, beta, n_pass);
    return EA_OK;
  }
  const int g = group_lanes(h);
  const int rpw = 32 / g;
  const int64_t warps_needed = (h->n + rpw - 1) / rpw;
  const int blocks = (int)std::min<int64_t>((warps_needed + 7) / 8, (int64_t)h->sm_count * 16);
  k_threshold<TCTX><<<std::max(blocks, 1), 256, 0, h->stream>>>(csr_of(h), tp, g, pass, beta, n_pass);
  return EA_OK;
}

int launch_threshold(ea_handle* h, const ThresholdSpec& spec, const ThrParams& tp, uint8_t* pass, double* beta,
                     unsigned long long* n_pass) {
  if (h->n == 0) return EA_OK;
  int rc = EA_OK;
  const bool planar_ok = h->pl_ok && (g_kernel_policy == 0 || g_kernel_policy == 6);
  if (!(planar_ok && spec.tctx >= 0)) NEED_BYTES(h);   // the generic letter-set kernel and the byte-row kernels read the byte rows
  if (spec.tctx < 0 || (h->max_len > 65535 && !(h->pl_ok && (g_kernel_policy == 0 || g_kernel_policy == 6)))) {
    k_threshold_generic<<<grid_for(h->n, 256), 256, 0, h->stream>>>(csr_of(h), spec.lw, tp, pass, beta, n_pass);
  } else {
    switch (spec.tctx) {
      case TC_CG: rc = launch_threshold_t<TC_CG>(h, tp, pass, beta, n_pass); break;
      case TC_CHG: rc = launch_threshold_t<TC_CHG>(h, tp, pass, beta, n_pass); break;
      case TC_CHH: rc = launch_threshold_t<TC_CHH>(h, tp, pass, beta, n_pass); break;
      case TC_CXG: rc = launch_threshold_t<TC_CXG>(h, tp, pass, beta, n_pass); break;
      default: rc = launch_threshold_t<TC_CX>(h, tp, pass, beta, n_pass); break;
    }
  }
  if (rc) return rc;
  CU(cudaGetLastError());
  return EA_OK;
}

template <int WANT>
size_t window_smem(int S, bool wide) {
  return sizeof(int32_t) * 2 * (S + 4) + (wide ? 8 : 4) * (size_t)n_ctx<WANT>() * 2 * S;
}

template <int TCTX, int WANT, typename CT, bool FUSED>
int launch_window_t(ea_handle* h, const WinIndex& ix, const ThrParams& tp, const uint8_t* pass_in, uint8_t* pass_out,
                    const ReportOut& out) {
  const size_t smem = window_smem<WANT>(ix.S, sizeof(CT) == 8);
  auto kern = k_window<TCTX, WANT, CT, FUSED>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 64));
  kern<<<grid, kThreads, smem, h->stream>>>(csr_of(h), wt, tp, pass_in, pass_out, group_lanes(h), out);
  CU(cudaGetLastError());
  return EA_OK;
}

template <int WANT>
int launch_window_w(ea_handle* h, const WinIndex& ix, int tctx_fused, bool wide, const ThrParams& tp,
                    const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  typedef unsigned long long u64;
  if (tctx_fused < 0) {
    return wide ? launch_window_t<TC_NONE, WANT, u64, false>(h, ix, tp, pass_in, pass_out, out)
                : launch_window_t<TC_NONE, WANT, uint32_t, false>(h, ix, tp, pass_in, pass_out, out);
  }
  switch (tctx_fused) {
    case TC_CG: return launch_window_t<TC_CG, WANT, uint32_t, true>(h, ix, tp, pass_in, pass_out, out);
#ifndef EA_FAST_BUILD
    case TC_CHG: return launch_window_t<TC_CHG, WANT, uint32_t, true>(h, ix, tp, pass_in, pass_out, out);
    case TC_CHH: return launch_window_t<TC_CHH, WANT, uint32_t, true>(h, ix, tp, pass_in, pass_out, out);
    case TC_CXG: return launch_window_t<TC_CXG, WANT, uint32_t, true>(h, ix, tp, pass_in, pass_out, out);
    case TC_CX: return launch_window_t<TC_CX, WANT, uint32_t, true>(h, ix, tp, pass_in, pass_out, out);
#endif
    default: return launch_window_t<TC_CG, WANT, uint32_t, true>(h, ix, tp, pass_in, pass_out, out);
  }
}

int launch_window(ea_handle* h, const WinIndex& ix, int want, int tctx_fused, bool wide, const ThrParams& tp,
                  const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_window_w<1>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
    case 2: return launch_window_w<2>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
    case 3: return launch_window_w<3>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
    case 5: return launch_window_w<5>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
    case 6: return launch_window_w<6>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_window_w<7>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
#endif
#ifndef EA_FAST_BUILD
#endif
    default: return launch_window_w<4>(h, ix, tctx_fused, wide, tp, pass_in, pass_out, out);
  }
}

// ---- thread-per-row TMA kernel (short reads, moderate depth) ----
struct TprPlan { bool ok = false; int tile_bytes = 0; };

TprPlan plan_tpr(const ea_handle* h, const WinIndex& ix) {
  TprPlan t;
  if (h->n == 0 || ix.n == 0) return t;
  const int max_pad = (h->max_len + 15) & ~15;
  const int64_t avg_pad = (h->npadded + h->n - 1) / h->n;
  if (max_pad > 512 || ix.max_rows > 32767 || ix.S != 2048) return t;  // long reads / very deep windows: other kernels
  int64_t tb = (32 * avg_pad + 127) & ~127ll;
  tb = std::min<int64_t>(std::max<int64_t>({tb, 4096, (int64_t)((max_pad + 127) & ~127)}), 8192);
  if (tb * 32 < 8 * avg_pad * 32) {}  // (tiles hold >= 8 rows on average by construction of the bounds above)
  if (avg_pad * 8 > tb) return t;      // fewer than 8 rows per tile: thread-per-row would idle most lanes
  t.ok = true; t.tile_bytes = (int)tb;
  return t;
}

int thr_tables(ea_handle* h, const ThrParams& thr, ThrTables* tb) {
  int rc;
  if (!h->d_thr_tables && (rc = dmalloc(&h->d_thr_tables, (size_t)2 * kThrTable))) return rc;
  tb->mmin = h->d_thr_tables; tb->omax = h->d_thr_tables + kThrTable;
  k_thr_tables<<<(kThrTable + 127) / 128, 128, 0, h->stream>>>(thr, h->d_thr_tables, h->d_thr_tables + kThrTable);
  return EA_OK;
}

template <int TCTX, int WANT, bool FUSED>
int launch_tpr_t(ea_handle* h, const WinIndex& ix, const TprPlan& tp, const ThrParams& thr, const uint8_t* pass_in,
                 uint8_t* pass_out, const ReportOut& out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int S = 2048;
  if (ix.S != S) return fail(EA_ERR_STATE, "window size mismatch");
  const size_t smem = sizeof(uint32_t) * (size_t)(NC + 1) * 2 * (kGuardLo + S + kGuardHi) + (size_t)kTprWarps * tp.tile_bytes;
  auto kern = k_window_tpr<TCTX, WANT, FUSED, S>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  ThrTables tb{nullptr, nullptr};
  int rc;
  if (FUSED && (rc = thr_tables(h, thr, &tb))) return rc;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 96));
  kern<<<grid, kTprThreads, smem, h->stream>>>(csr_of(h), wt, thr, tb, pass_in, pass_out, tp.tile_bytes, out);
  CU(cudaGetLastError());
  return EA_OK;
}

template <int WANT>
int launch_tpr_w(ea_handle* h, const WinIndex& ix, const TprPlan& tp, int tctx_fused, const ThrParams& thr,
                 const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (tctx_fused) {
    case TC_CG: return launch_tpr_t<TC_CG, WANT, true>(h, ix, tp, thr, pass_in, pass_out, out);
#ifndef EA_FAST_BUILD
    case TC_CHG: return launch_tpr_t<TC_CHG, WANT, true>(h, ix, tp, thr, pass_in, pass_out, out);
    case TC_CHH: return launch_tpr_t<TC_CHH, WANT, true>(h, ix, tp, thr, pass_in, pass_out, out);
    case TC_CXG: return launch_tpr_t<TC_CXG, WANT, true>(h, ix, tp, thr, pass_in, pass_out, out);
    case TC_CX: return launch_tpr_t<TC_CX, WANT, true>(h, ix, tp, thr, pass_in, pass_out, out);
#endif
    default: return launch_tpr_t<TC_NONE, WANT, false>(h, ix, tp, thr, pass_in, pass_out, out);
  }
}

int launch_tpr(ea_handle* h, const WinIndex& ix, const TprPlan& tp, int want, int tctx_fused, const ThrParams& thr,
               const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_tpr_w<1>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
    case 2: return launch_tpr_w<2>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
    case 3: return launch_tpr_w<3>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
    case 5: return launch_tpr_w<5>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
    case 6: return launch_tpr_w<6>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
    case 7: return launch_tpr_w<7>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
#endif
    default: return launch_tpr_w<4>(h, ix, tp, tctx_fused, thr, pass_in, pass_out, out);
  }
}

// ---- chunk-per-lane kernel (reads up to 1 KB, moderate depth): window_cpl.cuh ----
struct CplPlan { bool ok = false; int ub = 16; int rpt = 0; int tc = 0; size_t smem = 0; };

CplPlan plan_cpl(const ea_handle* h, const WinIndex& ix, int nc, bool fused) {
  CplPlan t;
  if (h->n == 0 || ix.n == 0 || ix.S != 2048) return t;
  const int max_nch = (h->max_len + 15) >> 4;
  if (max_nch < 1 || max_nch > 64 || ix.max_rows > 32767) return t;  // long reads / very deep windows: lane-group kernel
#ifdef EA_CPL_FORCE_UB16
  t.ub = 16;
#else
  t.ub = h->rows32 ? 32 : 16;  // unit a lane handles: a pair of chunks when every padded row is a multiple of 32 bytes
#endif
  const int cpu = t.ub / 16;
  const int max_nu = (max_nch + cpu - 1) / cpu;
  const int avg_nu = (int)((h->npadded / t.ub + h->n - 1) / h->n);
#ifndef EA_CPL_TCMAX
#define EA_CPL_TCMAX 320   // measured: 320 keeps two CTAs per SM for three-context (CX) reports (cfg 3: 7.2 -> 4.7 ms), 160 starves the unit phase
#endif
  int tc = std::min(EA_CPL_TCMAX / cpu, std::max(320 / cpu, 32 * std::max(avg_nu, 1)));  // units per tile
  tc = std::max(tc, max_nu);
  tc = (tc + 31) & ~31;
  t.rpt = std::min(32, tc / max_nu);
  t.tc = tc;
  const int mw = fused ? nc : 0;
  t.smem = sizeof(uint32_t) * ((size_t)(nc + 1) * 2 * (kGuardLo + 2048 + kGuardHi) + (size_t)kCplWarps * (1 + mw) * tc);
  t.ok = t.smem <= 200u * 1024u;
  return t;
}

template <int TCTX, int WANT, bool FUSED, int UB>
int launch_cpl_u(ea_handle* h, const WinIndex& ix, const CplPlan& cp, const ThrParams& thr, const uint8_t* pass_in,
                 uint8_t* pass_out, const ReportOut& out) {
  constexpr int S = 2048;
  if (ix.S != S) return fail(EA_ERR_STATE, "window size mismatch");
  auto kern = k_window_cpl<TCTX, WANT, FUSED, S, UB>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp.smem));
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  ThrTables tb{nullptr, nullptr};
  int rc;
  if (FUSED && (rc = thr_tables(h, thr, &tb))) return rc;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 64));
  kern<<<grid, kCplThreads, cp.smem, h->stream>>>(csr_of(h), wt, thr, tb, pass_in, pass_out, cp.rpt, cp.tc, out);
  CU(cudaGetLastError());
  return EA_OK;
}

template <int TCTX, int WANT, bool FUSED>
int launch_cpl_t(ea_handle* h, const WinIndex& ix, const CplPlan& cp, const ThrParams& thr, const uint8_t* pass_in,
                 uint8_t* pass_out, const ReportOut& out) {
  if (cp.ub == 32) return launch_cpl_u<TCTX, WANT, FUSED, 32>(h, ix, cp, thr, pass_in, pass_out, out);
  return launch_cpl_u<TCTX, WANT, FUSED, 16>(h, ix, cp, thr, pass_in, pass_out, out);
}

template <int WANT>
int launch_cpl_w(ea_handle* h, const WinIndex& ix, const CplPlan& cp, int tctx_fused, const ThrParams& thr,
                 const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (tctx_fused) {
    case TC_CG: return launch_cpl_t<TC_CG, WANT, true>(h, ix, cp, thr, pass_in, pass_out, out);
#ifndef EA_FAST_BUILD
    case TC_CHG: return launch_cpl_t<TC_CHG, WANT, true>(h, ix, cp, thr, pass_in, pass_out, out);
    case TC_CHH: return launch_cpl_t<TC_CHH, WANT, true>(h, ix, cp, thr, pass_in, pass_out, out);
    case TC_CXG: return launch_cpl_t<TC_CXG, WANT, true>(h, ix, cp, thr, pass_in, pass_out, out);
    case TC_CX: return launch_cpl_t<TC_CX, WANT, true>(h, ix, cp, thr, pass_in, pass_out, out);
#endif
    default: return launch_cpl_t<TC_NONE, WANT, false>(h, ix, cp, thr, pass_in, pass_out, out);
  }
}

int launch_cpl(ea_handle* h, const WinIndex& ix, const CplPlan& cp, int want, int tctx_fused, const ThrParams& thr,
               const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_cpl_w<1>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
    case 2: return launch_cpl_w<2>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
    case 3: return launch_cpl_w<3>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
    case 5: return launch_cpl_w<5>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
    case 6: return launch_cpl_w<6>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_cpl_w<7>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
#endif
#ifndef EA_FAST_BUILD
#endif
    default: return launch_cpl_w<4>(h, ix, cp, tctx_fused, thr, pass_in, pass_out, out);
  }
}

// ---- row-per-lane kernel (one requested context, 32-byte-multiple rows of at most 256 bytes): window_row.cuh ----

struct RowPlan { bool ok = false; int maxu = 0; int S = 0; size_t smem = 0; };

// Window size of the row kernel.  A window's rows are dealt to 8 warps in tiles of 32, so the sweep takes
// ceil(rows / 256) rounds and the last one is only partly filled; S is chosen per data set so that a window holds
// close to a whole number of rounds (k * 256 rows including the halo rows that start before the window).
int pick_row_window(const ea_handle* h, const WinIndex& ix2048) {
#ifdef EA_ROW_S
  return EA_ROW_S;
#else
  if (ix2048.n <= 0 || h->n <= 0) return 2048;
  const double dens = (double)h->n / ((double)ix2048.n * 2048.0);        // rows per covered position
  const double halo = ((double)h->nbytes / (double)h->n) * dens;          // rows that start before a window and reach into it
  if (!(dens > 0)) return 2048;
  int best = 2048; double best_eff = 0;
  for (int k = 1; k <= 8; k++) {
    const double s_exact = (256.0 * k - halo) / dens;
    int S = (int)(s_exact / 32.0) * 32;                                   // a little under k full rounds
    if (S < 1024 || S > 2560) continue;                                   // 5 CTAs/SM: (S + 80) * 16 B <= 42 KB
    const double rows = S * dens + halo;
    const double eff = rows / (256.0 * k) * (1.0 - halo / rows * 0.5);    // round filling, mild preference for less halo
    if (eff > best_eff) { best_eff = eff; best = S; }
  }
  return best;
#endif
}

RowPlan plan_row(const ea_handle* h, const WinIndex& ix, int nc) {
  RowPlan t;
  if (h->n == 0 || ix.n == 0 || nc != 1 || !h->rows32 || ix.max_rows > 32767 || ix.S > kRowMaxS || (ix.S & 31)) return t;
  const int max_nu = (h->max_len + 31) >> 5;
  if (max_nu < 1 || max_nu > 8) return t;
  t.maxu = max_nu <= 5 ? 5 : 8;
  t.S = ix.S;
  t.smem = sizeof(uint32_t) * (size_t)(nc + 1) * 2 * (kGuardLo + ix.S + kGuardHi);
  t.ok = true;
  return t;
}

template <int TCTX, int WANT, bool FUSED, int MAXU>
int launch_row_u(ea_handle* h, const WinIndex& ix, const RowPlan& rp, const ThrParams& thr, const uint8_t* pass_in,
                 uint8_t* pass_out, const ReportOut& out) {
  if (ix.S != rp.S) return fail(EA_ERR_STATE, "window size mismatch");
  auto kern = k_window_row<TCTX, WANT, FUSED, MAXU>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp.smem));
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  ThrTables tb{nullptr, nullptr};
  int rc;
  if (FUSED && (rc = thr_tables(h, thr, &tb))) return rc;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 64));
  kern<<<grid, kRowThreads, rp.smem, h->stream>>>(csr_of(h), wt, thr, tb, pass_in, pass_out, out);
  CU(cudaGetLastError());
  return EA_OK;
}

template <int TCTX, int WANT, bool FUSED>
int launch_row_t(ea_handle* h, const WinIndex& ix, const RowPlan& rp, const ThrParams& thr, const uint8_t* pass_in,
                 uint8_t* pass_out, const ReportOut& out) {
#ifndef EA_FAST_BUILD
  if (rp.maxu == 8) return launch_row_u<TCTX, WANT, FUSED, 8>(h, ix, rp, thr, pass_in, pass_out, out);
#endif
  return launch_row_u<TCTX, WANT, FUSED, 5>(h, ix, rp, thr, pass_in, pass_out, out);
}

template <int WANT>
int launch_row_w(ea_handle* h, const WinIndex& ix, const RowPlan& rp, int tctx_fused, const ThrParams& thr,
                 const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (tctx_fused) {
    case TC_CG: return launch_row_t<TC_CG, WANT, true>(h, ix, rp, thr, pass_in, pass_out, out);
#ifndef EA_FAST_BUILD
    case TC_CHG: return launch_row_t<TC_CHG, WANT, true>(h, ix, rp, thr, pass_in, pass_out, out);
    case TC_CHH: return launch_row_t<TC_CHH, WANT, true>(h, ix, rp, thr, pass_in, pass_out, out);
    case TC_CXG: return launch_row_t<TC_CXG, WANT, true>(h, ix, rp, thr, pass_in, pass_out, out);
    case TC_CX: return launch_row_t<TC_CX, WANT, true>(h, ix, rp, thr, pass_in, pass_out, out);
#endif
    default: return launch_row_t<TC_NONE, WANT, false>(h, ix, rp, thr, pass_in, pass_out, out);
  }
}

int launch_row(ea_handle* h, const WinIndex& ix, const RowPlan& rp, int want, int tctx_fused, const ThrParams& thr,
               const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_row_w<1>(h, ix, rp, tctx_fused, thr, pass_in, pass_out, out);
    case 2: return launch_row_w<2>(h, ix, rp, tctx_fused, thr, pass_in, pass_out, out);
#endif
    default: return launch_row_w<4>(h, ix, rp, tctx_fused, thr, pass_in, pass_out, out);
  }
}

// ---- long-read kernel (window_long.cuh): pass flags are an input ----
struct LongPlan { bool ok = false; int S = 0; size_t smem = 0; };

LongPlan plan_long(const ea_handle* h, const WinIndex& ix, int nc) {
  LongPlan t;
  if (h->n == 0 || ix.n == 0 || nc < 1 || nc > 3 || ix.max_rows > 32767 || ix.S > kLongMaxS || (ix.S & 31)) return t;
  t.S = ix.S;
  t.smem = sizeof(uint32_t) * (size_t)(nc + 1) * 2 * (kGuardLo + ix.S + kGuardHi);
  t.ok = true;
  return t;
}

template <int WANT>
int launch_long_w(ea_handle* h, const WinIndex& ix, const LongPlan& lp, const uint8_t* pass_in, const ReportOut& out) {
  if (ix.S != lp.S) return fail(EA_ERR_STATE, "window size mismatch");
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 64));
  if (h->rows32) {
    auto kern = k_window_long<WANT, true>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
    kern<<<grid, kRowThreads, lp.smem, h->stream>>>(csr_of(h), wt, pass_in, out);
  } else {
    auto kern = k_window_long<WANT, false>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
    kern<<<grid, kRowThreads, lp.smem, h->stream>>>(csr_of(h), wt, pass_in, out);
  }
  CU(cudaGetLastError());
  return EA_OK;
}

int launch_long(ea_handle* h, const WinIndex& ix, const LongPlan& lp, int want, const uint8_t* pass_in, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_long_w<1>(h, ix, lp, pass_in, out);
    case 2: return launch_long_w<2>(h, ix, lp, pass_in, out);
    case 3: return launch_long_w<3>(h, ix, lp, pass_in, out);
    case 5: return launch_long_w<5>(h, ix, lp, pass_in, out);
    case 6: return launch_long_w<6>(h, ix, lp, pass_in, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_long_w<7>(h, ix, lp, pass_in, out);
#endif
    default: return launch_long_w<4>(h, ix, lp, pass_in, out);
  }
}

// ---- window kernel on the planar call planes (planar.cuh) ----
struct PlPlan { bool ok = false; int maxu = 0; int S = 0; size_t smem = 0; };

// Window size: as many positions as keep EA_PL_MINB CTAs (one context) resident; more contexts take more state per position.
int pick_pl_window(const ea_handle* h, const WinIndex& ix2048, int nc) {
  if (const char* e = getenv("EA_PL_S")) { const int v = atoi(e) & ~31; if (v >= 256 && v <= kRowMaxS) return v; }
  if (h->pl_th == 1) return nc == 1 ? 3072 : 2048;   // long reads: few rows per window, larger windows amortise the finalize
  if (nc == 1) return pick_row_window(h, ix2048);
  return 2048;
}

PlPlan plan_pl(const ea_handle* h, const WinIndex& ix, int nc) {
  PlPlan t;
  if (h->pl_ok && h->pl_th == 1) {   // long reads: k_window_plr
    if (h->n == 0 || ix.n == 0 || nc < 1 || ix.max_rows > 32767 || ix.S > kRowMaxS || (ix.S & 31)) return t;
    t.maxu = -1; t.S = ix.S;
    t.smem = sizeof(uint32_t) * (size_t)(nc + 1) * 2 * (kGuardLo + ix.S + kGuardHi);
    t.ok = true;
    return t;
  }
  if (!h->pl_ok || h->n == 0 || ix.n == 0 || nc < 1 || ix.max_rows > 32767 || ix.S > kRowMaxS || (ix.S & 31)) return t;
  const int max_nu = (h->max_len + 31) >> 5;
  t.maxu = max_nu <= 5 ? 5 : 0;
  t.S = ix.S;
  t.smem = sizeof(uint32_t) * (size_t)(nc + 1) * 2 * (kGuardLo + ix.S + kGuardHi);
  t.ok = true;
  return t;
}

template <int TCTX, int WANT, bool FUSED, int MAXU, int PER>
int launch_pl_u(ea_handle* h, const WinIndex& ix, const PlPlan& pp, const ThrParams& thr, const uint8_t* pass_in,
                uint8_t* pass_out, const ReportOut& out) {
  if (ix.S != pp.S || ix.S > PER * kRowThreads) return fail(EA_ERR_STATE, "window size mismatch");
  auto kern = k_window_pl<TCTX, WANT, FUSED, MAXU, PER>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  ThrTables tb{nullptr, nullptr};
  int rc;
  if (FUSED && (rc = thr_tables(h, thr, &tb))) return rc;
  static const int grid_mult = []() { const char* e = getenv("EA_PL_GRID"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 64; }();
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * grid_mult));
  kern<<<grid, kRowThreads, pp.smem, h->stream>>>(planar_of(h), wt, thr, tb, pass_in, pass_out, out);
  CU(cudaGetLastError());
  return EA_OK;
}

template <int TCTX, int WANT, bool FUSED>
int launch_pl_t(ea_handle* h, const WinIndex& ix, const PlPlan& pp, const ThrParams& thr, const uint8_t* pass_in,
                uint8_t* pass_out, const ReportOut& out) {
  // positions per thread in the finalize: 8 for windows of up to 2048 positions, 12 up to 3072
  const bool per8 = ix.S <= 8 * kRowThreads;
  if (pp.maxu == 5) {
    if (per8) return launch_pl_u<TCTX, WANT, FUSED, 5, 8>(h, ix, pp, thr, pass_in, pass_out, out);
    return launch_pl_u<TCTX, WANT, FUSED, 5, 12>(h, ix, pp, thr, pass_in, pass_out, out);
  }
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
  if (per8) return launch_pl_u<TCTX, WANT, FUSED, 0, 8>(h, ix, pp, thr, pass_in, pass_out, out);
  return launch_pl_u<TCTX, WANT, FUSED, 0, 12>(h, ix, pp, thr, pass_in, pass_out, out);
#else
  return fail(EA_ERR_STATE, "fast build: rows of more than 160 positions are not compiled in");
#endif
}

template <int WANT>
int launch_pl_w(ea_handle* h, const WinIndex& ix, const PlPlan& pp, int tctx_fused, const ThrParams& thr,
                const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (tctx_fused) {
    case TC_CG: return launch_pl_t<TC_CG, WANT, true>(h, ix, pp, thr, pass_in, pass_out, out);
#ifndef EA_FAST_BUILD
    case TC_CHG: return launch_pl_t<TC_CHG, WANT, true>(h, ix, pp, thr, pass_in, pass_out, out);
    case TC_CHH: return launch_pl_t<TC_CHH, WANT, true>(h, ix, pp, thr, pass_in, pass_out, out);
    case TC_CXG: return launch_pl_t<TC_CXG, WANT, true>(h, ix, pp, thr, pass_in, pass_out, out);
    case TC_CX: return launch_pl_t<TC_CX, WANT, true>(h, ix, pp, thr, pass_in, pass_out, out);
#endif
    default: return launch_pl_t<TC_NONE, WANT, false>(h, ix, pp, thr, pass_in, pass_out, out);
  }
}

template <int WANT>
int launch_plr_w(ea_handle* h, const WinIndex& ix, const PlPlan& pp, const uint8_t* pass_in, const ReportOut& out) {
  if (ix.S != pp.S) return fail(EA_ERR_STATE, "window size mismatch");
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 64));
  if (ix.S <= 8 * kRowThreads) {
    auto kern = k_window_plr<WANT, 8>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
    kern<<<grid, kRowThreads, pp.smem, h->stream>>>(planar_of(h), wt, pass_in, out);
  } else {
    auto kern = k_window_plr<WANT, 12>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
    kern<<<grid, kRowThreads, pp.smem, h->stream>>>(planar_of(h), wt, pass_in, out);
  }
  CU(cudaGetLastError());
  return EA_OK;
}

int launch_plr(ea_handle* h, const WinIndex& ix, const PlPlan& pp, int want, const uint8_t* pass_in, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_plr_w<1>(h, ix, pp, pass_in, out);
    case 2: return launch_plr_w<2>(h, ix, pp, pass_in, out);
    case 3: return launch_plr_w<3>(h, ix, pp, pass_in, out);
    case 5: return launch_plr_w<5>(h, ix, pp, pass_in, out);
    case 6: return launch_plr_w<6>(h, ix, pp, pass_in, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_plr_w<7>(h, ix, pp, pass_in, out);
#endif
    default: return launch_plr_w<4>(h, ix, pp, pass_in, out);
  }
}

int launch_pl(ea_handle* h, const WinIndex& ix, const PlPlan& pp, int want, int tctx_fused, const ThrParams& thr,
              const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  if (pp.maxu < 0) return launch_plr(h, ix, pp, want, pass_in, out);
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_pl_w<1>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
    case 2: return launch_pl_w<2>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
    case 3: return launch_pl_w<3>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
    case 5: return launch_pl_w<5>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
    case 6: return launch_pl_w<6>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_pl_w<7>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
#endif
    default: return launch_pl_w<4>(h, ix, pp, tctx_fused, thr, pass_in, pass_out, out);
  }
}

// ---- window kernel with vertical (bit-sliced) counters on the planar call planes (window_vs.cuh) ----
struct VsPlan { bool ok = false; int k = 0; int lognb = 0; int S = 0; size_t smem = 0; };

int pick_vs_window(const ea_handle* h, const WinIndex& ix2048) {
  if (const char* e = getenv("EA_VS_S")) { const int v = atoi(e); if (v == 512 || v == 2048) return v; }
  (void)h;
  return ix2048.max_rows > 511 ? 512 : 2048;   // deep data: smaller windows, more lanes per block
}

VsPlan plan_vs(const ea_handle* h, const WinIndex& ix, int nc) {
  VsPlan t;
  if (!h->pl_ok || h->pl_th != 32 || h->n == 0 || ix.n == 0 || nc < 1 || (ix.S != 2048 && ix.S != 512)) return t;
  t.lognb = ix.S == 2048 ? 6 : 4;
  t.k = ix.S == 2048 ? 9 : 12;                  // planes per counter: 511 / 4095 rows over one window at most
  if (ix.max_rows > (1 << t.k) - 1) return t;
  t.S = ix.S;
  t.smem = (size_t)kVsCapUnitRows * 512 + sizeof(VsRec) * 2 * kVsRecCap;
  t.ok = true;
  return t;
}

template <int TCTX, int WANT, bool FUSED, int K, int LOGNB>
int launch_vs_u(ea_handle* h, const WinIndex& ix, const VsPlan& vp, const ThrParams& thr, const uint8_t* pass_in,
                uint8_t* pass_out, const ReportOut& out) {
  if (ix.S != vp.S || ix.S != (32 << LOGNB)) return fail(EA_ERR_STATE, "window size mismatch");
  auto kern = k_window_vs<TCTX, WANT, FUSED, K, LOGNB>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vp.smem));
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  ThrTables tb{nullptr, nullptr};
  int rc;
  if (FUSED && (rc = thr_tables(h, thr, &tb))) return rc;
  const int grid = std::max(1, std::min<int>(ix.n, h->sm_count * 32));
  kern<<<grid, kVsThreads, vp.smem, h->stream>>>(planar_of(h), wt, thr, tb, pass_in, pass_out, out, h->max_len);
  CU(cudaGetLastError());
  return EA_OK;
}

template <int TCTX, int WANT, bool FUSED>
int launch_vs_t(ea_handle* h, const WinIndex& ix, const VsPlan& vp, const ThrParams& thr, const uint8_t* pass_in,
                uint8_t* pass_out, const ReportOut& out) {
  if (vp.k == 9 && vp.lognb == 6) return launch_vs_u<TCTX, WANT, FUSED, 9, 6>(h, ix, vp, thr, pass_in, pass_out, out);
  if (vp.k == 12 && vp.lognb == 4) return launch_vs_u<TCTX, WANT, FUSED, 12, 4>(h, ix, vp, thr, pass_in, pass_out, out);
  return fail(EA_ERR_STATE, "no vertical-counter kernel for this window shape");
}

template <int WANT>
int launch_vs_w(ea_handle* h, const WinIndex& ix, const VsPlan& vp, int tctx_fused, const ThrParams& thr,
                const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (tctx_fused) {
    case TC_CG: return launch_vs_t<TC_CG, WANT, true>(h, ix, vp, thr, pass_in, pass_out, out);
#ifndef EA_FAST_BUILD
    case TC_CHG: return launch_vs_t<TC_CHG, WANT, true>(h, ix, vp, thr, pass_in, pass_out, out);
    case TC_CHH: return launch_vs_t<TC_CHH, WANT, true>(h, ix, vp, thr, pass_in, pass_out, out);
    case TC_CXG: return launch_vs_t<TC_CXG, WANT, true>(h, ix, vp, thr, pass_in, pass_out, out);
    case TC_CX: return launch_vs_t<TC_CX, WANT, true>(h, ix, vp, thr, pass_in, pass_out, out);
#endif
    default: return launch_vs_t<TC_NONE, WANT, false>(h, ix, vp, thr, pass_in, pass_out, out);
  }
}

int launch_vs(ea_handle* h, const WinIndex& ix, const VsPlan& vp, int want, int tctx_fused, const ThrParams& thr,
              const uint8_t* pass_in, uint8_t* pass_out, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_vs_w<1>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
    case 2: return launch_vs_w<2>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
    case 3: return launch_vs_w<3>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
    case 5: return launch_vs_w<5>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
    case 6: return launch_vs_w<6>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_vs_w<7>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
#endif
    default: return launch_vs_w<4>(h, ix, vp, tctx_fused, thr, pass_in, pass_out, out);
  }
}

// ---- long reads: a warp per 1024-position window, vertical counters (window_vs.cuh: k_window_vsr) ----
template <int WANT>
int launch_vsr_w(ea_handle* h, const WinIndex& ix, int k, const uint8_t* pass_in, const ReportOut& out) {
  if (ix.S != kVsrS) return fail(EA_ERR_STATE, "window size mismatch");
  WinTable wt;
  wt.rname = ix.rname; wt.pos0 = ix.pos0; wt.lo = ix.lo; wt.hi = ix.hi; wt.n = ix.n; wt.S = ix.S;
  const int grid = std::max(1, std::min<int>((ix.n + kVsWarps - 1) / kVsWarps, h->sm_count * 64));
  if (k == 6) k_window_vsr<WANT, 6><<<grid, kVsThreads, 0, h->stream>>>(planar_of(h), wt, pass_in, out);
  else k_window_vsr<WANT, 9><<<grid, kVsThreads, 0, h->stream>>>(planar_of(h), wt, pass_in, out);
  CU(cudaGetLastError());
  return EA_OK;
}

int launch_vsr(ea_handle* h, const WinIndex& ix, int k, int want, const uint8_t* pass_in, const ReportOut& out) {
  switch (want) {
#ifndef EA_FAST_BUILD
    case 1: return launch_vsr_w<1>(h, ix, k, pass_in, out);
    case 2: return launch_vsr_w<2>(h, ix, k, pass_in, out);
    case 3: return launch_vsr_w<3>(h, ix, k, pass_in, out);
    case 5: return launch_vsr_w<5>(h, ix, k, pass_in, out);
    case 6: return launch_vsr_w<6>(h, ix, k, pass_in, out);
#endif
#if !defined(EA_FAST_BUILD) || defined(EA_FAST_CX)
    case 7: return launch_vsr_w<7>(h, ix, k, pass_in, out);
#endif
    default: return launch_vsr_w<4>(h, ix, k, pass_in, out);
  }
}

// Ownership filter of a range shard: the report is ordered by (rname,pos), so the rows a shard owns are ONE contiguous run
// [a,b) of it; counters[2] = a, counters[3] = b.  counters[0] = rows written (read on the device: no host round trip).
__global__ void k_owned_rows(const int32_t* __restrict__ rname, const int32_t* __restrict__ pos, long long cap,
                             long long key_lo, long long key_hi, unsigned long long* counters) {
  if (threadIdx.x > 1 || blockIdx.x) return;
  const long long n = min((long long)counters[0], cap);
  const long long key = threadIdx.x ? key_hi : key_lo;
  long long a = 0, b = n;
  while (a < b) {
    const long long m = (a + b) >> 1;
    const long long k = ((long long)rname[m] << 32) | (long long)(uint32_t)pos[m];
    if (k < key) a = m + 1; else b = m;
  }
  counters[2 + threadIdx.x] = (unsigned long long)a;
}

int ensure_pass(ea_handle* h) {
  if (!h->d_pass) return dmalloc(&h->d_pass, (size_t)h->n);
  return EA_OK;
}

// exclusive prefix sum of the windows' row counts: one CTA for small indexes, CUB's multi-block scan beyond (the one-CTA
// kernel takes ~110 us for the 123 k windows of config 2, latency-bound on a single SM)
struct I32ToI64 { __host__ __device__ __forceinline__ long long operator()(const int32_t& v) const { return (long long)v; } };
int scan_windows(ea_handle* h, const WinIndex& ix) {
  if (ix.n <= 4096) {
    k_scan_windows<<<1, 1024, 0, h->stream>>>(ix.w_cnt, ix.w_dst, ix.n);
    return EA_OK;
  }
  auto in = thrust::make_transform_iterator(static_cast<const int32_t*>(ix.w_cnt), I32ToI64());
  size_t tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, ix.w_dst, ix.n, h->stream));
  void* tmp = nullptr;
  CU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
  CU(cub::DeviceScan::ExclusiveSum(tmp, tb, in, ix.w_dst, ix.n, h->stream));
  pool_free_on(tmp, h->stream);
  return EA_OK;
}

int ensure_rows(ea_handle* h, int64_t cap) {
  if (cap <= h->cap_rows) return EA_OK;
  pool_free(h->d_scratch); h->d_scratch = nullptr;
  for (auto& c : h->d_cols) { pool_free(c); c = nullptr; }
  h->cap_rows = 0;
  int rc;
  if ((rc = dmalloc(&h->d_scratch, (size_t)cap))) return rc;
  for (auto& c : h->d_cols)
    if ((rc = dmalloc(&c, (size_t)cap))) return rc;
  h->cap_rows = cap;
  return EA_OK;
}

struct ReportPlan {
  int want = 0;
  bool threshold = false;
  ThresholdSpec spec;
  ThrParams tp{};
  int mode = 0;    // 0 fused, 1 two-pass
  int kernel = 0;  // 0 auto (thread-per-row TMA kernel when the data allow it), 1 force the lane-group kernel
};

// The device pipeline of one report: (threshold) -> window kernel -> window scan -> ordered gather.
// pass_in_dev: device pass flags for the unthresholded/explicit-pass case (may be null = all TRUE).
std::mutex g_report_mu[kMaxDevices];   // one report at a time per DEVICE: the launchers set per-function attributes
                                        // (dynamic shared memory) that two host threads sharing a device would race on
int run_report(ea_handle* h, const ReportPlan& rp, const uint8_t* pass_in_dev, ea_timing* tm) {
  int rc;
  std::lock_guard<std::mutex> device_lock(g_report_mu[std::max(0, std::min(h->device, kMaxDevices - 1))]);
  Trace tr("run_report");
  h->n_rows = -1;
  if (tm) memset(tm, 0, sizeof *tm);
  const int nc = ((rp.want >> 0) & 1) + ((rp.want >> 1) & 1) + ((rp.want >> 2) & 1);
  WinIndex* ix = nullptr;
  TprPlan tpr;
  CplPlan cpl;
  RowPlan row;
  LongPlan lng;
  // kernel policy: 0 automatic (planar call planes when the handle has them, else as policy 5), 1 force the general
  // lane-group kernel, 2 prefer the thread-per-row TMA kernel, 3 prefer the chunk-per-lane kernel, 4 prefer the long-read
  // kernel, 5 automatic over the byte rows only (row-per-lane kernel, else chunk-per-lane kernel, else the lane-group kernel)
  const int policy = rp.kernel == 1 ? 1 : g_kernel_policy;
  PlPlan pln;
  if (rp.want && (rc = build_index(h, 2048, &ix))) return rc;
  if (rp.threshold && (rc = ensure_pass(h))) return rc;
  bool wide = ix && ix->max_rows > 65535;
  const bool can_fuse = rp.threshold && rp.spec.tctx >= 0 && !wide && h->max_len <= 1024 && rp.want;
  bool fused = can_fuse && rp.mode == 0 && policy != 4;  // the long-read kernel takes the flags as an input
  VsPlan vsp;
  if (rp.want && policy == 6 && h->pl_ok) {
    const int S = pick_vs_window(h, *ix);
    WinIndex* ixv = ix;
    if (S != 2048 && (rc = build_index(h, S, &ixv))) return rc;
    vsp = plan_vs(h, *ixv, nc);
    if (vsp.ok) ix = ixv;
  }
  int vsr_k = 0;   // long reads: k_window_vsr when at most 511 rows overlap a 1024-position window
  if (rp.want && (policy == 0 || policy == 6) && h->pl_ok && h->pl_th == 1 && !getenv("EA_NO_VSR")) {
    WinIndex* ixr = nullptr;
    if ((rc = build_index(h, kVsrS, &ixr))) return rc;
    if (ixr->n > 0 && ixr->max_rows <= 511) { vsr_k = ixr->max_rows <= 63 ? 6 : 9; ix = ixr; }
  }
  if (rp.want && !vsr_k && (policy == 0 || (policy == 6 && !vsp.ok)) && h->pl_ok) {
    const int S = pick_pl_window(h, *ix, nc);
    WinIndex* ixp = ix;
    if (S != 2048 && (rc = build_index(h, S, &ixp))) return rc;
    pln = plan_pl(h, *ixp, nc);
    if (pln.ok) ix = ixp;
  }
  const bool byte_auto = ((policy == 0 || policy == 6) && !pln.ok && !vsp.ok && !vsr_k) || policy == 5;
  if (rp.want && !pln.ok && !vsp.ok && !vsr_k) NEED_BYTES(h);   // the byte-row kernels
  if (rp.want && !pln.ok && !vsp.ok && !vsr_k) {
    if (byte_auto && nc == 1 && h->rows32 && ((h->max_len + 31) >> 5) <= 8) {
      const int S = pick_row_window(h, *ix);
      WinIndex* ixr = ix;
      if (S != 2048 && (rc = build_index(h, S, &ixr))) return rc;
      row = plan_row(h, *ixr, nc);
      if (row.ok) ix = ixr;
    }
    if ((byte_auto && !row.ok) || policy == 3) cpl = plan_cpl(h, *ix, nc, fused);
    if (policy == 2) tpr = plan_tpr(h, *ix);
    if ((byte_auto && !row.ok && !cpl.ok && !fused) || policy == 4) {  // long reads: (row, unit block) items
      WinIndex* ixl = nullptr;
      if ((rc = build_index(h, nc == 1 ? EA_LONG_S : 2048, &ixl))) return rc;
      lng = plan_long(h, *ixl, nc);
      if (lng.ok) ix = ixl;
    }
    if (!row.ok && !cpl.ok && !tpr.ok && !lng.ok && nc <= 1) {
      if ((rc = build_index(h, 4096, &ix))) return rc;
      wide = ix->max_rows > 65535;
      if (wide) fused = false;  // 64-bit counters: thresholds run as their own pass
    }
  }
  tr.mark("window index");
  // First guess of the row capacity (40 bytes per row): one report row per ~512 bytes of calls for a CpG report, more for the
  // denser contexts.  The window kernel counts the rows it would have written, so an overflow costs one exact-size retry
  // (a second pass over the data) - far cheaper than reserving the worst case of one row per call up front.
  if (h->cap_rows == 0) {
    const int64_t per = (rp.want & W_CHH) ? 64 : (rp.want & W_CHG) ? 192 : 512;
    if ((rc = ensure_rows(h, h->nbytes / per + 4096))) return rc;
  }
  tr.mark("row buffers");

  for (int attempt = 0; attempt < 2; attempt++) {
    int launches = 0;
    CU(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev[0], h->stream));
    const uint8_t* pass_for_window = pass_in_dev;
    if (rp.threshold) {
      if (fused || h->n == 0) {
        CU(cudaMemsetAsync(h->d_pass, 0, (size_t)std::max<int64_t>(h->n, 1), h->stream));  // empty rows keep FALSE
      } else {
        if ((rc = launch_threshold(h, rp.spec, rp.tp, h->d_pass, nullptr, h->d_counters + 1))) return rc;
        launches++;
      }
      pass_for_window = h->d_pass;
    }
    CU(cudaEventRecord(h->ev[1], h->stream));
    POLL();
    if (rp.want && ix->n > 0) {
      ReportOut out;
      out.scratch = h->d_scratch; out.cap = h->cap_rows; out.cursor = h->d_counters; out.w_base = ix->w_base;
      out.w_cnt = ix->w_cnt; out.n_pass = h->d_counters + 1;
      h->report_kernel = vsr_k ? "k_window_vsr (planar call planes, warp per 1024-position window, vertical bit-sliced counters)" : (pln.ok && pln.maxu < 0) ? "k_window_plr (planar call planes, warp per row, 512-byte coalesced loads)" : vsp.ok ? (vsp.lognb == 6 ? "k_window_vs<K=9,S=2048> (planar call planes, vertical bit-sliced counters)" : "k_window_vs<K=12,S=512> (planar call planes, vertical bit-sliced counters)") : pln.ok ? (pln.maxu ? "k_window_pl<MAXU=5> (planar call planes, row per lane, tile-coalesced 128-bit loads)" : "k_window_pl<any> (planar call planes, row per lane, two sweeps)") : row.ok ? "k_window_row (row per lane, 256-bit loads)" : lng.ok ? "k_window_long (warp per row, lane per 32-byte unit)" : cpl.ok ? (cpl.ub == 32 ? "k_window_cpl<UB=32> (chunk-per-lane, 256-bit loads)" : "k_window_cpl<UB=16> (chunk-per-lane)")
                                : tpr.ok ? "k_window_tpr (thread-per-row over TMA tiles)" : "k_window (lane-group)";
      if (vsr_k) rc = launch_vsr(h, *ix, vsr_k, rp.want, pass_for_window, out);
      else if (vsp.ok) rc = launch_vs(h, *ix, vsp, rp.want, fused ? rp.spec.tctx : -1, rp.tp, pass_for_window, h->d_pass, out);
      else if (pln.ok) rc = launch_pl(h, *ix, pln, rp.want, fused ? rp.spec.tctx : -1, rp.tp, pass_for_window, h->d_pass, out);
      else if (row.ok) rc = launch_row(h, *ix, row, rp.want, fused ? rp.spec.tctx : -1, rp.tp, pass_for_window, h->d_pass, out);
      else if (lng.ok) rc = launch_long(h, *ix, lng, rp.want, pass_for_window, out);
      else if (cpl.ok) rc = launch_cpl(h, *ix, cpl, rp.want, fused ? rp.spec.tctx : -1, rp.tp, pass_for_window, h->d_pass, out);
      else if (tpr.ok) rc = launch_tpr(h, *ix, tpr, rp.want, fused ? rp.spec.tctx : -1, rp.tp, pass_for_window, h->d_pass, out);
      else rc = launch_window(h, *ix, rp.want, fused ? rp.spec.tctx : -1, wide, rp.tp, pass_for_window, h->d_pass, out);
      if (rc) return rc;
      launches += ((vsp.ok || pln.ok || row.ok || cpl.ok || tpr.ok) && fused) ? 2 : 1;  // the short-read kernels are preceded by the threshold-table kernel
      CU(cudaEventRecord(h->ev[2], h->stream));
      ReportCols cols{h->d_cols[0], h->d_cols[1], h->d_cols[2], h->d_cols[3], h->d_cols[4], h->d_cols[5]};
      if ((rc = scan_windows(h, *ix))) return rc;
      launches += ix->n > 4096 ? 1 : 0;   // the multi-block scan is two kernels (CUB: tile-state init + scan)
      {  // windows per warp step: 32 when there are plenty of windows, fewer so that a small index still fills the GPU
        int group = 32;
        while (group > 1 && (int64_t)ix->n / group < (int64_t)h->sm_count * 32) group >>= 1;
        const int warps = (ix->n + group - 1) / group;
        k_gather_rows<<<std::max(1, std::min((warps + 3) / 4, h->sm_count * 16)), 128, 0, h->stream>>>(
            h->d_scratch, ix->w_base, ix->w_dst, ix->w_cnt, ix->rname, ix->n, h->cap_rows, cols, group);
      }
      launches += 2;
      CU(cudaGetLastError());
    } else {
      CU(cudaEventRecord(h->ev[2], h->stream));
    }
    if (h->own_set) {
      k_owned_rows<<<1, 32, 0, h->stream>>>(h->d_cols[0], h->d_cols[2], h->cap_rows, h->own_key_lo, h->own_key_hi, h->d_counters);
      launches++;
    }
    CU(cudaEventRecord(h->ev[3], h->stream));
    unsigned long long counters[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(counters, h->d_counters, sizeof counters, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->report_n_launches = launches;
    if ((int64_t)counters[0] > h->cap_rows) {  // scratch too small: grow to the exact size and redo
      if (attempt == 1) return fail(EA_ERR_STATE, "report row capacity exceeded twice");
      if ((rc = ensure_rows(h, (int64_t)counters[0] + 1024))) return rc;
      continue;
    }
    h->n_rows = (int64_t)counters[0];
    h->row_first = 0;
    if (h->own_set) { h->row_first = (int64_t)counters[2]; h->n_rows = std::max<int64_t>(0, (int64_t)counters[3] - (int64_t)counters[2]); }
    tr.mark("kernels");
    if (tm) {
      CU(cudaEventElapsedTime(&tm->total_ms, h->ev[0], h->ev[3]));
      CU(cudaEventElapsedTime(&tm->kernel_ms[0], h->ev[0], h->ev[1]));
      CU(cudaEventElapsedTime(&tm->kernel_ms[1], h->ev[1], h->ev[2]));
      CU(cudaEventElapsedTime(&tm->kernel_ms[2], h->ev[2], h->ev[3]));
      tm->n_launches = launches;
      tm->n_rows = h->n_rows;
      tm->n_pass = rp.threshold ? (int64_t)counters[1] : h->n;
      tm->alg_bytes = h->nbytes + (rp.threshold ? 14 : 13) * h->n + 24 * h->n_rows;
    }
    return EA_OK;
  }
  return EA_OK;
}

int make_plan_from_params(const ea_report_params* p, int mode, ReportPlan* rp) {
  if (!p) return fail(EA_ERR_ARG, "null params");
  rp->want = want_of(p->report_ctx);
  rp->threshold = p->threshold_reads != 0;
  rp->mode = mode & 1;
  rp->kernel = (mode >> 1) & 1;
  if (rp->threshold) {
    if (!p->ctx_meth || !p->ctx_unmeth) return fail(EA_ERR_ARG, "ctx_meth/ctx_unmeth must be non-null");
    parse_threshold_letters(p->ctx_meth, p->ctx_unmeth, p->ooctx_meth, p->ooctx_unmeth, &rp->spec);
    rp->tp.min_n = p->min_n_ctx; rp->tp.min_beta = p->min_ctx_meth_frac; rp->tp.max_oo = p->max_ooctx_meth_frac;
  }
  return EA_OK;
}

int adopt_and_finish(ea_handle* h, ea_handle** out) {
  int rc = finish_handle(h);
  if (rc == EA_OK) rc = build_planar(h);
  if (rc != EA_OK) { ea_free(h); return rc; }
  *out = h;
  return EA_OK;
}

// Build the padded CSR from a tight device blob (file order) + per-row metadata already on the device.
int pack_from_device_blob(ea_handle* h, const uint8_t* d_blob, const int64_t* d_soff, int64_t nstr,
                          const int32_t* d_templid /* null = identity */) {
  int rc;
  unsigned* d_flags; int64_t* d_padded;
  StreamTmp tmps(nullptr);   // every kernel of this function runs on the default stream
  if ((rc = dmalloc(&d_flags, 1))) return rc;
  tmps.keep(d_flags);
  if ((rc = dmalloc(&d_padded, (size_t)h->n + 1))) return rc;
  tmps.keep(d_padded);
  if ((rc = dmalloc(&h->d_len, (size_t)h->n))) return rc;
  if ((rc = dmalloc(&h->d_off, (size_t)h->n + 1))) return rc;
  CU(cudaMemset(d_flags, 0, 4));
  CU(cudaMemset(d_padded, 0, sizeof(int64_t) * (size_t)(h->n + 1)));
  if (h->n > 0) k_lens_from_templid<<<grid_for(h->n, 256), 256>>>(d_soff, d_templid, h->n, nstr, h->d_len, d_padded, d_flags);
  {  // domain check on the caller's bytes (all strings, referenced or not)
    int64_t total = 0;
    CU(cudaMemcpy(&total, d_soff + nstr, 8, cudaMemcpyDeviceToHost));
    if (total > 0) k_validate_domain<<<(int)std::min<int64_t>((total / 16 + 256) / 256, 148 * 32), 256>>>(d_blob, total, d_flags);
  }
  size_t tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_padded, h->d_off, (int)(h->n + 1)));
  void* tmp; CU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
  tmps.keep(tmp);
  CU(cub::DeviceScan::ExclusiveSum(tmp, tb, d_padded, h->d_off, (int)(h->n + 1)));
  unsigned f = 0;
  CU(cudaMemcpy(&f, d_flags, 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&h->npadded, h->d_off + h->n, 8, cudaMemcpyDeviceToHost));
  if (f & ~(unsigned)F_NONIDENT) return flags_error(f);   // F_NONIDENT is informational (rows reach their strings through templid)
  // true byte total
  {
    int64_t* d_l64; int64_t* d_sum;
    if ((rc = dmalloc(&d_l64, (size_t)h->n))) return rc;
    tmps.keep(d_l64);
    if ((rc = dmalloc(&d_sum, 1))) return rc;
    tmps.keep(d_sum);
    if (h->n > 0) k_len_i64<<<grid_for(h->n, 256), 256>>>(h->d_len, h->n, d_l64);
    size_t tb2 = 0;
    CU(cub::DeviceReduce::Sum(nullptr, tb2, d_l64, d_sum, (int)h->n));
    void* tmp2; CU(pool_malloc(&tmp2, std::max<size_t>(tb2, 1)));
    tmps.keep(tmp2);
    CU(cub::DeviceReduce::Sum(tmp2, tb2, d_l64, d_sum, (int)h->n));
    CU(cudaMemcpy(&h->nbytes, d_sum, 8, cudaMemcpyDeviceToHost));
  }
  if ((rc = dmalloc(&h->d_data, (size_t)h->npadded + 16))) return rc;
  if (h->n > 0) k_repack<<<std::max(1, std::min(grid_for(h->n, 256), h->sm_count * 8)), 256>>>(d_blob, d_soff, d_templid, 0, h->n, h->d_off, h->d_len, h->d_data);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  return EA_OK;
}

__global__ void k_synth_rows(easynth::Plan plan, int64_t row_begin, int64_t n, int32_t* rname, uint8_t* strand,
                             int32_t* start, int32_t* len, int64_t* padded) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const easynth::RowInfo r = easynth::row_info(plan, row_begin + i);
  rname[i] = r.rname; strand[i] = (uint8_t)r.strand; start[i] = r.start; len[i] = r.len;
  padded[i] = (r.len + 15) & ~15;
}

// one thread per 16-byte chunk of the padded data
__global__ void k_synth_bytes(easynth::Plan plan, int64_t row_begin, int64_t n, const int64_t* __restrict__ off,
                              uint8_t* __restrict__ data, int uniform_chunks) {
  const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_chunks = off[n] >> 4;
  if (chunk >= n_chunks) return;
  int64_t r;
  if (uniform_chunks > 0) {
    r = chunk / uniform_chunks;
  } else {  // last row with off[r] <= chunk*16
    int64_t a = 0, b = n;
    const int64_t target = chunk << 4;
    while (b - a > 1) {
      const int64_t m = (a + b) >> 1;
      if (off[m] <= target) a = m; else b = m;
    }
    r = a;
  }
  const easynth::RowInfo ri = easynth::row_info(plan, row_begin + r);
  const int j0 = (int)((chunk << 4) - off[r]);
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int j = j0 + 4 * k + b;
      uint32_t byte = j < ri.len ? easynth::read_byte(plan, row_begin + r, ri, j) : 0x03u;
      if (4 * k + b == 15 && j >= ri.len) byte |= (uint32_t)(j0 + 16 - ri.len) << 4;  // pad count in the last pad byte
      v |= byte << (8 * b);
    }
    w[k] = v;
  }
  *reinterpret_cast<uint4*>(data + (chunk << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace

namespace {
cudaError_t d2h_result(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  constexpr size_t kPiece = 16u << 20;
  if (bytes == 0) return cudaSuccess;
  bool pageable = false;
  if (bytes >= 2 * kPiece && !getenv("EA_NO_STAGED_D2H")) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, dst) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
  }
  if (!pageable) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
    return e != cudaSuccess ? e : cudaStreamSynchronize(s);
  }
  uint8_t* buf[2] = {static_cast<uint8_t*>(ea_host_alloc(kPiece)), static_cast<uint8_t*>(ea_host_alloc(kPiece))};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaError_t e = (buf[0] && buf[1]) ? cudaSuccess : cudaErrorMemoryAllocation;
  for (auto& x : ev) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming | cudaEventBlockingSync);
  const size_t n_pieces = (bytes + kPiece - 1) / kPiece;
  const int nt = (int)std::max(1u, std::min(std::thread::hardware_concurrency() / 2, 8u));
  auto copy_out = [&](size_t k) {   // piece k: staging block -> destination, nt threads
    const size_t o = k * kPiece, n = std::min(kPiece, bytes - o);
    const uint8_t* from = buf[k & 1];
    uint8_t* to = static_cast<uint8_t*>(dst) + o;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back([=]() { memcpy(to + n * t / nt, from + n * t / nt, n * (t + 1) / nt - n * t / nt); });
    memcpy(to, from, n / nt);
    for (auto& x : th) x.join();
  };
  for (size_t k = 0; k < n_pieces && e == cudaSuccess; k++) {
    const size_t o = k * kPiece, n = std::min(kPiece, bytes - o);
    e = cudaMemcpyAsync(buf[k & 1], static_cast<const uint8_t*>(src) + o, n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaEventRecord(ev[k & 1], s);
    if (k > 0 && e == cudaSuccess) {   // piece k is on its way: copy piece k-1 out meanwhile
      e = cudaEventSynchronize(ev[(k - 1) & 1]);
      if (e == cudaSuccess) copy_out(k - 1);
    }
  }
  if (e == cudaSuccess) e = cudaEventSynchronize(ev[(n_pieces - 1) & 1]);
  if (e == cudaSuccess) copy_out(n_pieces - 1);
  else cudaStreamSynchronize(s);
  for (auto& x : ev) if (x) cudaEventDestroy(x);
  ea_host_free(buf[0]); ea_host_free(buf[1]);
  return e;
}
}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int ea_version(void) { return 100; }
const char* ea_last_error(void) { return g_err.c_str(); }
int ea_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
void ea_set_poll(ea_poll_fn fn) { g_poll = fn; }
void ea_set_handle_poll(ea_handle* h, ea_poll_fn fn) { if (h) h->poll = fn; }
int ea_set_kernel_policy(int policy) { const int old = g_kernel_policy; if (policy >= 0) g_kernel_policy = policy; return old; }

void ea_free(ea_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  Trace tr("ea_free");
  FreeScope fs;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->owns) {
    pool_free(h->d_rname); pool_free(h->d_strand); pool_free(h->d_start); pool_free(h->d_off); pool_free(h->d_data);
  }
  pool_free(h->d_len);
  pool_free(h->d_maxend);
  pool_free(h->d_sp);
  pool_free(h->d_pl_units); pool_free(h->d_pl_toff); pool_free(h->d_pl_ls);
  for (auto& c : h->idx) c.release();
  pool_free(h->d_pass); pool_free(h->d_beta); pool_free(h->d_counters); pool_free(h->d_thr_tables); pool_free(h->d_match); pool_free(h->d_mhl_scratch);
  for (auto& c : h->d_mhl_i) pool_free(c);
  for (auto& c : h->d_mhl_d) pool_free(c); pool_free(h->d_scratch); pool_free(h->d_flush); pool_free(h->d_text);
  for (auto& c : h->d_cols) pool_free(c);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  tr.mark("cudaFree");
}

// Page-locked host blocks for results (ea_host_alloc / ea_host_free): cudaHostAlloc of a few hundred MB costs more than the
// pageable copy it would speed up, so freed blocks are kept in a small cache and handed out again (first fit within 2x).
namespace {
struct HostBlock { void* p; size_t bytes; };
std::mutex g_host_mu;
std::vector<HostBlock> g_host_free;                 // cached, not in use
std::unordered_map<void*, size_t> g_host_live;      // handed out
}  // namespace

void* ea_host_alloc(size_t bytes) {
  if (bytes == 0) bytes = 1;
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    int best = -1;
    for (int i = 0; i < (int)g_host_free.size(); i++)
      if (g_host_free[i].bytes >= bytes && g_host_free[i].bytes <= 2 * bytes + (1u << 20) &&
          (best < 0 || g_host_free[i].bytes < g_host_free[best].bytes)) best = i;
    if (best >= 0) {
      const HostBlock b = g_host_free[best];
      g_host_free.erase(g_host_free.begin() + best);
      g_host_live[b.p] = b.bytes;
      return b.p;
    }
  }
  void* p = nullptr;
  const size_t rounded = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
  if (cudaHostAlloc(&p, rounded, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    fail(EA_ERR_CUDA, "cudaHostAlloc(%zu) failed", rounded);
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_host_mu);
  g_host_live[p] = rounded;
  return p;
}

void ea_host_free(void* p) {
  if (!p) return;
  std::vector<HostBlock> drop;
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_live.find(p);
    if (it == g_host_live.end()) return;
    g_host_free.push_back(HostBlock{p, it->second});
    g_host_live.erase(it);
    // the cache of unused page-locked blocks is bounded (EA_PINNED_CACHE_MB, default 16 GiB): the oldest blocks go first
    static const size_t cap = []() { const char* e = getenv("EA_PINNED_CACHE_MB"); return (size_t)(e ? std::max<long long>(0, atoll(e)) : 16384) << 20; }();
    size_t held = 0;
    for (auto& b : g_host_free) held += b.bytes;
    while (held > cap && !g_host_free.empty()) {
      held -= g_host_free.front().bytes;
      drop.push_back(g_host_free.front());
      g_host_free.erase(g_host_free.begin());
    }
  }
  for (auto& b : drop) cudaFreeHost(b.p);
}

int ea_trim(int device) {
  int rc;
  if ((rc = pick_device(device))) return rc;
  cudaMemPool_t pool;
  CU(cudaDeviceSynchronize());
  CU(ea_pool(device, &pool));
  CU(cudaMemPoolTrimTo(pool, 0));
  std::vector<HostBlock> blocks;
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    blocks.swap(g_host_free);
  }
  for (auto& b : blocks) cudaFreeHost(b.p);
  return EA_OK;
}

int64_t ea_nrow(const ea_handle* h) { return h ? h->n : -1; }
int64_t ea_nbytes(const ea_handle* h) { return h ? h->nbytes : -1; }
int ea_device(const ea_handle* h) { return h ? h->device : -1; }

// Host columns + one tight string blob -> packed handle, as a two-stream pipeline:
//   copy stream    : row metadata, then the blob in chunks of whole strings (one event per chunk);
//   kernel stream  : strand/length/offset kernels as soon as the metadata landed, then per chunk the domain check and -
//                    when templid is the identity (rows are the strings in file order, what preprocessBam produces) - the
//                    repack of that chunk's rows, so only the last chunk's repack is not hidden behind the PCIe transfer;
//   handle stream  : sortedness validation, row-end keys and pad counts (finish_handle) while the blob is still in flight.
// With page-locked caller memory the copies are asynchronous and everything overlaps; with pageable memory the same code
// runs with the copies serialised by the driver.
// fill (optional): called with a range of strings right before their chunk of the blob is sent - ea_pack gathers the
// scattered strings of chunk k into the page-locked blob while chunk k-1 is on its way to the device.
namespace {
// fill (ea_pack): the caller's strings are scattered; fill(s0, s1, dst) gathers strings [s0, s1) back to back into dst, one
// of `kRing` page-locked blocks the implementation cycles through (a block is refilled once its previous copy has landed)
int pack_blob_impl(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                   int64_t nrow, const uint8_t* blob, const int64_t* off, int64_t nstr, int device, ea_handle** out,
                   const std::function<void(int64_t, int64_t, uint8_t*)>* fill);
}
int ea_pack_blob(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                 int64_t nrow, const uint8_t* blob, const int64_t* off, int64_t nstr, int device, ea_handle** out) {
  return pack_blob_impl(rname, strand, start, templid, nrow, blob, off, nstr, device, out, nullptr);
}
namespace {
int pack_blob_impl(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                   int64_t nrow, const uint8_t* blob, const int64_t* off, int64_t nstr, int device, ea_handle** out,
                   const std::function<void(int64_t, int64_t, uint8_t*)>* fill) {
  if (!out) return fail(EA_ERR_ARG, "null out");
  *out = nullptr;
  if (nrow < 0 || nstr < 0 || nrow > 0x7FFFFFFEll) return fail(EA_ERR_ARG, "bad row/string count (at most 2^31-2 rows per handle)");
  if (nrow > 0 && (!rname || !strand || !start || !off)) return fail(EA_ERR_ARG, "null column");
  if (nstr > 0 && !off) return fail(EA_ERR_ARG, "null offsets");
  int rc;
  if ((rc = pick_device(device))) return rc;
  if (nstr > 0 && off[0] < 0) return fail(EA_ERR_ARG, "negative string offset");
  const int64_t base = nstr > 0 ? off[0] : 0;   // string i = blob[off[i], off[i+1]): the offsets need not start at 0 (a shard of a larger blob)
  const int64_t total = nstr > 0 ? off[nstr] - base : 0;
  if (total < 0 || (total > 0 && !blob && !fill)) return fail(EA_ERR_ARG, "bad string blob");
  Trace tr("ea_pack_blob");
  // tight copy + padded rows + call planes + ~80 bytes per row of columns, keys, index scratch and report rows
  if (!getenv("EA_NO_POOL_RESERVE")) pool_reserve((size_t)((double)total * 2.7) + (size_t)nrow * 64 + (64u << 20));
  tr.mark("pool reserve");
  ea_handle* h = new ea_handle;
  h->device = device; h->n = nrow;
  // chunks of whole strings: ~256 MB (~5 ms of PCIe 5 x16) from a caller's blob, at most 32 of them; ~64 MB when the strings
  // are gathered through the ring of page-locked blocks (a first call then pins 200 MB, not the whole data set);
  // EA_PACK_CHUNK_MB overrides (tests)
  const int kMaxChunks = fill ? 4096 : 32;
  constexpr int kRing = 3;
  int64_t chunk_bytes = fill ? 64ll << 20 : 256ll << 20;
  if (const char* e = getenv("EA_PACK_CHUNK_MB")) chunk_bytes = std::max<int64_t>(1, atoll(e)) << 20;
  const int n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)kMaxChunks, total / chunk_bytes + 1, std::max<int64_t>(nstr, 1)}));
  int32_t *d_strand32 = nullptr, *d_templid = nullptr; int64_t *d_soff = nullptr, *d_padded = nullptr; uint8_t* d_blob = nullptr;
  unsigned* d_flags = nullptr; void* d_tmp = nullptr;
  cudaStream_t sc = nullptr, sk = nullptr;
  cudaEvent_t ev_meta = nullptr;
  std::vector<cudaEvent_t> ev_chunk((size_t)n_chunks, nullptr);
  uint8_t* ring[kRing] = {};
  auto cleanup = [&]() {
    for (auto& p : ring) ea_host_free(p);
    FreeScope fs;
    pool_free(d_strand32); pool_free(d_templid); pool_free(d_soff); pool_free(d_blob); pool_free(d_flags); pool_free(d_padded); pool_free(d_tmp);
    if (sc) cudaStreamDestroy(sc);
    if (sk) cudaStreamDestroy(sk);
    if (ev_meta) cudaEventDestroy(ev_meta);
    for (auto& e : ev_chunk) if (e) cudaEventDestroy(e);
  };
#define TRY(x) do { rc = (x); if (rc != EA_OK) { cleanup(); ea_free(h); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ea_free(h); return fail(EA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
  TRYCU(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
  TRYCU(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
  TRYCU(cudaEventCreateWithFlags(&ev_meta, cudaEventDisableTiming));
  for (int k = 0; k < n_chunks; k++) TRYCU(cudaEventCreateWithFlags(&ev_chunk[k], cudaEventDisableTiming));
  TRY(dmalloc(&h->d_rname, (size_t)nrow));
  TRY(dmalloc(&h->d_start, (size_t)nrow));
  TRY(dmalloc(&h->d_strand, (size_t)nrow));
  TRY(dmalloc(&h->d_len, (size_t)nrow));
  TRY(dmalloc(&h->d_off, (size_t)nrow + 1));
  TRY(dmalloc(&d_strand32, (size_t)nrow));
  TRY(dmalloc(&d_padded, (size_t)nrow + 1));
  TRY(dmalloc(&d_soff, (size_t)nstr + 1));
  TRY(dmalloc(&d_blob, (size_t)total + 16));
  TRY(dmalloc(&d_flags, 1));
  if (templid && nrow > 0) TRY(dmalloc(&d_templid, (size_t)nrow));
  size_t tb = 0;
  TRYCU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_padded, h->d_off, (int)(nrow + 1), sk));
  TRYCU(pool_malloc(&d_tmp, std::max<size_t>(tb, 1)));
  tr.mark("cudaMalloc staging");

  // copy stream: metadata, then the blob in chunks of whole strings
  if (nrow > 0) {
    TRYCU(cudaMemcpyAsync(h->d_rname, rname, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sc));
    TRYCU(cudaMemcpyAsync(h->d_start, start, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sc));
    TRYCU(cudaMemcpyAsync(d_strand32, strand, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sc));
    if (d_templid) TRYCU(cudaMemcpyAsync(d_templid, templid, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sc));
  }
  if (nstr > 0) TRYCU(cudaMemcpyAsync(d_soff, off, 8 * (size_t)(nstr + 1), cudaMemcpyHostToDevice, sc));
  else TRYCU(cudaMemsetAsync(d_soff, 0, 8, sc));
  TRYCU(cudaEventRecord(ev_meta, sc));
  // kernel stream: everything that only needs the metadata
  TRYCU(cudaMemsetAsync(d_flags, 0, 4, sk));
  TRYCU(cudaMemsetAsync(d_padded + nrow, 0, 8, sk));
  TRYCU(cudaStreamWaitEvent(sk, ev_meta, 0));
  if (nrow > 0) {
    k_strand_u8<<<grid_for(nrow, 256), 256, 0, sk>>>(d_strand32, nrow, h->d_strand, d_flags);
    k_lens_from_templid<<<grid_for(nrow, 256), 256, 0, sk>>>(d_soff, d_templid, nrow, nstr, h->d_len, d_padded, d_flags);
  }
  TRYCU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_padded, h->d_off, (int)(nrow + 1), sk));
  std::vector<int64_t> s_begin((size_t)n_chunks + 1);  // string index range of every chunk
  for (int k = 0; k <= n_chunks; k++) s_begin[(size_t)k] = nstr * k / n_chunks;
  if (fill) {
    int64_t slot = 1;
    for (int k = 0; k < n_chunks; k++) slot = std::max(slot, off[s_begin[(size_t)k + 1]] - off[s_begin[(size_t)k]]);
    for (auto& p : ring) { p = static_cast<uint8_t*>(ea_host_alloc((size_t)slot)); if (!p) { cleanup(); ea_free(h); return EA_ERR_CUDA; } }
  }
  for (int k = 0; k < n_chunks; k++) {
    const int64_t b0 = nstr > 0 ? off[s_begin[k]] - base : 0, b1 = nstr > 0 ? off[s_begin[k + 1]] - base : 0;
    if (b1 < b0 || b1 > total) { cleanup(); ea_free(h); return fail(EA_ERR_ARG, "string offsets are not ascending"); }
    const uint8_t* src = blob ? blob + base + b0 : nullptr;
    if (fill && b1 > b0) {
      if (k >= kRing) TRYCU(cudaEventSynchronize(ev_chunk[(size_t)(k - kRing)]));   // the block's previous copy has landed
      (*fill)(s_begin[k], s_begin[k + 1], ring[k % kRing]);
      src = ring[k % kRing];
    }
    if (b1 > b0) TRYCU(cudaMemcpyAsync(d_blob + b0, src, (size_t)(b1 - b0), cudaMemcpyHostToDevice, sc));
    TRYCU(cudaEventRecord(ev_chunk[k], sc));
  }
  unsigned f = 0;
  TRYCU(cudaMemcpyAsync(&f, d_flags, 4, cudaMemcpyDeviceToHost, sk));
  TRYCU(cudaMemcpyAsync(&h->npadded, h->d_off + nrow, 8, cudaMemcpyDeviceToHost, sk));
  TRYCU(cudaStreamSynchronize(sk));
  tr.mark("H2D metadata + row kernels");
  if (f & ~(unsigned)F_NONIDENT) { cleanup(); ea_free(h); return flags_error(f); }
  const bool by_chunk = !(f & F_NONIDENT) && nstr >= nrow;  // rows == strings in file order: repack chunk by chunk
  TRY(dmalloc(&h->d_data, (size_t)h->npadded + 16));
  rc = finish_handle(h);  // validation + row keys on the handle's own stream, while the blob is in flight
  if (rc != EA_OK) { cleanup(); ea_free(h); return rc; }
  // the planar call planes (planar.cuh) are filled tile by tile behind the repack of every chunk, like the byte rows
  const bool planar = planar_wanted(h);
  if (planar) {
    TRY(planar_tiles(h, sk));
    TRY(dmalloc(&h->d_pl_units, (size_t)h->pl_unit_rows * (h->pl_th == 32 ? 32 : 1)));
  }
  int64_t tiles_done = 0;
  const int rp_grid = std::max(1, h->sm_count * 8);
  for (int k = 0; k < n_chunks; k++) {
    const int64_t b0 = nstr > 0 ? off[s_begin[k]] - base : 0, b1 = nstr > 0 ? off[s_begin[k + 1]] - base : 0;
    TRYCU(cudaStreamWaitEvent(sk, ev_chunk[k], 0));
    if (b1 > b0) k_validate_domain<<<(int)std::min<int64_t>(((b1 - b0) / 16 + 256) / 256, h->sm_count * 16), 256, 0, sk>>>(d_blob + b0, b1 - b0, d_flags);
    if (by_chunk) {
      const int64_t r0 = std::min(s_begin[k], nrow), r1 = std::min(s_begin[k + 1], nrow);
      if (r1 > r0) k_repack<<<std::min(grid_for(r1 - r0, 256), rp_grid), 256, 0, sk>>>(d_blob - base, d_soff, nullptr, r0, r1, h->d_off, h->d_len, h->d_data);
      if (planar) {  // the tiles whose rows are all repacked by now
        const int64_t t1 = (k == n_chunks - 1 || r1 >= nrow) ? h->pl_tiles : (h->pl_th == 32 ? r1 >> 5 : r1);
        TRY(planar_fill(h, tiles_done, t1, sk));
        tiles_done = std::max(tiles_done, t1);
      }
    }
  }
  if (!by_chunk && nrow > 0) {
    k_repack<<<std::min(grid_for(nrow, 256), rp_grid), 256, 0, sk>>>(d_blob - base, d_soff, d_templid, 0, nrow, h->d_off, h->d_len, h->d_data);
    if (planar) TRY(planar_fill(h, 0, h->pl_tiles, sk));
  }
  TRYCU(cudaMemcpyAsync(&f, d_flags, 4, cudaMemcpyDeviceToHost, sk));
  TRYCU(cudaStreamSynchronize(sk));
  TRYCU(cudaGetLastError());
  tr.mark("H2D blob + repack");
  if (f & ~(unsigned)F_NONIDENT) { cleanup(); ea_free(h); return flags_error(f); }
  if (by_chunk && nstr == nrow) {
    h->nbytes = total;
  } else {  // sum of the referenced strings' lengths
    int64_t* d_l64; int64_t* d_sum;
    TRY(dmalloc(&d_l64, (size_t)nrow));
    TRY(dmalloc(&d_sum, 1));
    if (nrow > 0) k_len_i64<<<grid_for(nrow, 256), 256, 0, sk>>>(h->d_len, nrow, d_l64);
    size_t tb2 = 0;
    TRYCU(cub::DeviceReduce::Sum(nullptr, tb2, d_l64, d_sum, (int)nrow, sk));
    void* tmp2; TRYCU(pool_malloc(&tmp2, std::max<size_t>(tb2, 1)));
    TRYCU(cub::DeviceReduce::Sum(tmp2, tb2, d_l64, d_sum, (int)nrow, sk));
    TRYCU(cudaMemcpyAsync(&h->nbytes, d_sum, 8, cudaMemcpyDeviceToHost, sk));
    TRYCU(cudaStreamSynchronize(sk));
    pool_free(tmp2); pool_free(d_l64); pool_free(d_sum);
  }
  h->pl_ok = planar;   // filled on sk, which has been synchronized above
  h->h2d_bytes = total + 8 * (nstr + 1) + 4 * nrow * (templid ? 4 : 3);
  cleanup();
  tr.mark("free staging");
#undef TRY
#undef TRYCU
  *out = h;
  return EA_OK;
}
}  // namespace

// Calls-only packer: the host transposes the strings into the planar call planes (hostpack.cpp, all cores, AVX2) chunk by
// chunk into a ring of page-locked staging blocks while the previous chunk crosses PCIe, so 4 bits per position travel
// instead of 8 and the device does no repacking at all.  The handle has no byte rows (has_bytes = false): everything on
// the cytosine-report path works on it (thresholds, betas, reports, BED matching), the base-level paths (VCF, lMHL,
// ea_export_host) answer EA_ERR_NEEDS_BYTES.  Falls back to ea_pack_blob when rows are not the strings in file order.
namespace {
// worker threads that live for one pack call: run(fn) executes fn(thread index) on all of them and returns when all are done
class PackPool {
 public:
  explicit PackPool(int n) : n_(n) {
    for (int t = 0; t < n_; t++) th_.emplace_back([this, t]() { loop(t); });
  }
  ~PackPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
    cv_start_.notify_all();
    for (auto& x : th_) x.join();
  }
  void run(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn; pending_ = n_; gen_++;
    cv_start_.notify_all();
    cv_done_.wait(lk, [this]() { return pending_ == 0; });
  }
  // [lo, hi) in pieces of `grain`, handed out dynamically
  void parallel_for(int64_t lo, int64_t hi, int64_t grain, const std::function<void(int64_t, int64_t)>& body) {
    std::atomic<int64_t> next{lo};
    run([&](int) {
      for (;;) {
        const int64_t a = next.fetch_add(grain);
        if (a >= hi) break;
        body(a, std::min(hi, a + grain));
      }
    });
  }

 private:
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_start_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
      }
      (*fn)(t);
      { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) cv_done_.notify_one(); }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};
}  // namespace

int ea_pack_blob_calls(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                       int64_t nrow, const uint8_t* blob, const int64_t* off, int64_t nstr, int device, ea_handle** out) {
  if (!out) return fail(EA_ERR_ARG, "null out");
  *out = nullptr;
  if (nrow != nstr || nrow <= 0) return ea_pack_blob(rname, strand, start, templid, nrow, blob, off, nstr, device, out);
  if (!rname || !strand || !start || !off) return fail(EA_ERR_ARG, "null column");
  if (off[0] < 0) return fail(EA_ERR_ARG, "negative string offset");
  const int64_t total = off[nstr] - off[0];   // the offsets need not start at 0 (a shard of a larger blob)
  if (total < 0 || (total > 0 && !blob)) return fail(EA_ERR_ARG, "bad string blob");
  if (nrow > 0x7FFFFFFEll) return fail(EA_ERR_ARG, "bad row/string count (at most 2^31-2 rows per handle)");
  int rc;
  if ((rc = pick_device(device))) return rc;
  Trace tr("ea_pack_blob_calls");
  // call planes (half the bytes, ragged tiles up to ~1.6x) + four raw chunks + ~80 bytes per row
  if (!getenv("EA_NO_POOL_RESERVE")) pool_reserve((size_t)((double)total * 0.85) + (size_t)nrow * 48 + (832u << 20));
  tr.mark("pool reserve");
  int nt = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  if (g_pack_threads > 0) nt = std::min(nt, g_pack_threads);
  if (const char* e = getenv("EA_HOST_PACK_THREADS")) nt = std::max(1, atoi(e));
  if (nt < 6 && !getenv("EA_HOST_PACK_FORCE"))   // too few cores for the host half to pay: the general packer (all bytes over PCIe)
    return ea_pack_blob(rname, strand, start, templid, nrow, blob, off, nstr, device, out);
  PackPool pool(nt);
  const eahp::Job job{blob, off, strand, templid, nrow};
  const int64_t n_tiles32 = (nrow + 31) >> 5;

  // Chunks of whole 32-row tiles, ~192 MB of strings each.  The chunks are worked off FROM BOTH ENDS: this thread sends
  // chunks from the front as they are (raw bytes, transposed on the device, k_planarize_tight), a second thread lets the
  // pool transpose chunks from the back on the host and sends the planes (half the bytes).  PCIe and the host's cores are
  // both kept busy and the split between them finds itself, whatever the number of cores this process may use.
  int64_t raw_chunk = 192ll << 20;
  if (const char* e = getenv("EA_PACK_CHUNK_MB")) raw_chunk = std::max<int64_t>(1, atoll(e)) << 20;
  const int64_t avg_len = std::max<int64_t>(1, total / nrow);
  const int64_t rows_per_chunk = std::max<int64_t>(32, (raw_chunk / avg_len) & ~31ll);
  const int n_chunks = (int)((nrow + rows_per_chunk - 1) / rows_per_chunk);
  auto chunk_rows = [&](int c, int64_t& r0, int64_t& r1) { r0 = (int64_t)c * rows_per_chunk; r1 = std::min(nrow, r0 + rows_per_chunk); };
  int64_t max_raw = 0;
  for (int c = 0; c < n_chunks; c++) { int64_t r0, r1; chunk_rows(c, r0, r1); max_raw = std::max(max_raw, off[r1] - off[r0]); }

  ea_handle* h = new ea_handle;
  h->device = device; h->n = nrow; h->has_bytes = false; h->nbytes = total; h->npadded = 0;
  constexpr int kSlots = 2;                                // page-locked staging blocks of host-packed chunks
  constexpr int kRawSlots = 4;                             // raw chunks on the device (four keep PCIe busy during the host's first pass)
  cudaStream_t sr = nullptr, sh = nullptr;                 // raw chunks + metadata / host-packed chunks
  cudaEvent_t ev_raw[kRawSlots] = {}, ev_host[kSlots] = {}, ev_meta = nullptr;
  uint8_t* stage[kSlots] = {};
  uint8_t* d_raw[kRawSlots] = {}; int64_t* d_soff[kRawSlots] = {};
  uint32_t* ls = nullptr; int32_t* tile_nu = nullptr; int64_t* toff = nullptr;
  unsigned* d_flags = nullptr;
  auto cleanup = [&]() {
    if (sr) { cudaStreamSynchronize(sr); cudaStreamDestroy(sr); }
    if (sh) { cudaStreamSynchronize(sh); cudaStreamDestroy(sh); }
    for (auto& e : ev_raw) if (e) cudaEventDestroy(e);
    for (auto& e : ev_host) if (e) cudaEventDestroy(e);
    if (ev_meta) cudaEventDestroy(ev_meta);
    for (auto& p : stage) ea_host_free(p);
    ea_host_free(ls); ea_host_free(tile_nu); ea_host_free(toff);
    FreeScope fs;
    for (auto& p : d_raw) pool_free(p);
    for (auto& p : d_soff) pool_free(p);
    pool_free(d_flags);
  };
#define TRY(x) do { rc = (x); if (rc != EA_OK) { cleanup(); ea_free(h); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ea_free(h); return fail(EA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
  TRYCU(cudaStreamCreateWithFlags(&sr, cudaStreamNonBlocking));
  TRYCU(cudaStreamCreateWithFlags(&sh, cudaStreamNonBlocking));
  for (auto& e : ev_raw) TRYCU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
  for (auto& e : ev_host) TRYCU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
  TRYCU(cudaEventCreateWithFlags(&ev_meta, cudaEventDisableTiming | cudaEventBlockingSync));
  TRY(dmalloc(&h->d_rname, (size_t)nrow));
  TRY(dmalloc(&h->d_start, (size_t)nrow));
  for (int i = 0; i < kRawSlots; i++) {
    TRY(dmalloc(&d_raw[i], (size_t)max_raw + 64));
    TRY(dmalloc(&d_soff[i], (size_t)rows_per_chunk + 1));
  }
  // ---- PCIe starts right away: the two columns the device takes as they are, then the first raw chunks ----
  TRYCU(cudaMemcpyAsync(h->d_rname, rname, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sr));
  TRYCU(cudaMemcpyAsync(h->d_start, start, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sr));
  std::mutex qmu;
  int front = 0, back = n_chunks - 1;                      // next chunk from the front (raw) / from the back (host-packed)
  auto claim_front = [&]() { std::lock_guard<std::mutex> lk(qmu); return front <= back ? front++ : -1; };
  auto claim_back = [&]() { std::lock_guard<std::mutex> lk(qmu); return back >= front ? back-- : -1; };
  std::vector<int> raw_chunks;                             // chunks whose bytes are on their way; their kernels follow once toff is known
  auto send_raw = [&](int c, int slot) -> cudaError_t {
    int64_t r0, r1; chunk_rows(c, r0, r1);
    cudaError_t e = cudaSuccess;
    if (off[r1] > off[r0]) e = cudaMemcpyAsync(d_raw[slot], blob + off[r0], (size_t)(off[r1] - off[r0]), cudaMemcpyHostToDevice, sr);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_soff[slot], off + r0, 8 * (size_t)(r1 - r0 + 1), cudaMemcpyHostToDevice, sr);
    return e;
  };
  const int n_early = std::min(kRawSlots, std::max(0, n_chunks - 1));   // (at least one chunk stays for the host side)
  for (int i = 0; i < n_early; i++) { const int c = claim_front(); if (c < 0) break; TRYCU(send_raw(c, (int)raw_chunks.size() % kRawSlots)); raw_chunks.push_back(c); }
  // ---- pass 1 on the host meanwhile: lengths, strands, unit rows per tile ----
  ls = static_cast<uint32_t*>(ea_host_alloc(4 * (size_t)nrow));
  tile_nu = static_cast<int32_t*>(ea_host_alloc(4 * (size_t)n_tiles32));
  if (!ls || !tile_nu) { cleanup(); ea_free(h); return EA_ERR_CUDA; }
  std::atomic<unsigned> flags{0};
  std::atomic<int64_t> true_max{0};
  pool.parallel_for(0, n_tiles32, 4096, [&](int64_t a, int64_t b) {
    flags |= eahp::scan_rows(job, a, b, ls, tile_nu);
    int64_t m = 0;
    for (int64_t r = a * 32; r < std::min(b * 32, nrow); r++) m = std::max<int64_t>(m, ls[r] & 0x7FFFFFFFu);
    int64_t cur = true_max.load();
    while (m > cur && !true_max.compare_exchange_weak(cur, m)) {}
  });
  const int64_t max_l = true_max.load();
  if ((flags & eahp::kNotIdentity) || max_l < 1) {   // rows are not the strings in file order (or there is nothing to pack): the general packer
    cleanup(); ea_free(h);
    return ea_pack_blob(rname, strand, start, templid, nrow, blob, off, nstr, device, out);
  }
  if (flags & eahp::kBadLength) { cleanup(); ea_free(h); return fail(EA_ERR_ARG, "string length / end position does not fit 31 bits"); }
  if (flags & eahp::kBadStrand) { cleanup(); ea_free(h); return fail(EA_ERR_ARG, "strand must be 1 ('+') or 2 ('-')"); }
  const int th = max_l <= 1024 ? 32 : 1;
  const int64_t n_tiles = th == 32 ? n_tiles32 : nrow;
  toff = static_cast<int64_t*>(ea_host_alloc(8 * (size_t)(n_tiles + 1)));
  if (!toff) { cleanup(); ea_free(h); return EA_ERR_CUDA; }
  toff[0] = 0;
  if (th == 32) for (int64_t t = 0; t < n_tiles; t++) toff[t + 1] = toff[t] + tile_nu[t];
  else for (int64_t r = 0; r < nrow; r++) toff[r + 1] = toff[r] + (((ls[r] & 0x7FFFFFFFu) + 31u) >> 5);
  const int64_t unit_bytes = th == 32 ? 512 : 16;     // bytes per toff step
  const int64_t total_units_bytes = toff[n_tiles] * unit_bytes;
  h->pl_th = th; h->pl_tiles = n_tiles; h->pl_unit_rows = toff[n_tiles];
  auto chunk_tiles = [&](int c, int64_t& t0, int64_t& t1) { int64_t r0, r1; chunk_rows(c, r0, r1); t0 = th == 32 ? r0 >> 5 : r0; t1 = th == 32 ? (r1 + 31) >> 5 : r1; };
  int64_t stage_bytes = 0;
  for (int c = 0; c < n_chunks; c++) { int64_t t0, t1; chunk_tiles(c, t0, t1); stage_bytes = std::max(stage_bytes, (toff[t1] - toff[t0]) * unit_bytes); }
  tr.mark("host pass 1");

  TRY(dmalloc(&h->d_strand, (size_t)nrow));
  TRY(dmalloc(&h->d_len, (size_t)nrow));
  TRY(dmalloc(&h->d_pl_ls, (size_t)nrow));
  TRY(dmalloc(&h->d_pl_toff, (size_t)n_tiles + 1));
  TRY(dmalloc(&h->d_pl_units, (size_t)(total_units_bytes / 16)));
  TRY(dmalloc(&d_flags, 1));
  for (auto& p : stage) { p = static_cast<uint8_t*>(ea_host_alloc((size_t)std::max<int64_t>(stage_bytes, 1))); if (!p) { cleanup(); ea_free(h); return EA_ERR_CUDA; } }
  TRYCU(cudaMemsetAsync(d_flags, 0, 4, sh));
  TRYCU(cudaMemcpyAsync(h->d_pl_ls, ls, 4 * (size_t)nrow, cudaMemcpyHostToDevice, sh));
  TRYCU(cudaMemcpyAsync(h->d_pl_toff, toff, 8 * (size_t)(n_tiles + 1), cudaMemcpyHostToDevice, sh));
  k_unpack_ls<<<grid_for(nrow, 256), 256, 0, sh>>>(h->d_pl_ls, nrow, h->d_len, h->d_strand);
  TRYCU(cudaEventRecord(ev_meta, sh));
  TRYCU(cudaStreamWaitEvent(sr, ev_meta, 0));            // the raw chunks' kernels need toff
  int64_t h2d = 12 * nrow + 8 * (n_tiles + 1);

  // ---- host side: chunks from the back, transposed by the pool, planes sent on sh ----
  std::atomic<int64_t> h2d_host{0};
  std::atomic<int> host_err{0};
  std::atomic<bool> abort_pack{false};
  std::thread host_thread([&]() {
    cudaSetDevice(device);
    for (int k = 0; !abort_pack; k++) {
      const int c = claim_back();
      if (c < 0) break;
      const int slot = k % kSlots;
      if (k >= kSlots && cudaEventSynchronize(ev_host[slot]) != cudaSuccess) { host_err = 1; break; }
      int64_t t0, t1; chunk_tiles(c, t0, t1);
      uint8_t* dst = stage[slot];
      const int64_t grain = std::max<int64_t>(1, th == 32 ? 512 : 4096 / std::max<int64_t>(1, (max_l + 31) >> 5));
      pool.parallel_for(t0, t1, grain, [&](int64_t a, int64_t b) {
        uint8_t* d = dst + (toff[a] - toff[t0]) * unit_bytes;
        flags |= th == 32 ? eahp::pack_tiles(job, toff, a, b, d) : eahp::pack_rows(job, toff, a, b, d);
      });
      const int64_t bytes = (toff[t1] - toff[t0]) * unit_bytes;
      if (bytes > 0 && cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->d_pl_units) + toff[t0] * unit_bytes, dst, (size_t)bytes, cudaMemcpyHostToDevice, sh) != cudaSuccess) { host_err = 1; break; }
      if (cudaEventRecord(ev_host[slot], sh) != cudaSuccess) { host_err = 1; break; }
      h2d_host += bytes;
    }
  });
  // ---- this thread: raw chunks from the front ----
  cudaError_t raw_err = cudaSuccess;
  auto launch_tight = [&](int c, int slot) {
    int64_t r0, r1, t0, t1; chunk_rows(c, r0, r1); chunk_tiles(c, t0, t1);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((t1 - t0 + 7) / 8, (int64_t)148 * 16));
    k_planarize_tight<<<grid, 256, 0, sr>>>(d_raw[slot], off[r0], d_soff[slot], r0, nrow, h->d_pl_toff, t0, t1, th, h->d_pl_units);
    h2d += (off[r1] - off[r0]) + 8 * (r1 - r0 + 1);
    return cudaEventRecord(ev_raw[slot], sr);
  };
  for (size_t i = 0; i < raw_chunks.size() && raw_err == cudaSuccess; i++) raw_err = launch_tight(raw_chunks[i], (int)i % kRawSlots);
  // While the early raw chunks and the first host-packed chunk travel: validation + row keys (finish_handle) and the window
  // indexes a one-context report will ask for - they only need the metadata, which is on the device by now.
  int rc_meta = EA_OK;
  if (raw_err == cudaSuccess) raw_err = cudaEventSynchronize(ev_meta);
  if (raw_err == cudaSuccess) {
    rc_meta = finish_handle(h);
    WinIndex* ixp = nullptr;
    if (rc_meta == EA_OK && h->max_len != (int32_t)max_l) rc_meta = fail(EA_ERR_STATE, "length bookkeeping mismatch");
    if (rc_meta == EA_OK) rc_meta = build_index(h, h->pl_th == 32 ? 2048 : kVsrS, &ixp);
    if (rc_meta == EA_OK && h->pl_th == 32) {
      const int S = pick_pl_window(h, *ixp, 1);
      if (S != 2048) rc_meta = build_index(h, S, &ixp);
    }
    if (rc_meta != EA_OK) abort_pack = true;
  }
  for (int k = (int)raw_chunks.size(); raw_err == cudaSuccess && rc_meta == EA_OK; k++) {
    const int slot = k % kRawSlots;
    if ((raw_err = cudaEventSynchronize(ev_raw[slot])) != cudaSuccess) break;   // the chunk that used this slot has been transposed
    const int c = claim_front();
    if (c < 0) break;
    if ((raw_err = send_raw(c, slot)) != cudaSuccess) break;
    raw_err = launch_tight(c, slot);
  }
  host_thread.join();
  if (rc_meta != EA_OK) { const std::string msg = g_err; cleanup(); ea_free(h); g_err = msg; return rc_meta; }
  if (raw_err != cudaSuccess || host_err) { cleanup(); ea_free(h); return fail(EA_ERR_CUDA, "pack pipeline: %s", cudaGetErrorString(raw_err != cudaSuccess ? raw_err : cudaGetLastError())); }
  TRYCU(cudaStreamSynchronize(sr));
  h->h2d_bytes = h2d + h2d_host.load();
  // the device checks what it has: codes inside the domain, nothing beyond the rows' ends (covers both ways in)
  {
    Planar pl; pl.units = h->d_pl_units; pl.toff = h->d_pl_toff; pl.ls = h->d_pl_ls; pl.start = h->d_start; pl.n = nrow;
    k_validate_planes<<<(int)std::max<int64_t>(1, std::min<int64_t>((n_tiles + 7) / 8, 148 * 32)), 256, 0, sh>>>(pl, n_tiles, th, d_flags);
  }
  unsigned dflags = 0;
  TRYCU(cudaMemcpyAsync(&dflags, d_flags, 4, cudaMemcpyDeviceToHost, sh));
  TRYCU(cudaStreamSynchronize(sh));
  TRYCU(cudaGetLastError());
  tr.mark("transposition (host + device) + H2D");
  if ((flags & eahp::kBadCode) || (dflags & F_DOMAIN)) { cleanup(); ea_free(h); return flags_error(F_DOMAIN); }
  h->pl_ok = true;
  cleanup();
#undef TRY
#undef TRYCU
  *out = h;
  return EA_OK;
}

int ea_has_bytes(const ea_handle* h) { return h && h->has_bytes ? 1 : 0; }
int64_t ea_pack_h2d_bytes(const ea_handle* h) { return h ? h->h2d_bytes : -1; }

int ea_pack(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid, int64_t nrow,
            const uint8_t* const* str, const int64_t* len, int64_t nstr, int device, ea_handle** out) {
  if (!out) return fail(EA_ERR_ARG, "null out");
  *out = nullptr;
  if (nstr < 0 || (nstr > 0 && (!str || !len))) return fail(EA_ERR_ARG, "bad string table");
  std::vector<int64_t> off((size_t)nstr + 1, 0);
  for (int64_t i = 0; i < nstr; i++) {
    if (len[i] < 0) return fail(EA_ERR_ARG, "negative string length");
    off[(size_t)i + 1] = off[(size_t)i] + len[i];
  }
  int rc;
  if ((rc = pick_device(device))) return rc;
  // the strings are gathered chunk by chunk into a ring of three page-locked blocks (pack_blob_impl), chunk k while chunk
  // k-1 crosses PCIe: the H2D copies run at PCIe speed and a first call pins ~200 MB instead of the whole data set
  const int nt_max = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::thread::hardware_concurrency(), 16));
  const std::function<void(int64_t, int64_t, uint8_t*)> gather = [&](int64_t s0, int64_t s1, uint8_t* dst) {  // strings [s0, s1) back to back
    const int64_t b0 = off[(size_t)s0], bytes = off[(size_t)s1] - b0;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt_max, bytes / (4 << 20) + 1));
    auto work = [&](int t) {
      for (int64_t i = s0 + (s1 - s0) * t / nt; i < s0 + (s1 - s0) * (t + 1) / nt; i++)
        if (len[i]) memcpy(dst + (off[(size_t)i] - b0), str[i], (size_t)len[i]);
    };
    if (nt == 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  };
  for (int64_t i = 0; i < nstr; i++) if (len[i] > 0 && !str[i]) return fail(EA_ERR_ARG, "null string %lld", (long long)i);
  return pack_blob_impl(rname, strand, start, templid, nrow, nullptr, off.data(), nstr, device, out, &gather);
}

int ea_attach_device(int32_t* d_rname, uint8_t* d_strand, int32_t* d_start, uint8_t* d_data, int64_t* d_off,
                     int64_t nrow, int device, int take_ownership, ea_handle** out) {
  if (!out) return fail(EA_ERR_ARG, "null out");
  *out = nullptr;
  if (nrow < 0 || nrow > 0x7FFFFFFEll) return fail(EA_ERR_ARG, "bad row count (at most 2^31-2 rows per handle)");
  if (nrow > 0 && (!d_rname || !d_strand || !d_start || !d_data || !d_off)) return fail(EA_ERR_ARG, "null device pointer");
  int rc;
  if ((rc = pick_device(device))) return rc;
  ea_handle* h = new ea_handle;
  h->device = device; h->n = nrow;
  // the caller's CSR is tight; rows are re-laid out into the padded form, metadata is copied
  if ((rc = dmalloc(&h->d_rname, (size_t)nrow)) || (rc = dmalloc(&h->d_start, (size_t)nrow)) || (rc = dmalloc(&h->d_strand, (size_t)nrow))) { ea_free(h); return rc; }
  if (nrow > 0) {
    cudaMemcpy(h->d_rname, d_rname, 4 * (size_t)nrow, cudaMemcpyDeviceToDevice);
    cudaMemcpy(h->d_start, d_start, 4 * (size_t)nrow, cudaMemcpyDeviceToDevice);
    cudaMemcpy(h->d_strand, d_strand, (size_t)nrow, cudaMemcpyDeviceToDevice);
  }
  int64_t zero = 0; int64_t* d_soff = d_off;
  if (nrow == 0) { if ((rc = dmalloc(&d_soff, 1))) { ea_free(h); return rc; } cudaMemcpy(d_soff, &zero, 8, cudaMemcpyHostToDevice); }
  rc = pack_from_device_blob(h, d_data, d_soff, nrow, nullptr);
  if (nrow == 0) pool_free(d_soff);
  if (rc != EA_OK) { ea_free(h); return rc; }
  if (take_ownership) { pool_free(d_rname); pool_free(d_strand); pool_free(d_start); pool_free(d_data); pool_free(d_off); }
  return adopt_and_finish(h, out);
}

int ea_threshold_reads(ea_handle* h, const char* ctx_meth, const char* ctx_unmeth, const char* ooctx_meth,
                       const char* ooctx_unmeth, uint32_t min_n_ctx, double min_ctx_meth_frac,
                       double max_ooctx_meth_frac, uint8_t* pass_out) {
  if (!h || !ctx_meth || !ctx_unmeth || (h->n > 0 && !pass_out)) return fail(EA_ERR_ARG, "null argument");
  int rc;
  CU(cudaSetDevice(h->device));
  if ((rc = ensure_pass(h))) return rc;
  ThresholdSpec spec;
  parse_threshold_letters(ctx_meth, ctx_unmeth, ooctx_meth, ooctx_unmeth, &spec);
  ThrParams tp{min_n_ctx, min_ctx_meth_frac, max_ooctx_meth_frac};
  CU(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned long long), h->stream));
  if ((rc = launch_threshold(h, spec, tp, h->d_pass, nullptr, h->d_counters + 1))) return rc;
  if (h->n > 0) CU(d2h_result(pass_out, h->d_pass, (size_t)h->n, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return EA_OK;
}

int ea_get_xm_beta(ea_handle* h, const char* ctx_meth, const char* ctx_unmeth, double* beta_out) {
  if (!h || !ctx_meth || !ctx_unmeth || (h->n > 0 && !beta_out)) return fail(EA_ERR_ARG, "null argument");
  int rc;
  CU(cudaSetDevice(h->device));
  if (!h->d_beta && (rc = dmalloc(&h->d_beta, (size_t)h->n))) return rc;
  ThresholdSpec spec;
  parse_beta_letters(ctx_meth, ctx_unmeth, &spec);
  ThrParams tp{0, 0.0, 0.0};
  if ((rc = launch_threshold(h, spec, tp, nullptr, h->d_beta, nullptr))) return rc;
  if (h->n > 0) CU(d2h_result(beta_out, h->d_beta, 8 * (size_t)h->n, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return EA_OK;
}

// ---- BED-assisted report ------------------------------------------------
namespace {
struct BedDev {
  int64_t *s_key = nullptr, *e_key = nullptr;
  int32_t *s_start = nullptr, *s_end = nullptr, *s_pmax = nullptr, *s_idx = nullptr, *e_idx = nullptr;
  int32_t n = 0;
  void release() { FreeScope fs; pool_free(s_key); pool_free(e_key); pool_free(s_start); pool_free(s_end); pool_free(s_pmax); pool_free(s_idx); pool_free(e_idx); }
};

// Sort the BED rows whose chromosome the BAM knows by (chr, start) and by (chr, end); file order index kept.
int upload_bed(ea_handle* h, const int32_t* chr, const int32_t* start, const int32_t* end, int64_t nbed, BedDev* d) {
  std::vector<int32_t> ord;
  for (int64_t i = 0; i < nbed; i++)
    if (chr[i] != kNaInt) ord.push_back((int32_t)i);
  const size_t n = ord.size();
  d->n = (int32_t)n;
  std::vector<int32_t> oe(ord);
  std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return chr[a] != chr[b] ? chr[a] < chr[b] : start[a] < start[b]; });
  std::stable_sort(oe.begin(), oe.end(), [&](int32_t a, int32_t b) { return chr[a] != chr[b] ? chr[a] < chr[b] : end[a] < end[b]; });
  std::vector<int64_t> sk(n), ek(n);
  std::vector<int32_t> ss(n), se(n), sp(n);
  for (size_t k = 0; k < n; k++) {
    const int32_t i = ord[k];
    sk[k] = bed_key(chr[i], start[i]); ss[k] = start[i]; se[k] = end[i];
    sp[k] = (k > 0 && chr[ord[k - 1]] == chr[i]) ? std::max(sp[k - 1], end[i]) : end[i];
    ek[k] = bed_key(chr[oe[k]], end[oe[k]]);
  }
  int rc;
  if ((rc = dmalloc(&d->s_key, n)) || (rc = dmalloc(&d->e_key, n)) || (rc = dmalloc(&d->s_start, n)) || (rc = dmalloc(&d->s_end, n)) ||
      (rc = dmalloc(&d->s_pmax, n)) || (rc = dmalloc(&d->s_idx, n)) || (rc = dmalloc(&d->e_idx, n)))
    return rc;
  if (n) {
    CU(cudaMemcpyAsync(d->s_key, sk.data(), 8 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d->e_key, ek.data(), 8 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d->s_start, ss.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d->s_end, se.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d->s_pmax, sp.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d->s_idx, ord.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d->e_idx, oe.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));  // the host vectors die with this frame
  }
  return EA_OK;
}

int match_on_device(ea_handle* h, int mode, const int32_t* chr, const int32_t* start, const int32_t* end, int64_t nbed, int32_t param) {
  if (mode != 0 && mode != 1) return fail(EA_ERR_ARG, "mode must be 0 (amplicon) or 1 (capture)");
  if (nbed < 0 || nbed > 0x7FFFFFF0ll || (nbed > 0 && (!chr || !start || !end))) return fail(EA_ERR_ARG, "bad BED table");
  int rc;
  if (!h->d_match && (rc = dmalloc(&h->d_match, (size_t)h->n))) return rc;
  BedDev bd;
  if ((rc = upload_bed(h, chr, start, end, nbed, &bd))) { bd.release(); return rc; }
  BedIndex bi{bd.s_key, bd.s_start, bd.s_end, bd.s_pmax, bd.s_idx, bd.e_key, bd.e_idx, bd.n};
  if (h->n > 0) k_match_target<<<grid_for(h->n, 256), 256, 0, h->stream>>>(csr_of(h), bi, mode, param, h->d_match);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  bd.release();
  return EA_OK;
}
}  // namespace

int ea_match_target(ea_handle* h, int mode, const int32_t* bed_chr, const int32_t* bed_start, const int32_t* bed_end, int64_t nbed,
                    int32_t param, int32_t* match_out) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  int rc;
  CU(cudaSetDevice(h->device));
  if ((rc = match_on_device(h, mode, bed_chr, bed_start, bed_end, nbed, param))) return rc;
  if (match_out && h->n > 0) {
    CU(d2h_result(match_out, h->d_match, 4 * (size_t)h->n, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return EA_OK;
}

// rcpp_get_base_freqs (src/rcpp_get_base_freqs.cpp:15-61): see vcf.cuh.  The VCF entries must be sorted by
// (chromosome code, position), as the reference's cursor walk assumes; entries with an NA chromosome never match.
int ea_base_freqs(ea_handle* h, const uint8_t* pass, const int32_t* vcf_chr, const int32_t* vcf_pos, int64_t n_vcf, double* freqs_out) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  if (n_vcf < 0 || n_vcf > 0x7FFFFFFFll || (n_vcf > 0 && (!vcf_chr || !vcf_pos || !freqs_out))) return fail(EA_ERR_ARG, "bad VCF table");
  if (n_vcf == 0) return EA_OK;
  int rc;
  NEED_BYTES(h);   // base codes are the HIGH nibble
  CU(cudaSetDevice(h->device));
  std::vector<uint64_t> keys, na_key;     // valid entries; per NA entry the key of the last valid entry in front of it
  std::vector<int32_t> idx, na_after;     // row in the caller's table; per NA entry the number of valid entries in front of it
  keys.reserve((size_t)n_vcf); idx.reserve((size_t)n_vcf);
  for (int64_t i = 0; i < n_vcf; i++) {
    if (vcf_chr[i] == INT32_MIN) {
      if (!keys.empty()) { na_key.push_back(keys.back()); na_after.push_back((int32_t)keys.size()); }
      continue;
    }
    if (vcf_chr[i] < 0 || vcf_pos[i] < 0) return fail(EA_ERR_ARG, "negative VCF chromosome code or position");
    const uint64_t k = ((uint64_t)(uint32_t)vcf_chr[i] << 32) | (uint32_t)vcf_pos[i];
    if (!keys.empty() && k < keys.back())
      return fail(EA_ERR_UNSORTED, "VCF entries are not sorted by (seqnames, start): rcpp_get_base_freqs assumes sorted reads and VCF entries (rcpp_get_base_freqs.cpp:6)");
    keys.push_back(k); idx.push_back((int32_t)i);
  }
  const size_t cells = (size_t)n_vcf * 20;
  unsigned int* d_counts = nullptr; double* d_out = nullptr; uint64_t* d_keys = nullptr; int32_t* d_idx = nullptr; uint8_t* d_p = nullptr;
  uint64_t* d_na_key = nullptr; int32_t* d_na_limit = nullptr; int32_t* d_limit = nullptr;
  auto cleanup = [&]() { FreeScope fs; pool_free(d_counts); pool_free(d_out); pool_free(d_keys); pool_free(d_idx); pool_free(d_p); pool_free(d_na_key); pool_free(d_na_limit); pool_free(d_limit); };
#define TRY(x) do { rc = (x); if (rc != EA_OK) { cleanup(); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return fail(EA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
  TRY(dmalloc(&d_counts, cells));
  TRY(dmalloc(&d_out, cells));
  TRY(dmalloc(&d_keys, keys.size()));
  TRY(dmalloc(&d_idx, idx.size()));
  TRY(dmalloc(&d_limit, keys.size()));
  TRYCU(cudaMemsetAsync(d_counts, 0, 4 * cells, h->stream));
  if (!keys.empty() && h->n > 0) {
    // visibility limits (see vcf.cuh): one binary search per NA entry on the prefix maximum of row ends, spread on the host
    std::vector<int32_t> limit(keys.size(), 0x7FFFFFFF);
    if (!na_key.empty()) {
      std::vector<int32_t> na_limit(na_key.size());
      TRY(dmalloc(&d_na_key, na_key.size()));
      TRY(dmalloc(&d_na_limit, na_key.size()));
      TRYCU(cudaMemcpyAsync(d_na_key, na_key.data(), 8 * na_key.size(), cudaMemcpyHostToDevice, h->stream));
      k_vcf_na_limits<<<grid_for((int64_t)na_key.size(), 128), 128, 0, h->stream>>>(h->d_maxend, h->n, d_na_key, (int32_t)na_key.size(), d_na_limit);
      TRYCU(cudaMemcpyAsync(na_limit.data(), d_na_limit, 4 * na_key.size(), cudaMemcpyDeviceToHost, h->stream));
      TRYCU(cudaStreamSynchronize(h->stream));
      // entry v (0-based among the valid ones) is governed by the first NA entry with na_after > v
      size_t j = 0;
      for (size_t v = 0; v < keys.size(); v++) {
        while (j < na_after.size() && (size_t)na_after[j] <= v) j++;
        if (j < na_after.size()) limit[v] = na_limit[j];
      }
    }
    TRYCU(cudaMemcpyAsync(d_keys, keys.data(), 8 * keys.size(), cudaMemcpyHostToDevice, h->stream));
    TRYCU(cudaMemcpyAsync(d_idx, idx.data(), 4 * idx.size(), cudaMemcpyHostToDevice, h->stream));
    TRYCU(cudaMemcpyAsync(d_limit, limit.data(), 4 * limit.size(), cudaMemcpyHostToDevice, h->stream));
    if (pass) {
      TRY(dmalloc(&d_p, (size_t)h->n));
      TRYCU(cudaMemcpyAsync(d_p, pass, (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    }
    k_base_freqs<<<grid_for(h->n, 256), 256, 0, h->stream>>>(csr_of(h), d_p, d_keys, d_idx, d_limit, (int32_t)keys.size(), n_vcf, d_counts);
    TRYCU(cudaStreamSynchronize(h->stream));  // the host vectors above must outlive their copies
  }
  k_counts_to_double<<<grid_for((int64_t)cells, 256), 256, 0, h->stream>>>(d_counts, (int64_t)cells, d_out);
  TRYCU(cudaGetLastError());
  TRYCU(d2h_result(freqs_out, d_out, 8 * cells, h->stream));
  TRYCU(cudaStreamSynchronize(h->stream));
#undef TRY
#undef TRYCU
  cleanup();
  return EA_OK;
}

// rcpp_fep (src/rcpp_fep.cpp:9-36): two-sided Fisher's exact test per row of a small table, HOST arithmetic like the
// reference's call into HTSlib's kt_fisher_exact, whose published algorithm is restated here (HTSlib is not a
// dependency): hypergeometric probabilities of the tables with the observed margins, summed from the smallest and from
// the largest n11 while they stay below the observed table's (relative guards 0.99999999 / 1.00000001), each step
// derived from the previous one and recomputed from log-gamma every 11th table.  INT32_MIN (NA) in any cell -> NaN.
namespace {
double fe_lbinom(int n, int k) { return (n == 0 || k == 0 || n == k) ? 0.0 : std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma(n - k + 1.0); }
struct FeState {
  int n11, n1_, n_1, n; double p;
  double at(int k11) {  // probability of the table with n11 = k11, stepping from the current one when adjacent
    if (k11 % 11 && k11 + n - n1_ - n_1) {
      if (k11 == n11 + 1) { p *= (double)(n1_ - n11) / k11 * (n_1 - n11) / (k11 + n - n1_ - n_1); n11 = k11; return p; }
      if (k11 == n11 - 1) { p *= (double)n11 / (n1_ - k11) * (n11 + n - n1_ - n_1) / (n_1 - k11); n11 = k11; return p; }
    }
    n11 = k11;
    p = std::exp(fe_lbinom(n1_, n11) + fe_lbinom(n - n1_, n_1 - n11) - fe_lbinom(n, n_1));
    return p;
  }
};
double fisher_two_sided(int a, int b, int c, int d) {
  FeState s{a, a + b, a + c, a + b + c + d, 0.0};
  const int mx = s.n_1 < s.n1_ ? s.n_1 : s.n1_;
  const int mn = std::max(0, s.n1_ + s.n_1 - s.n);
  if (mn == mx) return 1.0;
  s.p = std::exp(fe_lbinom(s.n1_, a) + fe_lbinom(s.n - s.n1_, s.n_1 - a) - fe_lbinom(s.n, s.n_1));
  const double q = s.p;
  if (q == 0.0) return 0.0;
  double p = s.at(mn), left = 0.0, right = 0.0;
  int i, j;
  for (i = mn + 1; p < 0.99999999 * q && i <= mx; ++i) { left += p; p = s.at(i); }
  if (p < 1.00000001 * q) left += p;
  p = s.at(mx);
  for (j = mx - 1; p < 0.99999999 * q && j >= 0; --j) { right += p; p = s.at(j); }
  if (p < 1.00000001 * q) right += p;
  return std::min(1.0, left + right);
}
}  // namespace

int ea_fisher_exact(const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d, int64_t n, double* p_out) {
  if (n < 0 || (n > 0 && (!a || !b || !c || !d || !p_out))) return fail(EA_ERR_ARG, "bad table");
  for (int64_t i = 0; i < n; i++) {
    if (a[i] == INT32_MIN || b[i] == INT32_MIN || c[i] == INT32_MIN || d[i] == INT32_MIN) { p_out[i] = std::nan(""); continue; }
    if (a[i] < 0 || b[i] < 0 || c[i] < 0 || d[i] < 0) return fail(EA_ERR_ARG, "negative count in row %lld", (long long)i);
    p_out[i] = fisher_two_sided(a[i], b[i], c[i], d[i]);
  }
  return EA_OK;
}

int ea_bed_report(ea_handle* h, int mode, const int32_t* bed_chr, const int32_t* bed_start, const int32_t* bed_end, int64_t nbed,
                  int32_t param, const ea_report_params* thr, const uint8_t* pass, int32_t* match_out, uint8_t* pass_out,
                  int64_t* counts_out) {
  if (!h || !counts_out) return fail(EA_ERR_ARG, "null argument");
  int rc;
  CU(cudaSetDevice(h->device));
  if ((rc = match_on_device(h, mode, bed_chr, bed_start, bed_end, nbed, param))) return rc;
  // pass flags: thresholds evaluated on the device, or the caller's flags, or all TRUE
  const uint8_t* d_pass = nullptr;
  if (thr && thr->threshold_reads) {
    if (!thr->ctx_meth || !thr->ctx_unmeth) return fail(EA_ERR_ARG, "ctx_meth/ctx_unmeth must be non-null");
    if ((rc = ensure_pass(h))) return rc;
    ThresholdSpec spec;
    parse_threshold_letters(thr->ctx_meth, thr->ctx_unmeth, thr->ooctx_meth, thr->ooctx_unmeth, &spec);
    ThrParams tp{thr->min_n_ctx, thr->min_ctx_meth_frac, thr->max_ooctx_meth_frac};
    CU(cudaMemsetAsync(h->d_pass, 0, (size_t)std::max<int64_t>(h->n, 1), h->stream));
    if ((rc = launch_threshold(h, spec, tp, h->d_pass, nullptr, nullptr))) return rc;
    d_pass = h->d_pass;
  } else if (pass && h->n > 0) {
    if ((rc = ensure_pass(h))) return rc;
    CU(cudaMemcpyAsync(h->d_pass, pass, (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    d_pass = h->d_pass;
  }
  const size_t n_entries = (size_t)(nbed + 1) * 4;
  unsigned long long* d_counts;
  if ((rc = dmalloc(&d_counts, n_entries))) return rc;
  CU(cudaMemsetAsync(d_counts, 0, 8 * n_entries, h->stream));
  if (h->n > 0) {
    const int smem_entries = n_entries <= 8192 ? (int)n_entries : 0;
    const int grid = std::max(1, std::min(grid_for(h->n, 256), h->sm_count * 8));
    k_bed_counts<<<grid, 256, (size_t)smem_entries * 4, h->stream>>>(csr_of(h), h->d_match, d_pass, (int)nbed, d_counts, smem_entries);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(counts_out, d_counts, 8 * n_entries, cudaMemcpyDeviceToHost, h->stream));
  if (match_out && h->n > 0) CU(cudaMemcpyAsync(match_out, h->d_match, 4 * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
  if (pass_out && h->n > 0) {
    if (d_pass) CU(cudaMemcpyAsync(pass_out, d_pass, (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    else memset(pass_out, 1, (size_t)h->n);
  }
  CU(cudaStreamSynchronize(h->stream));
  pool_free(d_counts);
  return EA_OK;
}

// ---- lMHL report -----------------------------------------------------------
int ea_mhl_report(ea_handle* h, const char* ctx, int32_t hmax, int32_t hmin, double max_ooctx_meth_frac, int64_t* nrow_out) {
  if (!h || !ctx || !nrow_out) return fail(EA_ERR_ARG, "null argument");
  int rc;
  CU(cudaSetDevice(h->device));
  h->mhl_rows = -1;
  if (h->max_len > 65535) return fail(EA_ERR_ARG, "lMHL report: rows longer than 65535 bytes are not supported (the reference's lookup table ends there)");
  MhlParams prm;
  prm.ctx_mask = 0;
  for (const unsigned char* p = (const unsigned char*)ctx; *p; p++) prm.ctx_mask |= 1u << code_of(*p);
  prm.n_cls = 0;
  for (int cls = 0; cls < 3; cls++) {  // CHH (H h: 2 10), CHG (X x: 6 14), CG (Z z: 7 15): a counter plane if any of its letters is in ctx
    const int code = cls == 0 ? 2 : (cls == 1 ? 6 : 7);
    prm.slot[cls] = ((prm.ctx_mask >> code) & 1u) ? prm.n_cls++ : -1;   // rows need the METHYLATED letter in ctx (rcpp_mhl_report.cpp:72-101)
  }
  if (prm.n_cls == 0) { prm.n_cls = 1; }
  prm.cap = hmax > 0 ? (uint32_t)std::min(hmax, 65536) : 65536u;
  prm.hmin = hmin;
  prm.max_ooctx = max_ooctx_meth_frac;
  WinIndex* ix = nullptr;
  if ((rc = build_index(h, kMhlS, &ix))) return rc;
  if (h->n == 0 || ix->n == 0) { h->mhl_rows = 0; *nrow_out = 0; return EA_OK; }
  if (h->mhl_cap == 0) {
    const int64_t cap = h->nbytes / 8 + 4096;
    if ((rc = dmalloc(&h->d_mhl_scratch, (size_t)cap))) return rc;
    for (auto& c : h->d_mhl_i) if ((rc = dmalloc(&c, (size_t)cap))) return rc;
    for (auto& c : h->d_mhl_d) if ((rc = dmalloc(&c, (size_t)cap))) return rc;
    h->mhl_cap = cap;
  }
  for (int attempt = 0; attempt < 2; attempt++) {
    CU(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned long long), h->stream));
    MhlOut out{h->d_mhl_scratch, h->mhl_cap, h->d_counters, ix->w_base, ix->w_cnt};
    WinTable wt;
    wt.rname = ix->rname; wt.pos0 = ix->pos0; wt.lo = ix->lo; wt.hi = ix->hi; wt.n = ix->n; wt.S = ix->S;
    const size_t smem = (size_t)2 * 3 * kMhlS * 8 + (size_t)2 * kMhlS * 4 + (size_t)2 * prm.n_cls * kMhlS * 4;
    // `ctx` made of whole letter pairs Hh / Xx / Zz (and nothing else): the mask sweep; otherwise the byte-serial one
    const uint32_t pair_bits = (1u << 2) | (1u << 10) | (1u << 6) | (1u << 14) | (1u << 7) | (1u << 15);
    const bool pairs = (prm.ctx_mask & ~pair_bits) == 0 && ((prm.ctx_mask >> 2) & 1u) == ((prm.ctx_mask >> 10) & 1u) &&
                       ((prm.ctx_mask >> 6) & 1u) == ((prm.ctx_mask >> 14) & 1u) && ((prm.ctx_mask >> 7) & 1u) == ((prm.ctx_mask >> 15) & 1u);
    const int cls = pairs ? (int)(((prm.ctx_mask >> 2) & 1u) | (((prm.ctx_mask >> 6) & 1u) << 1) | (((prm.ctx_mask >> 7) & 1u) << 2)) : 0;
    const int grid = std::max(1, std::min<int>(ix->n, h->sm_count * 32));
    // letter pairs on rows of at most 1024 positions: the planar kernel (lane per row over the call planes; also serves
    // handles without byte rows); anything else - and kernel policy 5, which keeps the byte-row kernel tested - reads bytes
    const bool use_pl = pairs && h->pl_ok && h->pl_th == 32 && g_kernel_policy == 0 && !getenv("EA_MHL_BYTES");
    if (use_pl) {
      // every per-position sum of a window is at most (rows over the window) x S(min(longest row, hmax)): below 2^31 the
      // difference arrays are 32-bit (native shared atomics, 40 KB per CTA)
      const unsigned long long hbound = std::min<unsigned long long>((unsigned long long)h->max_len, prm.cap);
      const unsigned long long sbound = hbound < 2 ? hbound : hbound * (hbound + 1) * (hbound + 2) / 6;
      const bool narrow = (unsigned long long)ix->max_rows * sbound < (1ull << 31) && !getenv("EA_MHL_WIDE");
      const size_t smem_pl = narrow ? (size_t)2 * 3 * kMhlS * 4 + (size_t)2 * kMhlS * 4 + (size_t)2 * prm.n_cls * kMhlS * 4 : smem;
      void (*kern)(Planar, WinTable, MhlParams, MhlOut) = nullptr;
#define EA_MHL_PL(C_) (narrow ? (void (*)(Planar, WinTable, MhlParams, MhlOut))k_window_mhl_pl<C_, int> : (void (*)(Planar, WinTable, MhlParams, MhlOut))k_window_mhl_pl<C_, long long>)
      switch (cls) {
        case 1: kern = EA_MHL_PL(1); break;
        case 2: kern = EA_MHL_PL(2); break;
        case 3: kern = EA_MHL_PL(3); break;
        case 4: kern = EA_MHL_PL(4); break;
        case 5: kern = EA_MHL_PL(5); break;
        case 6: kern = EA_MHL_PL(6); break;
        default: kern = EA_MHL_PL(7); break;
      }
#undef EA_MHL_PL
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pl));
      kern<<<grid, kMhlThreads, smem_pl, h->stream>>>(planar_of(h), wt, prm, out);
      h->report_kernel = narrow ? "k_window_mhl_pl<int> (planar call planes, row per lane, 32-bit difference arrays)" : "k_window_mhl_pl<int64> (planar call planes, row per lane)";
    } else {
      NEED_BYTES(h);
      void (*kern)(Csr, WinTable, MhlParams, MhlOut) = k_window_mhl<0>;
      switch (cls) {
        case 1: kern = k_window_mhl<1>; break;
        case 2: kern = k_window_mhl<2>; break;
        case 3: kern = k_window_mhl<3>; break;
        case 4: kern = k_window_mhl<4>; break;
        case 5: kern = k_window_mhl<5>; break;
        case 6: kern = k_window_mhl<6>; break;
        case 7: kern = k_window_mhl<7>; break;
        default: break;
      }
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kMhlThreads, smem, h->stream>>>(csr_of(h), wt, prm, out);
      h->report_kernel = "k_window_mhl (byte rows, thread per read)";
    }
    if ((rc = scan_windows(h, *ix))) return rc;   // multi-block scan beyond a few thousand windows
    MhlCols cols{h->d_mhl_i[0], h->d_mhl_i[1], h->d_mhl_i[2], h->d_mhl_i[3], h->d_mhl_i[4], h->d_mhl_d[0], h->d_mhl_d[1]};
    k_gather_mhl<<<std::max(1, std::min(ix->n / 4 + 1, h->sm_count * 16)), 128, 0, h->stream>>>(
        h->d_mhl_scratch, ix->w_base, ix->w_dst, ix->w_cnt, ix->rname, ix->n, h->mhl_cap, cols);
    CU(cudaGetLastError());
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, h->d_counters, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if ((int64_t)total > h->mhl_cap) {  // scratch too small: grow to the exact size and redo
      if (attempt == 1) return fail(EA_ERR_STATE, "lMHL row capacity exceeded twice");
      pool_free(h->d_mhl_scratch); h->d_mhl_scratch = nullptr;
      for (auto& c : h->d_mhl_i) { pool_free(c); c = nullptr; }
      for (auto& c : h->d_mhl_d) { pool_free(c); c = nullptr; }
      const int64_t cap = (int64_t)total + 1024;
      if ((rc = dmalloc(&h->d_mhl_scratch, (size_t)cap))) return rc;
      for (auto& c : h->d_mhl_i) if ((rc = dmalloc(&c, (size_t)cap))) return rc;
      for (auto& c : h->d_mhl_d) if ((rc = dmalloc(&c, (size_t)cap))) return rc;
      h->mhl_cap = cap;
      continue;
    }
    h->mhl_rows = (int64_t)total;
    break;
  }
  *nrow_out = h->mhl_rows;
  return EA_OK;
}

int ea_mhl_report_fetch(ea_handle* h, int32_t* rname, int32_t* strand, int32_t* pos, int32_t* context, int32_t* coverage,
                        double* length, double* lmhl) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  if (h->mhl_rows < 0) return fail(EA_ERR_STATE, "no lMHL report has been computed on this handle");
  CU(cudaSetDevice(h->device));
  int32_t* di[5] = {rname, strand, pos, context, coverage};
  double* dd[2] = {length, lmhl};
  if (h->mhl_rows > 0) {
    for (int c = 0; c < 5; c++)
      if (di[c]) CU(d2h_result(di[c], h->d_mhl_i[c], 4 * (size_t)h->mhl_rows, h->stream));
    for (int c = 0; c < 2; c++)
      if (dd[c]) CU(d2h_result(dd[c], h->d_mhl_d[c], 8 * (size_t)h->mhl_rows, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return EA_OK;
}

int ea_cx_report(ea_handle* h, const uint8_t* pass, const char* ctx, int64_t* nrow_out) {
  if (!h || !ctx || !nrow_out) return fail(EA_ERR_ARG, "null argument");
  int rc;
  CU(cudaSetDevice(h->device));
  const uint8_t* d_pass_in = nullptr;
  if (pass && h->n > 0) {
    if ((rc = ensure_pass(h))) return rc;
    CU(cudaMemcpyAsync(h->d_pass, pass, (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    d_pass_in = h->d_pass;
  }
  ReportPlan rp;
  rp.want = want_of(ctx);
  rp.threshold = false;
  rp.mode = 1;
  if ((rc = run_report(h, rp, d_pass_in, nullptr))) return rc;
  *nrow_out = h->n_rows;
  return EA_OK;
}

int ea_cx_report_fetch(ea_handle* h, int32_t* rname, int32_t* strand, int32_t* pos, int32_t* context, int32_t* meth,
                       int32_t* unmeth) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  if (h->n_rows < 0) return fail(EA_ERR_STATE, "no report has been computed on this handle");
  CU(cudaSetDevice(h->device));
  Trace tr("ea_cx_report_fetch");
  int32_t* dst[6] = {rname, strand, pos, context, meth, unmeth};
  if (h->n_rows > 0)
    for (int c = 0; c < 6; c++)
      if (dst[c]) CU(d2h_result(dst[c], h->d_cols[c] + h->row_first, 4 * (size_t)h->n_rows, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  tr.mark("D2H rows");
  return EA_OK;
}

// The last report as text (writer.cuh).  rname_levels: n_levels NUL-terminated chromosome names (the factor levels of
// bam$rname); header != 0 prepends the column-name line.  Returns the byte count; ea_cx_report_text_fetch copies the text.
int ea_cx_report_text(ea_handle* h, const char* const* rname_levels, int32_t n_levels, int header, int64_t* nbytes_out) {
  if (!h || !nbytes_out) return fail(EA_ERR_ARG, "null argument");
  if (h->n_rows < 0) return fail(EA_ERR_STATE, "no report has been computed on this handle");
  if (n_levels < 0 || (n_levels > 0 && !rname_levels)) return fail(EA_ERR_ARG, "bad rname levels");
  int rc;
  CU(cudaSetDevice(h->device));
  static const char kHeader[] = "rname\tstrand\tpos\tcontext\tmeth\tunmeth\n";
  const int64_t hdr = header ? (int64_t)(sizeof(kHeader) - 1) : 0;
  std::string chars;
  std::vector<int32_t> off((size_t)n_levels + 1, 0);
  for (int32_t i = 0; i < n_levels; i++) {
    if (!rname_levels[i]) return fail(EA_ERR_ARG, "null rname level %d", i);
    chars += rname_levels[i];
    off[(size_t)i + 1] = (int32_t)chars.size();
  }
  pool_free(h->d_text); h->d_text = nullptr; h->text_bytes = -1;
  const int64_t n = h->n_rows;
  char* d_chars = nullptr; int32_t* d_off = nullptr; int64_t* d_len = nullptr; int64_t* d_pos = nullptr; void* d_tmp = nullptr;
  auto cleanup = [&]() { FreeScope fs; pool_free(d_chars); pool_free(d_off); pool_free(d_len); pool_free(d_pos); pool_free(d_tmp); };
#define TRY(x) do { rc = (x); if (rc != EA_OK) { cleanup(); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return fail(EA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
  TRY(dmalloc(&d_chars, chars.size()));
  TRY(dmalloc(&d_off, off.size()));
  TRY(dmalloc(&d_len, (size_t)n + 1));
  TRY(dmalloc(&d_pos, (size_t)n + 1));
  if (!chars.empty()) TRYCU(cudaMemcpyAsync(d_chars, chars.data(), chars.size(), cudaMemcpyHostToDevice, h->stream));
  TRYCU(cudaMemcpyAsync(d_off, off.data(), 4 * off.size(), cudaMemcpyHostToDevice, h->stream));
  TRYCU(cudaMemsetAsync(d_len + n, 0, 8, h->stream));
  const TextLevels lv{d_chars, d_off, n_levels};
  const int64_t f0 = h->row_first;
  const ReportCols cols{h->d_cols[0] + f0, h->d_cols[1] + f0, h->d_cols[2] + f0, h->d_cols[3] + f0, h->d_cols[4] + f0, h->d_cols[5] + f0};
  if (n > 0) k_text_len<<<grid_for(n, 256), 256, 0, h->stream>>>(cols, n, lv, d_len);
  size_t tb = 0;
  TRYCU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_len, d_pos, (int)(n + 1), h->stream));
  TRYCU(pool_malloc(&d_tmp, std::max<size_t>(tb, 1)));
  TRYCU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_len, d_pos, (int)(n + 1), h->stream));
  int64_t body = 0;
  TRYCU(cudaMemcpyAsync(&body, d_pos + n, 8, cudaMemcpyDeviceToHost, h->stream));
  TRYCU(cudaStreamSynchronize(h->stream));  // also: chars / off must outlive their copies
  TRY(dmalloc(&h->d_text, (size_t)(hdr + body)));
  if (hdr) TRYCU(cudaMemcpyAsync(h->d_text, kHeader, (size_t)hdr, cudaMemcpyHostToDevice, h->stream));
  if (n > 0) k_text_write<<<grid_for(n, 256), 256, 0, h->stream>>>(cols, n, lv, d_pos, hdr, h->d_text);
  TRYCU(cudaGetLastError());
  TRYCU(cudaStreamSynchronize(h->stream));
#undef TRY
#undef TRYCU
  cleanup();
  h->text_bytes = hdr + body;
  *nbytes_out = h->text_bytes;
  return EA_OK;
}

int ea_cx_report_text_fetch(ea_handle* h, char* out) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  if (h->text_bytes < 0) return fail(EA_ERR_STATE, "no report text has been formatted on this handle");
  CU(cudaSetDevice(h->device));
  if (h->text_bytes > 0) {
    if (!out) return fail(EA_ERR_ARG, "null output buffer");
    CU(d2h_result(out, h->d_text, (size_t)h->text_bytes, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return EA_OK;
}

int ea_cytosine_report(ea_handle* h, const ea_report_params* p, uint8_t* pass_out, int64_t* nrow_out) {
  if (!h || !nrow_out) return fail(EA_ERR_ARG, "null argument");
  int rc;
  CU(cudaSetDevice(h->device));
  ReportPlan rp;
  if ((rc = make_plan_from_params(p, 0, &rp))) return rc;
  if ((rc = run_report(h, rp, nullptr, nullptr))) return rc;
  const int64_t r0 = h->own_set ? std::min(h->own_row_lo, h->n) : 0;   // a range shard reports the flags of the rows it owns
  if (pass_out && h->n > r0) {
    if (rp.threshold) CU(d2h_result(pass_out, h->d_pass + r0, (size_t)(h->n - r0), h->stream));
    else memset(pass_out, 1, (size_t)(h->n - r0));
    CU(cudaStreamSynchronize(h->stream));
  }
  *nrow_out = h->n_rows;
  return EA_OK;
}

int ea_set_owned_range(ea_handle* h, int64_t key_lo, int64_t key_hi, int64_t own_row_lo) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  if (key_lo < 0) { h->own_set = false; h->own_row_lo = 0; return EA_OK; }
  if (key_hi < key_lo || own_row_lo < 0 || own_row_lo > h->n) return fail(EA_ERR_ARG, "bad owned range");
  h->own_set = true; h->own_key_lo = key_lo; h->own_key_hi = key_hi; h->own_row_lo = own_row_lo;
  h->n_rows = -1;
  return EA_OK;
}

int ea_cytosine_report_device(ea_handle* h, const ea_report_params* p, int mode, ea_timing* t) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  int rc;
  CU(cudaSetDevice(h->device));
  ReportPlan rp;
  if ((rc = make_plan_from_params(p, mode, &rp))) return rc;
  return run_report(h, rp, nullptr, t);
}

const char* ea_report_kernel(const ea_handle* h) { return h ? h->report_kernel : ""; }

const char* ea_stage_name(int i) {
  static const char* names[3] = {"threshold", "window", "scan+gather"};
  return (i >= 0 && i < 3) ? names[i] : "";
}

int ea_flush_l2(ea_handle* h, int64_t bytes) {
  if (!h || bytes <= 0) return fail(EA_ERR_ARG, "bad argument");
  CU(cudaSetDevice(h->device));
  if (h->flush_bytes < bytes) {
    pool_free(h->d_flush); h->d_flush = nullptr; h->flush_bytes = 0;
    CU(pool_malloc(&h->d_flush, (size_t)bytes));
    h->flush_bytes = bytes;
  }
  CU(cudaMemsetAsync(h->d_flush, 0x5A, (size_t)bytes, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return EA_OK;
}

int ea_synth_device(const ea_synth_cfg* cfg, int64_t row_begin, int64_t row_end, int device, ea_handle** out) {
  if (!out) return fail(EA_ERR_ARG, "null out");
  *out = nullptr;
  easynth::Plan plan;
  if (easynth::make_plan(cfg, &plan)) return fail(EA_ERR_ARG, "bad synthetic configuration");
  if (row_begin < 0 || row_end < row_begin || row_end > plan.n_reads || row_end - row_begin > 0x7FFFFFFEll) return fail(EA_ERR_ARG, "bad row range");
  int rc;
  if ((rc = pick_device(device))) return rc;
  ea_handle* h = new ea_handle;
  h->device = device; h->n = row_end - row_begin;
  const int64_t n = h->n;
  int64_t* d_padded = nullptr;
#define TRY(x) do { rc = (x); if (rc != EA_OK) { pool_free(d_padded); ea_free(h); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { pool_free(d_padded); ea_free(h); return fail(EA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
  TRY(dmalloc(&h->d_rname, (size_t)n)); TRY(dmalloc(&h->d_start, (size_t)n)); TRY(dmalloc(&h->d_strand, (size_t)n));
  TRY(dmalloc(&h->d_len, (size_t)n)); TRY(dmalloc(&h->d_off, (size_t)n + 1)); TRY(dmalloc(&d_padded, (size_t)n + 1));
  TRYCU(cudaMemset(d_padded, 0, 8 * (size_t)(n + 1)));
  if (n > 0) k_synth_rows<<<grid_for(n, 256), 256>>>(plan, row_begin, n, h->d_rname, h->d_strand, h->d_start, h->d_len, d_padded);
  size_t tb = 0;
  TRYCU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_padded, h->d_off, (int)(n + 1)));
  void* tmp; TRYCU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
  TRYCU(cub::DeviceScan::ExclusiveSum(tmp, tb, d_padded, h->d_off, (int)(n + 1)));
  TRYCU(cudaMemcpy(&h->npadded, h->d_off + n, 8, cudaMemcpyDeviceToHost));
  pool_free(tmp);
  {
    int64_t* d_l64; int64_t* d_sum;
    TRY(dmalloc(&d_l64, (size_t)n)); TRY(dmalloc(&d_sum, 1));
    if (n > 0) k_len_i64<<<grid_for(n, 256), 256>>>(h->d_len, n, d_l64);
    size_t tb2 = 0;
    TRYCU(cub::DeviceReduce::Sum(nullptr, tb2, d_l64, d_sum, (int)n));
    void* tmp2; TRYCU(pool_malloc(&tmp2, std::max<size_t>(tb2, 1)));
    TRYCU(cub::DeviceReduce::Sum(tmp2, tb2, d_l64, d_sum, (int)n));
    TRYCU(cudaMemcpy(&h->nbytes, d_sum, 8, cudaMemcpyDeviceToHost));
    pool_free(tmp2); pool_free(d_l64); pool_free(d_sum);
  }
  TRY(dmalloc(&h->d_data, (size_t)h->npadded + 16));
  const int64_t n_chunks = h->npadded / 16;
  const int uniform = plan.len_min == plan.len_max ? (plan.len_min + 15) / 16 : 0;
  if (n_chunks > 0) k_synth_bytes<<<grid_for(n_chunks, 256), 256>>>(plan, row_begin, n, h->d_off, h->d_data, uniform);
  TRYCU(cudaDeviceSynchronize());
  pool_free(d_padded); d_padded = nullptr;
#undef TRY
#undef TRYCU
  return adopt_and_finish(h, out);
}

int ea_synth_host(const ea_synth_cfg* cfg, int64_t row_begin, int64_t row_end, int32_t* rname, int32_t* strand,
                  int32_t* start, int64_t* off, uint8_t* blob) {
  easynth::Plan plan;
  if (easynth::make_plan(cfg, &plan)) return fail(EA_ERR_ARG, "bad synthetic configuration");
  if (row_begin < 0 || row_end < row_begin || row_end > plan.n_reads || !off) return fail(EA_ERR_ARG, "bad row range");
  const int64_t n = row_end - row_begin;
  off[0] = 0;
  for (int64_t i = 0; i < n; i++) {
    const easynth::RowInfo r = easynth::row_info(plan, row_begin + i);
    if (rname) rname[i] = r.rname;
    if (strand) strand[i] = r.strand;
    if (start) start[i] = r.start;
    off[i + 1] = off[i] + r.len;
  }
  if (!blob) return EA_OK;
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
  return EA_OK;
}

int ea_export_host(ea_handle* h, int32_t* rname, int32_t* strand, int32_t* start, int64_t* off, uint8_t* blob) {
  if (!h) return fail(EA_ERR_ARG, "null handle");
  CU(cudaSetDevice(h->device));
  const int64_t n = h->n;
  if (n == 0) { if (off) off[0] = 0; return EA_OK; }
  int rc;
  if (off || blob) NEED_BYTES(h);
  if (rname) CU(cudaMemcpy(rname, h->d_rname, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  if (start) CU(cudaMemcpy(start, h->d_start, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  if (strand) {
    std::vector<uint8_t> s8((size_t)n);
    CU(cudaMemcpy(s8.data(), h->d_strand, (size_t)n, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; i++) strand[i] = s8[(size_t)i];
  }
  if (off || blob) {
    int64_t *d_l64, *d_toff;
    StreamTmp tmps(nullptr);   // default-stream work below
    if ((rc = dmalloc(&d_l64, (size_t)n + 1))) return rc;
    tmps.keep(d_l64);
    if ((rc = dmalloc(&d_toff, (size_t)n + 1))) return rc;
    tmps.keep(d_toff);
    CU(cudaMemset(d_l64, 0, 8 * (size_t)(n + 1)));
    k_len_i64<<<grid_for(n, 256), 256>>>(h->d_len, n, d_l64);
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_l64, d_toff, (int)(n + 1)));
    void* tmp; CU(pool_malloc(&tmp, std::max<size_t>(tb, 1)));
    tmps.keep(tmp);
    CU(cub::DeviceScan::ExclusiveSum(tmp, tb, d_l64, d_toff, (int)(n + 1)));
    if (off) CU(cudaMemcpy(off, d_toff, 8 * (size_t)(n + 1), cudaMemcpyDeviceToHost));
    if (blob && h->nbytes > 0) {
      uint8_t* d_tight;
      if ((rc = dmalloc(&d_tight, (size_t)h->nbytes))) return rc;
      tmps.keep(d_tight);
      k_unpad<<<std::max(1, std::min(grid_for(n * 32, 256), h->sm_count * 32)), 256>>>(h->d_data, h->d_off, h->d_len, d_toff, n, d_tight);
      CU(cudaMemcpy(blob, d_tight, (size_t)h->nbytes, cudaMemcpyDeviceToHost));
    }
  }
  CU(cudaGetLastError());
  return EA_OK;
}

}  // extern "C"

#include "multi.inc"
#include "scattered.inc"
