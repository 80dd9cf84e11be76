// Stable LSD radix sort of (u32 key, u32 value) pairs and an exclusive scan, written for the
// binning stage of the rasterizer.  They replace cub::DeviceRadixSort::SortPairs /
// cub::DeviceScan::InclusiveSum at rasterizer_impl.cu:419,447-452,467 and simple_knn.cu:210-213.
//
// Design for B200: the reference sorts R 64-bit (tile|depth) keys in 6 digit passes.  Here depth
// order is established once per Gaussian (P elements, 3-4 passes of 8 bits) and instances are
// emitted in that order, so only a stable sort by tile id (12-13 bits => 2 passes over 4-byte
// keys) touches the R-sized arrays.
//
// Small sorts (up to 2^20 keys; second version, round 2).  The first version ran three kernels per digit pass — per-CTA histogram,
// a scan of the 256 x nblk count matrix, ranked scatter — and read the keys twice per pass; the
// chain of ~15 short launches was most of the binning time of small scenes.  Now ONE kernel reads
// the keys once and produces the digit totals of EVERY pass, and each pass is a single kernel: a
// CTA ranks its chunk of keys (CH_CHUNK = 2048), publishes its 256 digit counts, and obtains the sum over the CTAs
// before it by looking back over their published counts (chained scan with decoupled look-back:
// status word = 2 flag bits + 30-bit count, "local count" first, "inclusive prefix" once known).
// CTAs take their chunk number from a ticket counter, so every chunk a CTA waits for is owned by
// a CTA that is already running: the wait cannot deadlock, whatever order the hardware schedules
// the grid in.  The output does not depend on that order either — a stable sort has exactly one
// result.  The chunk is staged in shared memory in sorted order BEFORE the look-back (the wait
// overlaps useful work) and written as contiguous runs per digit; the destination arrays
// (<= 68 MB each at R = 17 M) sit in the 126 MB L2, which merges the runs into full sectors.
#include "common.cuh"

namespace dogs {

constexpr int SORT_THREADS = 256;
#ifndef SORT_ITEMS_N
#define SORT_ITEMS_N 12    // keys per thread (a multiple of 4).  Measured at C3: 8 / 12 / 16 -> depth sort 0.101 / 0.098 / 0.100 ms,
                          // tile sort 0.183 / 0.173 / 0.185 ms (16: 89 registers, 2 CTAs/SM)
#endif
constexpr int SORT_ITEMS = SORT_ITEMS_N;
constexpr int SORT_CHUNK = SORT_THREADS * SORT_ITEMS;  // keys per CTA
constexpr int CH_ITEMS = 8;                            // chained single-kernel passes (small sorts): 8 keys per thread
constexpr int CH_CHUNK = SORT_THREADS * CH_ITEMS;      // (measured at C1: depth sort 0.036 ms with 8, 0.041 ms with 12)
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX = 256;
constexpr int SORT_MAX_PASSES = 4;
constexpr uint32_t ST_LOCAL = 1u << 30, ST_INCL = 2u << 30, ST_COUNT = (1u << 30) - 1u;

struct SortPasses { int n; int shift[SORT_MAX_PASSES]; uint32_t mask[SORT_MAX_PASSES]; };

// Device-side element count (SortCtl, common.cuh): the forward pass no longer waits for the host to read the
// instance count, so the tile sort is launched for the buffer capacity and takes the count from device memory.
__device__ __forceinline__ uint32_t sort_count(const SortCtl ctl, uint32_t n_cap) {
  return ctl.n_dev ? min(*ctl.n_dev, n_cap) : n_cap;
}
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Digit totals of every pass in one read of the keys (a fixed grid of CTAs striding over the chunks), and the
// look-back status words of the chunks set to "nothing published".  totals = [passes][256] (zeroed by the host
// before the launch), status = [passes][nblk_cap][256].
__global__ void __launch_bounds__(SORT_THREADS)
radix_totals_kernel(const uint32_t* __restrict__ keys, uint32_t n_cap, const SortPasses ps,
                    uint32_t* __restrict__ totals, uint32_t* __restrict__ status, uint32_t nblk_cap,
                    const SortCtl ctl) {
  pdl_wait();
  pdl_trigger();
  const uint32_t n = sort_count(ctl, n_cap);
  __shared__ uint32_t h[SORT_MAX_PASSES][RADIX];
#pragma unroll
  for (int p = 0; p < SORT_MAX_PASSES; p++) h[p][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t nblk = (n + CH_CHUNK - 1) / CH_CHUNK;
  for (int p = 0; p < ps.n; p++) {
    uint4* row = reinterpret_cast<uint4*>(status + (size_t)p * nblk_cap * RADIX);
    for (uint32_t i = blockIdx.x * SORT_THREADS + threadIdx.x; i < nblk * (RADIX / 4); i += gridDim.x * SORT_THREADS)
      row[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (uint32_t base = blockIdx.x * CH_CHUNK; base < n; base += gridDim.x * CH_CHUNK) {
    uint32_t k[CH_ITEMS];
#pragma unroll
    for (int i = 0; i < CH_ITEMS; i++) {
      const uint32_t idx = base + i * SORT_THREADS + threadIdx.x;
      k[i] = idx < n ? keys[idx] : 0u;
    }
#pragma unroll
    for (int i = 0; i < CH_ITEMS; i++) {
      const uint32_t idx = base + i * SORT_THREADS + threadIdx.x;
      if (idx < n)
        for (int p = 0; p < ps.n; p++) atomicAdd(&h[p][(k[i] >> ps.shift[p]) & ps.mask[p]], 1u);
    }
  }
  __syncthreads();
  for (int p = 0; p < ps.n; p++) {
    const uint32_t c = h[p][threadIdx.x];
    if (c) atomicAdd(&totals[p * RADIX + threadIdx.x], c);
  }
}

// One digit pass.  Keys are ranked stably inside the CTA (warp slices in order, ballot matching within a
// step); thread d owns digit d for the look-back.
__global__ void __launch_bounds__(SORT_THREADS)
radix_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n_cap,
                  int shift, uint32_t mask, const uint32_t* __restrict__ totals,
                  uint32_t* __restrict__ status, uint32_t* __restrict__ ticket, const SortCtl ctl) {
  pdl_wait();
  pdl_trigger();
  const uint32_t n = sort_count(ctl, n_cap);
  const int nbits = 32 - __clz(mask);
  __shared__ uint32_t cnt[SORT_WARPS][RADIX];
  __shared__ uint32_t lbase[RADIX];
  __shared__ uint32_t gbase[RADIX];
  __shared__ uint32_t skey[CH_CHUNK];
  __shared__ uint32_t sval[CH_CHUNK];
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t s_chunk;
  if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t chunk = s_chunk;
  if ((uint64_t)chunk * CH_CHUNK >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // issue every global load up front: digit totals, keys, values
  const uint32_t tot = totals[threadIdx.x];
  const uint32_t first = chunk * CH_CHUNK;
  const uint32_t wbase = first + warp * (32 * CH_ITEMS);
  uint32_t key[CH_ITEMS], val[CH_ITEMS], rank[CH_ITEMS];
#pragma unroll
  for (int s = 0; s < CH_ITEMS; s++) {
    uint32_t idx = wbase + s * 32 + lane;
    key[s] = idx < n ? keys_in[idx] : 0u;
    val[s] = idx < n ? (vals_in ? vals_in[idx] : idx) : 0u;   // null vals_in: the values are the indices
  }

  // global digit base = exclusive scan of the digit totals (thread d owns digit d)
  uint32_t x = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  uint32_t dbase = x - tot;
#pragma unroll
  for (int w = 0; w < 8; w++) if (w < warp) dbase += wsum[w];

  // Warp w ranks the contiguous slice [first + w*32*ITEMS, ...) in order, 32 keys per step.
#pragma unroll
  for (int s = 0; s < CH_ITEMS; s++) {
    uint32_t idx = wbase + s * 32 + lane;
    bool valid = idx < n;
    uint32_t d = valid ? ((key[s] >> shift) & mask) : 0xffffffffu;
    // Lanes holding the same digit, from one ballot per digit bit.  (__match_any_sync costs one round per
    // distinct value in the warp — up to 32 with 8-bit digits — and was 30 % of this kernel's time.)
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
    if (!valid) peers = ~peers;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      if (b < nbits) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
      }
    }
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[s] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // thread d: per-warp counts -> exclusive offsets over the warps; chunk total of digit d
  uint32_t run = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) {
    uint32_t t = cnt[w][threadIdx.x];
    cnt[w][threadIdx.x] = run;
    run += t;
  }
  // publish the chunk's count of digit d; chunk 0 has nothing before it
  uint32_t* my_status = status + (size_t)chunk * RADIX + threadIdx.x;
  st_status(my_status, (chunk == 0 ? ST_INCL : ST_LOCAL) | run);
  // local base = exclusive scan over the digits of the chunk totals
  uint32_t y2 = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, y2, o);
    if (lane >= o) y2 += y;
  }
  __syncthreads();            // wsum reuse
  if (lane == 31) wsum[warp] = y2;
  __syncthreads();
  uint32_t lb = y2 - run;
#pragma unroll
  for (int w = 0; w < 8; w++) if (w < warp) lb += wsum[w];
  lbase[threadIdx.x] = lb;
  __syncthreads();
  // stage in local sorted order
#pragma unroll
  for (int s = 0; s < CH_ITEMS; s++) {
    uint32_t idx = wbase + s * 32 + lane;
    if (idx < n) {
      uint32_t d = (key[s] >> shift) & mask;
      uint32_t lp = lbase[d] + cnt[warp][d] + rank[s];
      skey[lp] = key[s];
      sval[lp] = val[s];
    }
  }
  // look back: keys of digit d in the chunks before this one (only digits that exist in this pass).  Four
  // status words are requested at a time; a word that is not published yet is polled again.
  uint32_t before = 0;
  if (chunk > 0 && threadIdx.x <= mask) {
    int t = (int)chunk - 1;
    bool done = false;
    while (!done) {
      uint32_t sv[4];
#pragma unroll
      for (int q = 0; q < 4; q++) sv[q] = (t - q >= 0) ? ld_status(status + (size_t)(t - q) * RADIX + threadIdx.x) : ST_INCL;
      int used = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (done || used < q) continue;
        const uint32_t f = sv[q] >> 30;
        if (f == 0) continue;            // not published: stop here, poll from t - used again
        before += sv[q] & ST_COUNT;
        used = q + 1;
        if (f == 2) done = true;
      }
      t -= used;
    }
    st_status(my_status, ST_INCL | (before + run));
  }
  gbase[threadIdx.x] = dbase + before;
  __syncthreads();
  // contiguous runs per digit -> coalesced stores
  const uint32_t nvalid = min((uint32_t)CH_CHUNK, n - first);
#pragma unroll
  for (int s = 0; s < CH_ITEMS; s++) {
    uint32_t i = s * SORT_THREADS + threadIdx.x;
    if (i < nvalid) {
      uint32_t k = skey[i];
      uint32_t d = (k >> shift) & mask;
      uint32_t gp = gbase[d] + (i - lbase[d]);
      keys_out[gp] = k;
      vals_out[gp] = sval[i];
    }
  }
}

// ---- large sorts: three kernels per digit pass ---------------------------------------------------------------------
// Per-CTA digit histogram -> scan of the 256 x nblk count matrix (one CTA per digit) -> ranked scatter.  No CTA
// waits for another one.  Measured against the chained passes above at C3 (3 M depth keys, 10 M tile keys):
// depth sort 0.112 vs 0.119 ms, tile sort 0.199 vs 0.208 ms — with thousands of chunks the chain of inclusive
// prefixes, not the data movement, sets the pace of a chained pass; at C1 (100 k / 0.5 M keys) the chained passes
// win (0.036 vs 0.060 ms, 0.035 vs 0.042 ms) because the 15 short launches dominate there.
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n_cap, int shift, uint32_t mask,
                  uint32_t* __restrict__ hist, uint32_t nblk, const SortCtl ctl) {
  pdl_wait();
  pdl_trigger();
  const uint32_t n = sort_count(ctl, n_cap);
  __shared__ uint32_t h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * SORT_CHUNK;
  // every load of the chunk is in flight before the first count (ncu of the first version: 16 registers, one
  // load per thread at a time, 49 stall cycles per instruction on the scoreboard, 1.4 TB/s); a full chunk
  // comes as 16-byte loads, two per thread — the order of the keys does not matter to a histogram
  if (base + SORT_CHUNK <= n && (((uintptr_t)keys) & 15) == 0) {
    const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
    uint4 v[SORT_ITEMS / 4];
#pragma unroll
    for (int q = 0; q < SORT_ITEMS / 4; q++) v[q] = k4[q * SORT_THREADS + threadIdx.x];
#pragma unroll
    for (int q = 0; q < SORT_ITEMS / 4; q++) {
      atomicAdd(&h[(v[q].x >> shift) & mask], 1u); atomicAdd(&h[(v[q].y >> shift) & mask], 1u);
      atomicAdd(&h[(v[q].z >> shift) & mask], 1u); atomicAdd(&h[(v[q].w >> shift) & mask], 1u);
    }
  } else {
    uint32_t k[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
      const uint32_t idx = base + i * SORT_THREADS + threadIdx.x;
      k[i] = idx < n ? keys[idx] : 0u;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++)
      if (base + i * SORT_THREADS + threadIdx.x < n) atomicAdd(&h[(k[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// One CTA per digit: exclusive scan of that digit's per-block counts, digit total to totals[d].  A thread
// takes 8 consecutive counts, so one strip of the row is 8192 blocks (16 M keys) and the usual row needs a
// single round of barriers (the first version walked the row 256 counts at a time: 20 us at 6 k blocks).
constexpr int ROWSCAN_THREADS = 1024;
constexpr int ROWSCAN_ITEMS = 8;
__global__ void __launch_bounds__(ROWSCAN_THREADS)
radix_rowscan_kernel(uint32_t* __restrict__ hist, uint32_t nblk, uint32_t* __restrict__ totals) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t wsum[ROWSCAN_THREADS / 32];
  uint32_t* row = hist + (size_t)blockIdx.x * nblk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (uint32_t start = 0; start < nblk; start += ROWSCAN_THREADS * ROWSCAN_ITEMS) {
    const uint32_t i0 = start + threadIdx.x * ROWSCAN_ITEMS;
    uint32_t v[ROWSCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < ROWSCAN_ITEMS; k++) {
      v[k] = i0 + k < nblk ? row[i0 + k] : 0u;
      sum += v[k];
    }
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0, strip = 0;
#pragma unroll
    for (int w = 0; w < ROWSCAN_THREADS / 32; w++) {
      const uint32_t t = wsum[w];
      if (w < warp) woff += t;
      strip += t;
    }
    uint32_t ex = carry + woff + x - sum;
#pragma unroll
    for (int k = 0; k < ROWSCAN_ITEMS; k++) {
      if (i0 + k < nblk) row[i0 + k] = ex;
      ex += v[k];
    }
    carry += strip;
    __syncthreads();      // wsum is rewritten by the next strip
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// Ranked scatter.  Keys are ranked stably inside the CTA (warp slices in order, ballot matching
// within a step), staged in shared memory in local sorted order and written out as contiguous runs
// per digit: the L2 sees ~8x fewer (partial-sector) write transactions than with direct 4-byte
// scatter stores, which is what bounded the first version of this kernel (ncu: issue-active 9 %).
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n_cap,
                     int shift, uint32_t mask, const uint32_t* __restrict__ hist, uint32_t nblk,
                     const uint32_t* __restrict__ totals, const SortCtl ctl) {
  pdl_wait();
  pdl_trigger();
  const uint32_t n = sort_count(ctl, n_cap);
  if (blockIdx.x * SORT_CHUNK >= n) return;
  const int nbits = 32 - __clz(mask);
  __shared__ uint32_t cnt[SORT_WARPS][RADIX];
  __shared__ uint32_t lbase[RADIX];
  __shared__ uint32_t gbase[RADIX];
  __shared__ uint32_t skey[SORT_CHUNK];
  __shared__ uint32_t sval[SORT_CHUNK];
  __shared__ uint32_t wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) cnt[w][threadIdx.x] = 0;

  // issue every global load up front: digit totals, this block's histogram column, keys, values
  const uint32_t tot = totals[threadIdx.x];
  const uint32_t hcol = hist[threadIdx.x * nblk + blockIdx.x];
  const uint32_t wbase = blockIdx.x * SORT_CHUNK + warp * (32 * SORT_ITEMS);
  uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
  for (int s = 0; s < SORT_ITEMS; s++) {
    uint32_t idx = wbase + s * 32 + lane;
    key[s] = idx < n ? keys_in[idx] : 0u;
    val[s] = idx < n ? (vals_in ? vals_in[idx] : idx) : 0u;   // null vals_in: the values are the indices
  }

  // global digit base = exclusive scan of the digit totals (thread d owns digit d)
  uint32_t x = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  uint32_t dbase = x - tot;
#pragma unroll
  for (int w = 0; w < 8; w++) if (w < warp) dbase += wsum[w];

  // Warp w ranks the contiguous slice [base + w*32*ITEMS, ...) in order, 32 keys per step.
#pragma unroll
  for (int s = 0; s < SORT_ITEMS; s++) {
    uint32_t idx = wbase + s * 32 + lane;
    bool valid = idx < n;
    uint32_t d = valid ? ((key[s] >> shift) & mask) : 0xffffffffu;
    // Lanes holding the same digit, from one ballot per digit bit.  (__match_any_sync costs one round per
    // distinct value in the warp — up to 32 with 8-bit digits — and was 30 % of this kernel's time.)
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
    if (!valid) peers = ~peers;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      if (b < nbits) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
      }
    }
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[s] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // thread d: per-warp counts -> exclusive offsets over the warps; block total of digit d
  uint32_t run = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) {
    uint32_t t = cnt[w][threadIdx.x];
    cnt[w][threadIdx.x] = run;
    run += t;
  }
  // local base = exclusive scan over the digits of the block totals
  uint32_t y2 = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, y2, o);
    if (lane >= o) y2 += y;
  }
  __syncthreads();            // wsum reuse
  if (lane == 31) wsum[warp] = y2;
  __syncthreads();
  uint32_t lb = y2 - run;
#pragma unroll
  for (int w = 0; w < 8; w++) if (w < warp) lb += wsum[w];
  lbase[threadIdx.x] = lb;
  gbase[threadIdx.x] = hcol + dbase;
  __syncthreads();
  // stage in local sorted order
#pragma unroll
  for (int s = 0; s < SORT_ITEMS; s++) {
    uint32_t idx = wbase + s * 32 + lane;
    if (idx < n) {
      uint32_t d = (key[s] >> shift) & mask;
      uint32_t lp = lbase[d] + cnt[warp][d] + rank[s];
      skey[lp] = key[s];
      sval[lp] = val[s];
    }
  }
  __syncthreads();
  // contiguous runs per digit -> coalesced stores
  const uint32_t first = blockIdx.x * SORT_CHUNK;
  const uint32_t nvalid = min((uint32_t)SORT_CHUNK, n - first);
#pragma unroll
  for (int s = 0; s < SORT_ITEMS; s++) {
    uint32_t i = s * SORT_THREADS + threadIdx.x;
    if (i < nvalid) {
      uint32_t k = skey[i];
      uint32_t d = (k >> shift) & mask;
      uint32_t gp = gbase[d] + (i - lbase[d]);
      keys_out[gp] = k;
      vals_out[gp] = sval[i];
    }
  }
}

static inline size_t sort_head_bytes() { return align_up((SORT_MAX_PASSES * RADIX + SORT_MAX_PASSES) * sizeof(uint32_t)); }

// keys up to which a sort runs as chained single-kernel passes
constexpr size_t SORT_CHAINED_MAX = 1u << 20;

size_t sort_pairs_scratch_bytes(size_t n, int max_passes) {
  const size_t chunk = n <= SORT_CHAINED_MAX ? (size_t)CH_CHUNK : (size_t)SORT_CHUNK;
  size_t nblk = (n + chunk - 1) / chunk;
  if (nblk == 0) nblk = 1;
  if (max_passes < 1 || max_passes > SORT_MAX_PASSES) max_passes = SORT_MAX_PASSES;
  if (n <= SORT_CHAINED_MAX) return sort_head_bytes() + align_up((size_t)max_passes * nblk * RADIX * sizeof(uint32_t));
  return align_up(nblk * RADIX * sizeof(uint32_t)) + align_up(RADIX * sizeof(uint32_t));
}

int sort_pairs_u32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                   size_t n, int begin_bit, int end_bit, char* scratch, bool debug,
                   cudaStream_t s, int* result_in_b, bool vals_are_iota, const SortCtl* ctl_in) {
  const SortCtl ctl = ctl_in ? *ctl_in : SortCtl{nullptr};
  *result_in_b = 0;
  if (n == 0 || end_bit <= begin_bit) return 0;
  if (n > ST_COUNT) { set_error("sort_pairs: more than 2^30 - 1 elements"); return 1; }
  SortPasses ps;
  ps.n = 0;
  for (int p = 0; p < SORT_MAX_PASSES; p++) { ps.shift[p] = 0; ps.mask[p] = 0; }
  for (int b = begin_bit; b < end_bit; b += 8) {
    if (ps.n == SORT_MAX_PASSES) { set_error("sort_pairs: more than %d digit passes", SORT_MAX_PASSES); return 1; }
    const int nb = end_bit - b < 8 ? end_bit - b : 8;
    ps.shift[ps.n] = b;
    ps.mask[ps.n] = (1u << nb) - 1u;
    ps.n++;
  }
  const uint32_t chunk = n > SORT_CHAINED_MAX ? (uint32_t)SORT_CHUNK : (uint32_t)CH_CHUNK;
  const uint32_t nblk = (uint32_t)((n + chunk - 1) / chunk);
  if (n > SORT_CHAINED_MAX) {
    Carver cv(scratch);
    uint32_t* hist = cv.take<uint32_t>((size_t)nblk * RADIX);
    uint32_t* dtot = cv.take<uint32_t>(RADIX);
    uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
    for (int p = 0; p < ps.n; p++) {
      launch_pdl(radix_hist_kernel, dim3(nblk), dim3(SORT_THREADS), 0, s, ki, (uint32_t)n, ps.shift[p], ps.mask[p], hist, nblk, ctl);
      launch_pdl(radix_rowscan_kernel, dim3(RADIX), dim3(ROWSCAN_THREADS), 0, s, hist, nblk, dtot);
      launch_pdl(radix_scatter_kernel, dim3(nblk), dim3(SORT_THREADS), 0, s, ki,
                 (vals_are_iota && p == 0) ? (const uint32_t*)nullptr : (const uint32_t*)vi, ko, vo, (uint32_t)n,
                 ps.shift[p], ps.mask[p], hist, nblk, dtot, ctl);
      uint32_t* t;
      t = ki; ki = ko; ko = t;
      t = vi; vi = vo; vo = t;
      *result_in_b ^= 1;
    }
    return check_launch("radix sort", debug, s, 3 * ps.n);
  }
  uint32_t* totals = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* tickets = totals + SORT_MAX_PASSES * RADIX;
  uint32_t* status = reinterpret_cast<uint32_t*>(scratch + sort_head_bytes());
  cudaMemsetAsync(scratch, 0, sort_head_bytes(), s);
  const uint32_t tot_grid = nblk < 148u * 8u ? nblk : 148u * 8u;
  launch_pdl(radix_totals_kernel, dim3(tot_grid), dim3(SORT_THREADS), 0, s, keys_a, (uint32_t)n, ps, totals, status,
             nblk, ctl);
  uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  for (int p = 0; p < ps.n; p++) {
    launch_pdl(radix_pass_kernel, dim3(nblk), dim3(SORT_THREADS), 0, s, ki,
               (vals_are_iota && p == 0) ? (const uint32_t*)nullptr : (const uint32_t*)vi, ko, vo, (uint32_t)n,
               ps.shift[p], ps.mask[p], (const uint32_t*)(totals + p * RADIX), status + (size_t)p * nblk * RADIX,
               tickets + p, ctl);
    uint32_t* t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
    *result_in_b ^= 1;
  }
  return check_launch("radix sort", debug, s, ps.n + 1);
}

// ---- exclusive scan -----------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* wsum, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  uint32_t woff = 0, t = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++) {
    uint32_t ws = wsum[w];
    if (w < warp) woff += ws;
    t += ws;
  }
  *total = t;
  __syncthreads();
  return woff + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ idx, uint32_t n,
                   uint32_t* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t wsum[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS, sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t j = base + i;
    if (j < n) sum += in[idx ? idx[j] : j];
  }
  uint32_t total;
  block_exclusive_scan(sum, wsum, &total);
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_spine_kernel(uint32_t* __restrict__ partial, uint32_t nb, uint32_t* __restrict__ total_out) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t wsum[SCAN_THREADS / 32];
  uint32_t carry = 0;
  for (uint32_t start = 0; start < nb; start += SCAN_THREADS) {
    uint32_t i = start + threadIdx.x;
    uint32_t v = i < nb ? partial[i] : 0, tot;
    uint32_t ex = block_exclusive_scan(v, wsum, &tot);
    if (i < nb) partial[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ idx, uint32_t n,
                  const uint32_t* __restrict__ partial, uint32_t* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t wsum[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t j = base + i;
    v[i] = j < n ? in[idx ? idx[j] : j] : 0;
    sum += v[i];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(sum, wsum, &total) + partial[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t j = base + i;
    if (j < n) out[j] = ex;
    ex += v[i];
  }
}

// In-place exclusive scan of `nb` block sums by one CTA (the row scan above on a single row); total to *total_out.
int launch_block_base_scan(uint32_t* sums, uint32_t nb, uint32_t* total_out, bool debug, cudaStream_t s) {
  launch_pdl(radix_rowscan_kernel, dim3(1), dim3(ROWSCAN_THREADS), 0, s, sums, nb, total_out);
  return check_launch("block_base_scan", debug, s);
}

size_t scan_scratch_bytes(size_t n) {
  return align_up(((n + SCAN_CHUNK - 1) / SCAN_CHUNK + 1) * sizeof(uint32_t));
}

int exclusive_scan_u32(const uint32_t* in, const uint32_t* idx, uint32_t* out, uint32_t* total,
                       size_t n, char* scratch, bool debug, cudaStream_t s) {
  if (n == 0) {
    if (total) cudaMemsetAsync(total, 0, sizeof(uint32_t), s);
    return 0;
  }
  uint32_t nb = (uint32_t)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  uint32_t* partial = reinterpret_cast<uint32_t*>(scratch);
  launch_pdl(scan_reduce_kernel, dim3(nb), dim3(SCAN_THREADS), 0, s, in, idx, (uint32_t)n, partial);
  launch_pdl(scan_spine_kernel, dim3(1), dim3(SCAN_THREADS), 0, s, partial, nb, total);
  launch_pdl(scan_apply_kernel, dim3(nb), dim3(SCAN_THREADS), 0, s, in, idx, (uint32_t)n, (const uint32_t*)partial, out);
  return check_launch("exclusive scan", debug, s, 3);
}

}  // namespace dogs
