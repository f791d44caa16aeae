// binning.cu — tile-instance generation, per-tile depth sort and tile ranges.  All kernels here are hand-written for
// sm_100a; no library sort or scan runs on the forward path.
//
// Replaces, in order, cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:321), duplicateWithKeys (:70-115),
// cub::DeviceRadixSort::SortPairs on the 64-bit tile|depth key (:357-362, bits [0, 32+getHigherMsb(tiles))) and
// identifyTileRanges (:120-142) with its memset (:364).
//
// The reference sorts R tile instances (R ~ 5.4 M at config 2) GLOBALLY on 45..47-bit keys: six 8-bit radix passes
// over 12-byte pairs in HBM.  The sorted order it produces is "by tile, then by depth, ties by ascending slot" — and
// the tile part of that key is known before anything is sorted.  So the instances are BUCKETED by tile first and
// each tile's list is sorted on chip:
//   1. preprocess_fwd counts the instances of every tile while it computes the rectangles (one RED per instance);
//   2. tile_scan_kernel (ONE CTA) turns the counts into list offsets = the tile ranges, the total R, and four work
//      lists by list length; R and the list lengths travel to the host in one 32-byte copy (R is part of the API);
//   3. scatter_instances_kernel appends (depth bits, slot) to its tile's list through a per-tile cursor
//      (arrival order is arbitrary: 32 atomics in flight per warp, the instances of a warp's 32 splats are spread
//      evenly over its lanes whatever the splat sizes);
//   4. tile_sort kernels sort every list in SHARED MEMORY with a stable LSD radix sort (5-bit digits) over only the
//      depth bits in which the list's keys differ (min/max first: 24 bits = 5 passes at config 2).  Ranks come from
//      thread-private digit counters and one scan per pass — no atomics, no warp collectives in the element loops.  A
//      warp per short list (<= 1024, elements packed into one word after the first pass), a CTA per medium list
//      (<= 2048), a 512-thread CTA with 224 KB of shared memory per long list (<= 12288), the same routine over HBM
//      scratch for lists beyond that.  Ties in depth are put in ascending slot order afterwards (short runs in place;
//      a list with a long run is re-sorted slot-first, which the stable depth passes then preserve), so point_list is
//      bit-identical to the reference's, ties included.
// HBM traffic: 8 R written + 8 R read + 4 R written (the reference: 12 R written, then 6 x 24 R through the sort).
// The 64-bit keys themselves are only reconstructed on request (hg_raster_debug_keys) for the parity tests.
#include "common.cuh"
#include "tile_instances.cuh"

#include <mutex>

namespace hg {

namespace {

constexpr uint32_t kFull = 0xffffffffu;

// ---- list-length classes ------------------------------------------------------------------------------------------
constexpr int kSortWarpsA = 4;                      // tile_sort_kernel A: CTA of 4 warps
constexpr int kCapS = 1024;                         // short list: one warp
constexpr int kCapM = kCapS * kSortWarpsA / 2;      // medium list: the CTA (the short-list path holds 8 bytes per element)
constexpr int kSortWarpsB = 16;                     // kernel B: 512 threads, one CTA per SM
constexpr int kCapL = 12288;                        // 12288 * 16 B + 32 KB of counters = 224 KB
constexpr int kSortWarpsC = 32;                     // kernel C: lists beyond shared memory, sorted in HBM scratch

// keys + slots, twice (ping-pong), and 32 digits x threads 16-bit counters
__host__ __device__ constexpr size_t sort_smem_bytes(int warps, int cap) {
  return (size_t)cap * 16 + (size_t)warps * 32 * 32 * 2;
}

// ---- 2. counts -> offsets, ranges, work lists ------------------------------------------------------------------------
// header: [0] R  [1] #short  [2] #medium  [3] #long  [4] #beyond  [5] pairs in "beyond" lists  [6] longest list  [7] -
//         [8..11] work counters of the sort kernels (zeroed here)
// One CTA; rounds of 1024 consecutive tiles, one per thread, so that every global access of a warp is contiguous
// (a single SM issues all of this kernel's memory traffic: with four tiles per thread the 32-byte lane stride made it
// 40 k sector transactions and 18 us; now ~7 k).  `host_header` is mapped pinned host memory: the kernel stores the
// eight header words straight into it (no separate copy operation behind the kernel).
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int T, const int stride, const uint32_t cap_short, uint32_t* __restrict__ ctr,
                 uint2* __restrict__ ranges, uint32_t* __restrict__ list_a, uint32_t* __restrict__ list_b,
                 uint32_t* __restrict__ xl_off, uint32_t* __restrict__ header, volatile uint32_t* host_header) {
  __shared__ uint32_t s_warp[2][32];
  __shared__ uint32_t s_cnt[8];
  __shared__ uint32_t s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  if (tid < 8) s_cnt[tid] = 0;
  // total first (ranges of a view with a single instance stay (0, 0), see below)
  uint32_t part = 0;
  for (int t = tid; t < T; t += 1024) part += ctr[(size_t)t * stride];
  part = __reduce_add_sync(kFull, part);
  if (lane == 0) s_warp[1][warp] = part;
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = __reduce_add_sync(kFull, s_warp[1][lane]);
    if (lane == 0) s_total = v;
  }
  __syncthreads();
  const uint32_t total = s_total;
  uint32_t carry = 0, longest = 0;
  int round = 0;
  for (int base = 0; base < T; base += 1024, round ^= 1) {
    const int t = base + tid;
    const uint32_t n = t < T ? ctr[(size_t)t * stride] : 0u;
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[round][warp] = incl;  // double-buffered: one barrier per round
    __syncthreads();
    const uint32_t wsum = s_warp[round][lane];
    uint32_t wincl = wsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, wincl, o);
      if (lane >= o) wincl += v;
    }
    const uint32_t start = carry + __shfl_sync(kFull, wincl - wsum, warp) + incl - n;
    carry += __shfl_sync(kFull, wincl, 31);
    if (t < T) {
      ctr[(size_t)t * stride + 1] = start;
      // identifyTileRanges: empty tiles keep the memset's (0, 0); with a single instance in the whole view the end
      // marker is never written (rasterizer_impl.cu:133-141), so that tile reads (0, 0) as well.
      ranges[t] = (n && total != 1u) ? make_uint2(start, start + n) : make_uint2(0u, 0u);
    }
    longest = max(longest, n);
    const int cls = n == 0 ? 0 : n <= cap_short ? 1 : n <= (uint32_t)kCapM ? 2 : n <= (uint32_t)kCapL ? 3 : 4;
    uint32_t seen = __reduce_or_sync(kFull, 1u << cls) & ~1u;
    while (seen) {  // warp-aggregated appends, only for the classes this warp holds
      const int c = __ffs(seen) - 1;
      seen &= seen - 1;
      const uint32_t m = __ballot_sync(kFull, cls == c);
      uint32_t at = 0;
      if (lane == __ffs(m) - 1) at = atomicAdd(&s_cnt[c], (uint32_t)__popc(m));
      at = __shfl_sync(kFull, at, __ffs(m) - 1) + __popc(m & lt);
      if (cls == c) {
        if (c == 1) list_a[at] = (uint32_t)t;
        else if (c == 2) list_a[T - 1 - at] = (uint32_t)t;
        else if (c == 3) list_b[at] = (uint32_t)t;
        else {
          list_b[T - 1 - at] = (uint32_t)t;
          xl_off[at] = atomicAdd(&s_cnt[5], n);
        }
      }
    }
  }
  longest = __reduce_max_sync(kFull, longest);
  if (lane == 0) atomicMax(&s_cnt[6], longest);
  __syncthreads();
  if (tid < 8) {
    const uint32_t v = tid == 0 ? total : s_cnt[tid];
    header[tid] = v;
    if (host_header) host_header[tid] = v;
  } else if (tid < 12) {
    header[tid] = 0;
  }
}

// ---- 3. scatter ---------------------------------------------------------------------------------------------------
template <int PER>
__global__ void __launch_bounds__(256)
scatter_instances_kernel(const int P, const uint32_t* __restrict__ tiles_touched, const uint2* __restrict__ rects,
                         const float* __restrict__ depths, const uint32_t grid_x, const int stride,
                         uint32_t* __restrict__ ctr, uint2* __restrict__ pairs) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t cnt = idx < P ? tiles_touched[idx] : 0u;
  if (__ballot_sync(kFull, cnt != 0) == 0) return;
  uint2 r = make_uint2(0u, 0u);
  uint32_t dbits = 0;
  if (cnt) {
    r = rects[idx];
    dbits = __float_as_uint(depths[idx]);
  }
  const uint32_t first = (uint32_t)(idx - lane);
  WarpInstances wi(cnt, r.x, (r.y & 0xffffu) - (r.x & 0xffffu), lane);
  for (uint32_t j0 = 0; j0 < wi.total; j0 += 32 * PER) {  // PER instances per lane: PER atomics in flight
    int owner[PER];
    uint32_t tile[PER], d[PER], pos[PER];
    bool on[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      on[q] = wi.at(j0 + 32 * q, grid_x, owner[q], tile[q]);
      d[q] = __shfl_sync(kFull, dbits, owner[q]);
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) pos[q] = on[q] ? atomicAdd(ctr + (size_t)tile[q] * stride + 1, 1u) : 0u;
#pragma unroll
    for (int q = 0; q < PER; ++q)
      if (on[q]) pairs[pos[q]] = make_uint2(d[q], first + owner[q]);
  }
}

// ---- 4. per-list stable LSD radix sort ------------------------------------------------------------------------------
// NW warps cooperate on one list held in K0/V0 (keys / slots); K1/V1 is the other half of the ping-pong.  A pass
// ranks 5-bit digits WITHOUT warp collectives or atomics: thread t owns the m consecutive elements [t m, (t+1) m)
// (m odd: the strided shared-memory reads are conflict-free) and a private counter per digit, cnt[digit][t]; the
// flat array cnt[32][NT] read front to back IS the stable order (digit, owner thread, position), so one exclusive scan
// of it — 32 counters per thread, serial, then one warp scan of the per-thread totals — turns counts into
// destinations, which the owner hands out in a second walk over its elements.  (The warp-match ranking of 8-bit
// digits this replaced spent 3x the instructions: MATCH / eight ballots per 32 elements and a dependent
// shared-memory update chain.)
template <int NW>
__device__ __forceinline__ void group_sync() {
  if (NW == 1) __syncwarp();
  else __syncthreads();
}

constexpr int kDigitBits = 5;
constexpr uint32_t kDigits = 1u << kDigitBits;

template <typename CntT> struct CntPack;
template <> struct CntPack<uint16_t> { static constexpr int kWords = 16; };  // 32 counters in 16 words
template <> struct CntPack<uint32_t> { static constexpr int kWords = 32; };

// One counting pass over the list.  `ops` supplies the element type and three operations:
//   E load(int i)              the element at position i of the source buffer
//   uint32_t digit(const E&)   its 5-bit digit in this pass
//   void store(uint32_t pos, const E&)   put it at position pos of the destination buffer
template <int NW, typename CntT, typename Ops>
__device__ __forceinline__ void counting_pass(CntT* __restrict__ cnt, uint32_t* __restrict__ red, const int n,
                                              const int m, const int w, const int lane, const Ops ops) {
  constexpr int NT = NW * 32;
  constexpr int kWords = CntPack<CntT>::kWords;
  const int tid = w * 32 + lane;
  const int b = min(n, tid * m), e = min(n, b + m);
  {
    uint4* z = reinterpret_cast<uint4*>(cnt);
#pragma unroll
    for (int i = 0; i < kWords / 4; ++i) z[i * NT + tid] = make_uint4(0u, 0u, 0u, 0u);
  }
  group_sync<NW>();
  CntT* mine = cnt + tid;
  // Four elements per step: their counters are read together and written back in order, equal digits inside the
  // group folded into the increments — one shared-memory round trip per four elements instead of a read-modify-write
  // chain per element.
  int i = b;
  for (; i + 4 <= e; i += 4) {
    uint32_t d[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = ops.digit(ops.load(i + j));
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (uint32_t)mine[d[j] * NT];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t inc = 1;
#pragma unroll
      for (int q = 0; q < j; ++q) inc += d[q] == d[j];
      mine[d[j] * NT] = (CntT)(c[j] + inc);
    }
  }
  for (; i < e; ++i) mine[ops.digit(ops.load(i)) * NT] += 1;
  group_sync<NW>();
  // exclusive scan of the flat counter array: this thread's 32 consecutive counters, then across threads
  {
    uint4* row = reinterpret_cast<uint4*>(cnt) + tid * (kWords / 4);
    uint32_t c[kWords];
#pragma unroll
    for (int i = 0; i < kWords / 4; ++i) {
      const uint4 q = row[i];
      c[4 * i] = q.x; c[4 * i + 1] = q.y; c[4 * i + 2] = q.z; c[4 * i + 3] = q.w;
    }
    uint32_t total = 0;
    if (sizeof(CntT) == 2) {
      uint32_t acc = 0;  // two 16-bit lanes; neither overflows (n < 65536 in this variant)
#pragma unroll
      for (int i = 0; i < kWords; ++i) acc += c[i];
      total = (acc & 0xffffu) + (acc >> 16);
    } else {
#pragma unroll
      for (int i = 0; i < kWords; ++i) total += c[i];
    }
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    uint32_t run = incl - total;
    if (NW > 1) {
      if (lane == 31) red[w] = incl;
      __syncthreads();
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) run += ww < w ? red[ww] : 0u;
    }
    if (sizeof(CntT) == 2) {
#pragma unroll
      for (int i = 0; i < kWords; ++i) {
        const uint32_t lo = c[i] & 0xffffu, hi = c[i] >> 16;
        c[i] = run | ((run + lo) << 16);
        run += lo + hi;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kWords; ++i) {
        const uint32_t v = c[i];
        c[i] = run;
        run += v;
      }
    }
#pragma unroll
    for (int i = 0; i < kWords / 4; ++i) row[i] = make_uint4(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]);
  }
  group_sync<NW>();
  i = b;
  for (; i + 4 <= e; i += 4) {
    uint32_t d[4], c[4];
    typename Ops::E el[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      el[j] = ops.load(i + j);
      d[j] = ops.digit(el[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (uint32_t)mine[d[j] * NT];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t pos = c[j];
#pragma unroll
      for (int q = 0; q < j; ++q) pos += d[q] == d[j];
      mine[d[j] * NT] = (CntT)(pos + 1u);
      ops.store(pos, el[j]);
    }
  }
  for (; i < e; ++i) {
    const typename Ops::E el = ops.load(i);
    const uint32_t d = ops.digit(el);
    const uint32_t pos = mine[d * NT];
    mine[d * NT] = (CntT)(pos + 1u);
    ops.store(pos, el);
  }
  group_sync<NW>();
}

// (key, slot) pairs in two arrays each way; the digit comes from the key (depth bits above kmin) or from the slot
template <bool BY_SLOT>
struct PairOps {
  struct E { uint32_t k, v; };
  const uint32_t* __restrict__ K0;
  const uint32_t* __restrict__ V0;
  uint32_t* __restrict__ K1;
  uint32_t* __restrict__ V1;
  uint32_t kmin;
  int shift;
  __device__ __forceinline__ E load(int i) const { return E{K0[i], V0[i]}; }
  __device__ __forceinline__ uint32_t digit(const E& x) const {
    return ((BY_SLOT ? x.v : x.k - kmin) >> shift) & (kDigits - 1u);
  }
  __device__ __forceinline__ void store(uint32_t pos, const E& x) const {
    K1[pos] = x.k;
    V1[pos] = x.v;
  }
};

template <int NW, bool BY_SLOT, typename CntT>
__device__ __forceinline__ void radix_pass(const uint32_t* K0, const uint32_t* V0, uint32_t* K1, uint32_t* V1,
                                           CntT* cnt, uint32_t* red, const int n, const int m, const uint32_t kmin,
                                           const int shift, const int w, const int lane) {
  counting_pass<NW, CntT>(cnt, red, n, m, w, lane, PairOps<BY_SLOT>{K0, V0, K1, V1, kmin, shift});
}

// Equal depths: ascending slot (the reference's stable sort over its ascending-slot emission order).  Runs of up to 16
// equal keys are insertion-sorted by the thread that owns their head; returns whether a longer run exists.
template <int NT>
__device__ __forceinline__ uint32_t order_short_runs(const uint32_t* K, uint32_t* V, const int n, const int tid) {
  uint32_t long_run = 0;
  for (int i = tid; i < n - 1; i += NT) {
    const uint32_t k = K[i];
    if (K[i + 1] == k && (i == 0 || K[i - 1] != k)) {
      int e = i + 2;
      while (e < n && e < i + 17 && K[e] == k) ++e;
      if (e == i + 17) { long_run = 1; continue; }
      for (int a = i + 1; a < e; ++a) {
        const uint32_t v = V[a];
        int c = a - 1;
        while (c >= i && V[c] > v) { V[c + 1] = V[c]; --c; }
        V[c + 1] = v;
      }
    }
  }
  return long_run;
}

// Sorts list `src[0..n)` of (depth bits, slot) by (depth, slot) and writes the slots to dst[0..n).
template <int NW, typename CntT>
__device__ __forceinline__ void sort_list(uint32_t* K0, uint32_t* V0, uint32_t* K1, uint32_t* V1, CntT* cnt,
                                          uint32_t* red, const uint2* __restrict__ src,
                                          uint32_t* __restrict__ dst, const int n, const int slot_bits,
                                          const int w, const int lane) {
  constexpr int NT = NW * 32;
  const int tid = w * 32 + lane;
  const int m = ((n + NT - 1) / NT) | 1;
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll 4
  for (int i = tid; i < n; i += NT) {
    const uint2 p = src[i];
    K0[i] = p.x;
    V0[i] = p.y;
    kmin = min(kmin, p.x);
    kmax = max(kmax, p.x);
  }
  kmin = __reduce_min_sync(kFull, kmin);
  kmax = __reduce_max_sync(kFull, kmax);
  if (NW > 1) {
    if (lane == 0) {
      red[32 + w] = kmin;
      red[64 + w] = kmax;
    }
    __syncthreads();
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) {
      kmin = min(kmin, red[32 + ww]);
      kmax = max(kmax, red[64 + ww]);
    }
  } else {
    __syncwarp();
  }
  const int bits = 32 - __clz(kmax - kmin);  // __clz(0) == 32
  bool by_slot_done = false;
  for (;;) {
    for (int shift = 0; shift < bits; shift += kDigitBits) {
      radix_pass<NW, false, CntT>(K0, V0, K1, V1, cnt, red, n, m, kmin, shift, w, lane);
      uint32_t* t = K0; K0 = K1; K1 = t;
      t = V0; V0 = V1; V1 = t;
    }
    if (by_slot_done) break;
    uint32_t long_run = order_short_runs<NT>(K0, V0, n, tid);
    long_run = __any_sync(kFull, long_run);
    if (NW > 1) {
      if (tid == 0) red[96] = 0;
      __syncthreads();
      if (long_run && lane == 0) red[96] = 1;
      __syncthreads();
      long_run = red[96];
    } else {
      __syncwarp();
    }
    if (!long_run) break;
    // a long run of equal depths: order the whole list by slot first; the depth passes keep that order among equals
    for (int shift = 0; shift < slot_bits; shift += kDigitBits) {
      radix_pass<NW, true, CntT>(K0, V0, K1, V1, cnt, red, n, m, 0u, shift, w, lane);
      uint32_t* t = K0; K0 = K1; K1 = t;
      t = V0; V0 = V1; V1 = t;
    }
    by_slot_done = true;
  }
  group_sync<NW>();
  for (int i = tid; i < n; i += NT) dst[i] = V0[i];
  group_sync<NW>();
}

// Short lists, one warp, 8 bytes of shared memory per element instead of 16: once the first digit has been consumed
// an element is ONE word — its remaining key bits above its arrival position (key' >> 5) << idx_bits | position — so the
// later passes move 4 bytes per element, and twice as many warps fit an SM.  After the last pass the position part
// gathers (depth bits, slot) back from the list's own segment in HBM (L2 hits).  Needs bits - 5 + idx_bits <= 32, i.e.
// for a list of 1024 a depth spread below 2^27 ulps (a factor 65536 in depth); returns false, with nothing written, when
// the list does not qualify or holds a run of more than 16 equal depths — the caller then takes the general routine.
struct PackFirstOps {
  struct E { uint32_t a; int i; };
  const uint32_t* __restrict__ A;
  uint32_t* __restrict__ B;
  uint32_t kmin;
  int idx_bits;
  __device__ __forceinline__ E load(int i) const { return E{A[i] - kmin, i}; }
  __device__ __forceinline__ uint32_t digit(const E& x) const { return x.a & (kDigits - 1u); }
  __device__ __forceinline__ void store(uint32_t pos, const E& x) const {
    B[pos] = ((x.a >> kDigitBits) << idx_bits) | (uint32_t)x.i;
  }
};
struct PackNextOps {
  struct E { uint32_t a; };
  const uint32_t* __restrict__ X;
  uint32_t* __restrict__ Y;
  int shift;
  __device__ __forceinline__ E load(int i) const { return E{X[i]}; }
  __device__ __forceinline__ uint32_t digit(const E& x) const { return (x.a >> shift) & (kDigits - 1u); }
  __device__ __forceinline__ void store(uint32_t pos, const E& x) const { Y[pos] = x.a; }
};

__device__ __forceinline__ bool sort_short_packed(uint32_t* A, uint32_t* B, uint16_t* cnt, const uint2* __restrict__ src,
                                                  uint32_t* __restrict__ dst, const int n, const int lane) {
  const int m = ((n + 31) >> 5) | 1;
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll 4
  for (int i = lane; i < n; i += 32) {
    const uint32_t k = src[i].x;
    A[i] = k;
    kmin = min(kmin, k);
    kmax = max(kmax, k);
  }
  kmin = __reduce_min_sync(kFull, kmin);
  kmax = __reduce_max_sync(kFull, kmax);
  __syncwarp();
  const int bits = 32 - __clz(kmax - kmin);
  const int idx_bits = n > 1 ? 32 - __clz(n - 1) : 1;  // positions < n fit
  if (bits - kDigitBits + idx_bits > 32 || (bits == 0 && n > 16)) return false;
  uint32_t* X = A;  // buffer that holds the current order
  uint32_t* Y = B;
  if (bits > 0) {
    counting_pass<1, uint16_t>(cnt, nullptr, n, m, 0, lane, PackFirstOps{A, B, kmin, idx_bits});
    X = B;
    Y = A;
    for (int done = kDigitBits; done < bits; done += kDigitBits) {
      counting_pass<1, uint16_t>(cnt, nullptr, n, m, 0, lane, PackNextOps{X, Y, idx_bits + done - kDigitBits});
      uint32_t* t = X; X = Y; Y = t;
    }
  }
  // positions -> (depth bits, slot): keys into Y, slots over the packed words
  const uint32_t mask = (1u << idx_bits) - 1u;
#pragma unroll 4
  for (int i = lane; i < n; i += 32) {
    const uint32_t at = bits > 0 ? (X[i] & mask) : (uint32_t)i;
    const uint2 p = src[at];
    Y[i] = p.x;
    X[i] = p.y;
  }
  __syncwarp();
  if (__any_sync(kFull, order_short_runs<32>(Y, X, n, lane))) return false;
  __syncwarp();
  for (int i = lane; i < n; i += 32) dst[i] = X[i];
  __syncwarp();
  return true;
}

// Kernel A: CTA of kSortWarpsA warps; the CTA first takes medium lists (all warps on one list), then every warp takes
// short lists on its own.  Work is handed out through two device counters (header[8], header[9]).
__global__ void __launch_bounds__(kSortWarpsA * 32)
tile_sort_small_kernel(const int T, const int stride, const uint32_t* __restrict__ ctr,
                       const uint32_t* __restrict__ list_a, uint32_t* __restrict__ header, uint2* pairs,
                       uint32_t* __restrict__ vals, const int slot_bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_red[128];
  __shared__ uint32_t s_item;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n_short = header[1], n_medium = header[2];
  uint32_t* const words = reinterpret_cast<uint32_t*>(smem_raw);
  uint16_t* const counters = reinterpret_cast<uint16_t*>(words + 4 * kCapM);  // [32][128] for the CTA, [32][32] per warp
  {
    uint32_t* K0 = words;
    uint32_t* V0 = K0 + kCapM;
    uint32_t* K1 = V0 + kCapM;
    uint32_t* V1 = K1 + kCapM;
    for (;;) {
      if (threadIdx.x == 0) s_item = atomicAdd(header + 9, 1u);
      __syncthreads();
      const uint32_t item = s_item;
      __syncthreads();
      if (item >= n_medium) break;
      const uint32_t t = list_a[T - 1 - item];
      const uint2 ce = *reinterpret_cast<const uint2*>(ctr + (size_t)t * stride);  // (count, end of list)
      const uint32_t n = ce.x, start = ce.y - ce.x;
      sort_list<kSortWarpsA, uint16_t>(K0, V0, K1, V1, counters, s_red, pairs + start, vals + start, (int)n, slot_bits,
                                       w, lane);
    }
  }
  {
    uint32_t* A = words + w * (2 * kCapS);
    uint32_t* B = A + kCapS;
    uint16_t* cnt = counters + w * (kDigits * 32);
    // the next list is claimed, and its counters fetched, before the current one is sorted: the three dependent
    // global round trips of a hand-out (counter, list entry, tile counters) overlap the sort
    auto claim = [&](uint2& ce) -> bool {
      uint32_t item = 0;
      if (lane == 0) item = atomicAdd(header + 8, 1u);
      item = __shfl_sync(kFull, item, 0);
      if (item >= n_short) return false;
      ce = *reinterpret_cast<const uint2*>(ctr + (size_t)list_a[item] * stride);  // (count, end of list)
      return true;
    };
    uint2 ce = make_uint2(0u, 0u), ce_next = make_uint2(0u, 0u);
    bool have = claim(ce);
    while (have) {
      const bool have_next = claim(ce_next);
      const uint32_t n = ce.x, start = ce.y - ce.x;
      if (!sort_short_packed(A, B, cnt, pairs + start, vals + start, (int)n, lane)) {
        // general routine: keys / slots in the warp's two buffers, the other half of the ping-pong in the list's own
        // (already consumed) segment of `pairs`
        uint32_t* spill = reinterpret_cast<uint32_t*>(pairs + start);
        sort_list<1, uint16_t>(A, B, spill, spill + n, cnt, s_red, pairs + start, vals + start, (int)n, slot_bits, 0,
                               lane);
      }
      ce = ce_next;
      have = have_next;
    }
  }
}

// Kernel B: long lists, one 512-thread CTA per list, ~225 KB of shared memory.
__global__ void __launch_bounds__(kSortWarpsB * 32)
tile_sort_long_kernel(const int stride, const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ list_b,
                      const uint32_t* __restrict__ header, const uint2* __restrict__ pairs,
                      uint32_t* __restrict__ vals, const int slot_bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_red[128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t* K0 = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* V0 = K0 + kCapL;
  uint32_t* K1 = V0 + kCapL;
  uint32_t* V1 = K1 + kCapL;
  uint16_t* cnt = reinterpret_cast<uint16_t*>(V1 + kCapL);
  const uint32_t n_long = header[3];
  for (uint32_t item = blockIdx.x; item < n_long; item += gridDim.x) {
    const uint32_t t = list_b[item];
    const uint2 ce = *reinterpret_cast<const uint2*>(ctr + (size_t)t * stride);
    const uint32_t n = ce.x, start = ce.y - ce.x;
    sort_list<kSortWarpsB, uint16_t>(K0, V0, K1, V1, cnt, s_red, pairs + start, vals + start, (int)n, slot_bits, w,
                                     lane);
  }
}

// Kernel C: lists that do not fit shared memory; the same routine with keys / slots in 16 n bytes of HBM scratch per
// list and 32-bit counters in shared memory (128 KB).
__global__ void __launch_bounds__(kSortWarpsC * 32)
tile_sort_beyond_kernel(const int T, const int stride, const uint32_t* __restrict__ ctr,
                        const uint32_t* __restrict__ list_b, const uint32_t* __restrict__ xl_off,
                        const uint32_t* __restrict__ header, const uint2* __restrict__ pairs,
                        uint32_t* __restrict__ vals, uint32_t* __restrict__ scratch, const int slot_bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_red[128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n_beyond = header[4];
  for (uint32_t item = blockIdx.x; item < n_beyond; item += gridDim.x) {
    const uint32_t t = list_b[T - 1 - item];
    const uint2 ce = *reinterpret_cast<const uint2*>(ctr + (size_t)t * stride);
    const uint32_t n = ce.x, start = ce.y - ce.x;
    uint32_t* base = scratch + 4 * (size_t)xl_off[item];
    sort_list<kSortWarpsC, uint32_t>(base, base + n, base + 2 * (size_t)n, base + 3 * (size_t)n,
                                     reinterpret_cast<uint32_t*>(smem_raw), s_red, pairs + start, vals + start, (int)n,
                                     slot_bits, w, lane);
  }
}

// ---- diagnostics (hg_raster_debug_keys) ------------------------------------------------------------------------------
// Inclusive sum of tiles_touched in slot order (the reference's point_offsets), one CTA.
__global__ void __launch_bounds__(1024)
debug_slot_offsets_kernel(const int P, const uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ offsets) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < P; base += 1024) {
    const int i = base + tid;
    const uint32_t v = i < P ? tiles_touched[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t add = s_carry;
    for (int ww = 0; ww < warp; ++ww) add += s_warp[ww];
    if (i < P) offsets[i] = incl + add;
    __syncthreads();
    if (tid == 1023) s_carry = incl + add;
    __syncthreads();
  }
}

// Reference emission (duplicateWithKeys, rasterizer_impl.cu:70-115): 64-bit keys in ascending slot order.
__global__ void __launch_bounds__(256)
emit_keys_reference_kernel(const int P, const float* __restrict__ depths,
                           const uint32_t* __restrict__ offsets, const uint2* __restrict__ rects,
                           const int* __restrict__ radii, const uint32_t grid_x,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  if (radii[idx] <= 0) return;
  uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
  const uint2 r = rects[idx];
  const uint32_t minx = r.x & 0xffffu, miny = r.x >> 16;
  const uint32_t maxx = r.y & 0xffffu, maxy = r.y >> 16;
  const uint64_t dbits = (uint64_t)__float_as_uint(depths[idx]);
  for (uint32_t y = miny; y < maxy; ++y) {
    for (uint32_t x = minx; x < maxx; ++x) {
      if (keys) keys[off] = ((uint64_t)(y * grid_x + x) << 32) | dbits;
      if (vals) vals[off] = (uint32_t)idx;
      ++off;
    }
  }
}

__global__ void __launch_bounds__(128)
rebuild_sorted_keys_kernel(const int stride, const uint32_t* __restrict__ ctr,
                           const uint32_t* __restrict__ point_list, const float* __restrict__ depths,
                           uint64_t* __restrict__ keys) {
  const uint32_t t = blockIdx.x;
  const uint32_t n = ctr[(size_t)t * stride], start = ctr[(size_t)t * stride + 1] - n;
  for (uint32_t i = threadIdx.x; i < n; i += 128)
    keys[start + i] = ((uint64_t)t << 32) | (uint64_t)__float_as_uint(depths[point_list[start + i]]);
}

}  // namespace

int launch_tile_scan(const GeomState& g, const ImageState& img, int T, uint32_t* host_header, cudaStream_t stream,
                     bool debug) {
  tile_scan_kernel<<<1, 1024, 0, stream>>>(T, g.ctr_stride, (uint32_t)kCapS, g.tile_ctr, img.ranges, g.list_a, g.list_b,
                                           g.xl_off, g.bin_header, host_header);
  HG_POST_LAUNCH(debug, stream, "tile_scan");
  return HG_OK;
}

size_t binning_scratch_bytes(uint32_t beyond_pairs) { return beyond_pairs ? 16 * (size_t)beyond_pairs + 256 : 0; }

int launch_binning(const hg_raster_inputs& in, const GeomState& g, const BinState& b, int T, dim3 grid,
                   const uint32_t* header_host, cudaStream_t stream) {
  static std::mutex attr_mu;
  static bool attr_done[64] = {};
  int dev = 0;
  HG_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("binning: device index %d out of range", dev);
    return HG_ERR_CUDA;
  }
  {  // the opt-in for > 48 KB of dynamic shared memory is a per-device function attribute
    std::lock_guard<std::mutex> lk(attr_mu);
    if (!attr_done[dev]) {
      HG_CUDA_TRY(cudaFuncSetAttribute(tile_sort_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sort_smem_bytes(kSortWarpsA, kCapM)));
      HG_CUDA_TRY(cudaFuncSetAttribute(tile_sort_beyond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kSortWarpsC * 32 * 32 * 4)));
      HG_CUDA_TRY(cudaFuncSetAttribute(tile_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sort_smem_bytes(kSortWarpsB, kCapL)));
      attr_done[dev] = true;
    }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool debug = in.debug != 0;
  // (one, two or four instances per lane measure the same 63 us at config 2: the kernel runs at the rate the L2
  // serves 5.4 M atomics with return plus as many scattered 8-byte stores)
  scatter_instances_kernel<2><<<(in.P + 255) / 256, 256, 0, stream>>>(in.P, g.tiles_touched, g.rects, g.depths, grid.x,
                                                                      g.ctr_stride, g.tile_ctr, b.pairs);
  HG_POST_LAUNCH(debug, stream, "scatter_instances");
  uint32_t slot_bits = 1;
  while (slot_bits < 32 && ((uint32_t)(in.P - 1) >> slot_bits)) ++slot_bits;
  const uint32_t n_short = header_host[1], n_medium = header_host[2], n_long = header_host[3], n_beyond = header_host[4];
  if (n_long) {  // heaviest first
    const int blocks = (int)(n_long < (uint32_t)sms ? n_long : (uint32_t)sms);
    tile_sort_long_kernel<<<blocks, kSortWarpsB * 32, sort_smem_bytes(kSortWarpsB, kCapL), stream>>>(
        g.ctr_stride, g.tile_ctr, g.list_b, g.bin_header, b.pairs, b.vals, (int)slot_bits);
    HG_POST_LAUNCH(debug, stream, "tile_sort_long");
  }
  if (n_beyond) {
    const int blocks = (int)(n_beyond < (uint32_t)sms ? n_beyond : (uint32_t)sms);
    tile_sort_beyond_kernel<<<blocks, kSortWarpsC * 32, kSortWarpsC * 32 * 32 * 4, stream>>>(T, g.ctr_stride, g.tile_ctr, g.list_b, g.xl_off,
                                                                      g.bin_header, b.pairs, b.vals, b.scratch,
                                                                      (int)slot_bits);
    HG_POST_LAUNCH(debug, stream, "tile_sort_beyond");
  }
  if (n_short + n_medium) {
    const size_t smem = sort_smem_bytes(kSortWarpsA, kCapM);
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    const uint32_t want = n_medium + (n_short + kSortWarpsA - 1) / kSortWarpsA;
    const uint32_t cap = (uint32_t)(sms * (per_sm > 0 ? per_sm : 1));
    const int blocks = (int)(want < cap ? want : cap);
    tile_sort_small_kernel<<<blocks, kSortWarpsA * 32, smem, stream>>>(T, g.ctr_stride, g.tile_ctr, g.list_a,
                                                                        g.bin_header, b.pairs, b.vals, (int)slot_bits);
    HG_POST_LAUNCH(debug, stream, "tile_sort_small");
  }
  return HG_OK;
}

int launch_debug_keys(int P, int T, const GeomState& g, const BinState& b, const int* radii, int R, dim3 grid,
                      uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys_sorted, cudaStream_t stream) {
  (void)R;
  if (keys_unsorted || vals_unsorted) {
    debug_slot_offsets_kernel<<<1, 1024, 0, stream>>>(P, g.tiles_touched, g.point_offsets);
    HG_POST_LAUNCH(true, stream, "debug_slot_offsets");
    emit_keys_reference_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.depths, g.point_offsets, g.rects, radii,
                                                                       grid.x, keys_unsorted, vals_unsorted);
    HG_POST_LAUNCH(true, stream, "emit_keys_reference");
  }
  if (keys_sorted) {
    rebuild_sorted_keys_kernel<<<T, 128, 0, stream>>>(g.ctr_stride, g.tile_ctr, b.vals, g.depths, keys_sorted);
    HG_POST_LAUNCH(true, stream, "rebuild_sorted_keys");
  }
  return HG_OK;
}

}  // namespace hg
