// exchange.cu — the gradient exchange of view-sharded training as ONE kernel over NVSwitch multicast memory
// (include/hidegs_exchange.h).  Replaces the ring all-reduce a library collective would run after the backward.
//
// Every rank owns 1/world of the flat gradient arena.  For its slice it issues `multimem.ld_reduce.add.v4.f32` on the
// multicast address — the switch fetches the 16 bytes from every replica and returns their sum — and writes the result
// back through the same address with `multimem.st.v4.f32`, which the switch replicates into every GPU.  Per GPU that is
// ~1x the arena out (its replicas of all slices are read once) and ~1x in (the sums), with the additions done in the
// fabric.  Two cross-rank barriers per CTA (flags in symmetric memory, CAS put / CAS take, system scope) bracket the
// data phase; CTA b of one rank only ever pairs with CTA b of the others, so no co-residency is assumed.
//
// hg_nvls_exchange_f32 adds an all-GATHER range to the same launch: rank r owns a slice of that range in its replica and
// replicates it into every GPU (plain 128-bit loads + multimem.st) — the SH gradient factors of the factored exchange
// (DESIGN.md §6), which every rank then turns into the summed SH rows locally (sh_from_factors_kernel).
#include "common.cuh"

#include <cstdlib>
#include "../../include/hidegs_exchange.h"

namespace hg {
namespace {

constexpr int kExThreads = 512;
#ifndef HG_NVLS_UNROLL
#define HG_NVLS_UNROLL 4
#endif
constexpr int kExDefaultBlocks = 32;  // measured on 8 B200: 16..48 CTAs x 512 threads saturate the NVLS path, more only add barrier traffic

__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* p, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float mc_ld_reduce1(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st1(float* p, const float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Barrier between the CTAs with this block index on every rank.  Slot [block][src] of rank dst's flag words is written
// only by rank src (0 -> 1) and cleared only by rank dst (1 -> 0), so the same slots serve every barrier of every call.
__device__ __forceinline__ void rank_barrier(const uint64_t* __restrict__ flag_ptrs, int rank, int world) {
  __syncthreads();
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    const int peer = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(flag_ptrs[peer]) + (size_t)blockIdx.x * world + rank;
    uint32_t* mine = reinterpret_cast<uint32_t*>(flag_ptrs[rank]) + (size_t)blockIdx.x * world + peer;
    __threadfence_system();
    while (atomicCAS_system(remote, 0u, 1u) != 0u) {
    }
    while (atomicCAS_system(mine, 1u, 0u) != 1u) {
    }
    __threadfence_system();
  }
  __syncthreads();
}

constexpr int kExMaxRanges = 8;
struct ExRanges {  // in float4 units relative to the multicast base
  int64_t off4[kExMaxRanges];
  int64_t n4[kExMaxRanges];
  int n;
  // all-gather part: every rank owns `gather_n4` float4 at gather_off4 + rank * gather_n4 of its replica and
  // replicates them into every GPU (0 = none)
  int64_t gather_off4;
  int64_t gather_n4;
};

// This rank's 1/world share of one contiguous range.
template <int kExUnroll>
__device__ __forceinline__ void reduce_share(float4* __restrict__ mc, const int64_t n4, const int rank, const int world) {
  const int64_t per = (n4 + world - 1) / world;
  const int64_t begin = per * rank;
  const int64_t end = begin + per < n4 ? begin + per : n4;
  const int64_t stride = (int64_t)gridDim.x * kExThreads;
  int64_t i = begin + (int64_t)blockIdx.x * kExThreads + threadIdx.x;
  for (; i + (kExUnroll - 1) * stride < end; i += kExUnroll * stride) {
    float4 v[kExUnroll];
#pragma unroll
    for (int u = 0; u < kExUnroll; ++u) v[u] = mc_ld_reduce(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < kExUnroll; ++u) mc_st(mc + i + u * stride, v[u]);
  }
  for (; i < end; i += stride) mc_st(mc + i, mc_ld_reduce(mc + i));
}

// This rank's slice of the gather range: plain loads from the local replica, multimem.st into every replica.
template <int kExUnroll>
__device__ __forceinline__ void gather_share(float4* __restrict__ mc, const float4* __restrict__ local, const int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * kExThreads;
  int64_t i = (int64_t)blockIdx.x * kExThreads + threadIdx.x;
  for (; i + (kExUnroll - 1) * stride < n4; i += kExUnroll * stride) {
    float4 v[kExUnroll];
#pragma unroll
    for (int u = 0; u < kExUnroll; ++u) v[u] = local[i + u * stride];
#pragma unroll
    for (int u = 0; u < kExUnroll; ++u) mc_st(mc + i + u * stride, v[u]);
  }
  for (; i < n4; i += stride) mc_st(mc + i, local[i]);
}

template <int kExUnroll>
__global__ void __launch_bounds__(kExThreads)
nvls_allreduce_kernel(float4* __restrict__ mc, const float4* __restrict__ local, const uint64_t* __restrict__ flag_ptrs,
                      const int rank, const int world, const ExRanges ranges, const int tail) {
  rank_barrier(flag_ptrs, rank, world);  // every replica of the ranges is complete
  if (ranges.gather_n4) {
    const int64_t o = ranges.gather_off4 + ranges.gather_n4 * rank;
    gather_share<kExUnroll>(mc + o, local + o, ranges.gather_n4);
  }
  for (int r = 0; r < ranges.n; ++r) reduce_share<kExUnroll>(mc + ranges.off4[r], ranges.n4[r], rank, world);
  if (tail && rank == 0 && blockIdx.x == 0 && (int)threadIdx.x < tail) {  // (single-range call with n % 4 floats left over)
    float* p = reinterpret_cast<float*>(mc + ranges.off4[0] + ranges.n4[0]) + threadIdx.x;
    mc_st1(p, mc_ld_reduce1(p));
  }
  __threadfence_system();
  rank_barrier(flag_ptrs, rank, world);  // every replica holds the sums; nobody still reads this rank's memory
}

}  // namespace
}  // namespace hg

extern "C" {

size_t hg_nvls_flag_words(int32_t world, int32_t max_blocks) {
  if (world < 1 || max_blocks < 1) return 0;
  return (size_t)world * (size_t)max_blocks;
}

static int launch_exchange(void* mc_ptr, const void* local_ptr, const uint64_t* flag_ptrs, int32_t rank, int32_t world,
                           const hg::ExRanges& ranges, int tail, int32_t blocks, void* stream) {
  const int grid = blocks ? blocks : hg::kExDefaultBlocks;
  static const int unroll = [] {
    const char* e = getenv("HG_NVLS_UNROLL");
    return e ? atoi(e) : HG_NVLS_UNROLL;
  }();
#define HG_EX_LAUNCH(U_)                                                                  \
  hg::nvls_allreduce_kernel<U_><<<grid, hg::kExThreads, 0, (cudaStream_t)stream>>>(       \
      (float4*)mc_ptr, (const float4*)local_ptr, flag_ptrs, rank, world, ranges, tail)
  if (unroll >= 8) HG_EX_LAUNCH(8);
  else if (unroll >= 4) HG_EX_LAUNCH(4);
  else if (unroll >= 2) HG_EX_LAUNCH(2);
  else HG_EX_LAUNCH(1);
#undef HG_EX_LAUNCH
  HG_POST_LAUNCH(false, (cudaStream_t)stream, "nvls_allreduce");
  return HG_OK;
}

static int check_exchange_args(const char* who, void* mc_ptr, const uint64_t* flag_ptrs, int32_t rank, int32_t world,
                               int32_t blocks) {
  if (world < 1 || rank < 0 || rank >= world || blocks < 0 || world > hg::kExThreads) {
    hg::set_error("%s: bad argument (rank %d, world %d, blocks %d)", who, rank, world, blocks);
    return HG_ERR_INVALID_ARG;
  }
  if (world > 1 && (!mc_ptr || !flag_ptrs)) {
    hg::set_error("%s: multicast pointer and flag table are mandatory", who);
    return HG_ERR_INVALID_ARG;
  }
  if ((uintptr_t)mc_ptr & 15) {
    hg::set_error("%s: the arena must be 16-byte aligned", who);
    return HG_ERR_INVALID_ARG;
  }
  return HG_OK;
}

int hg_nvls_allreduce_f32(void* mc_ptr, float* local_ptr, const uint64_t* flag_ptrs, int32_t rank, int32_t world,
                          int64_t n_floats, int32_t blocks, void* stream) {
  if (n_floats < 0) {
    hg::set_error("hg_nvls_allreduce_f32: bad argument (n %lld)", (long long)n_floats);
    return HG_ERR_INVALID_ARG;
  }
  if (world == 1 && rank == 0) return HG_OK;
  if (world > 1 && n_floats > 0 && !local_ptr) {
    hg::set_error("hg_nvls_allreduce_f32: multicast pointer, local pointer and flag table are mandatory");
    return HG_ERR_INVALID_ARG;
  }
  if ((uintptr_t)local_ptr & 15) {
    hg::set_error("hg_nvls_allreduce_f32: the arena must be 16-byte aligned");
    return HG_ERR_INVALID_ARG;
  }
  const int rc = check_exchange_args("hg_nvls_allreduce_f32", mc_ptr, flag_ptrs, rank, world, blocks);
  if (rc) return rc;
  if (n_floats == 0) return HG_OK;
  hg::ExRanges r{};
  r.n = 1;
  r.off4[0] = 0;
  r.n4[0] = n_floats / 4;
  return launch_exchange(mc_ptr, local_ptr, flag_ptrs, rank, world, r, (int)(n_floats % 4), blocks, stream);
}

static int exchange_ranges(const char* who, void* mc_ptr, const float* local_ptr, const uint64_t* flag_ptrs, int32_t rank,
                           int32_t world, int32_t n_ranges, const int64_t* offsets, const int64_t* counts,
                           int64_t gather_offset, int64_t gather_count, int32_t blocks, void* stream) {
  if (n_ranges < 0 || n_ranges > hg::kExMaxRanges || (n_ranges && (!offsets || !counts))) {
    hg::set_error("%s: bad argument (%d ranges, at most %d)", who, n_ranges, hg::kExMaxRanges);
    return HG_ERR_INVALID_ARG;
  }
  if (gather_offset < 0 || gather_count < 0 || (gather_offset & 3) || (gather_count & 3) ||
      (gather_count && (!local_ptr || ((uintptr_t)local_ptr & 15)))) {
    hg::set_error("%s: gather range (offset %lld, %lld floats per rank) must be multiples of 4 floats of a 16-byte "
                  "aligned replica", who, (long long)gather_offset, (long long)gather_count);
    return HG_ERR_INVALID_ARG;
  }
  if (world == 1 && rank == 0) return HG_OK;
  const int rc = check_exchange_args(who, mc_ptr, flag_ptrs, rank, world, blocks);
  if (rc) return rc;
  hg::ExRanges r{};
  for (int i = 0; i < n_ranges; ++i) {
    if (offsets[i] < 0 || counts[i] < 0 || (offsets[i] & 3) || (counts[i] & 3)) {
      hg::set_error("%s: range %d (offset %lld, count %lld) must be multiples of 4 floats", who, i,
                    (long long)offsets[i], (long long)counts[i]);
      return HG_ERR_INVALID_ARG;
    }
    if (counts[i] == 0) continue;
    r.off4[r.n] = offsets[i] / 4;
    r.n4[r.n] = counts[i] / 4;
    ++r.n;
  }
  r.gather_off4 = gather_offset / 4;
  r.gather_n4 = gather_count / 4;
  if (r.n == 0 && r.gather_n4 == 0) return HG_OK;
  return launch_exchange(mc_ptr, local_ptr, flag_ptrs, rank, world, r, 0, blocks, stream);
}

int hg_nvls_allreduce_ranges_f32(void* mc_ptr, const uint64_t* flag_ptrs, int32_t rank, int32_t world, int32_t n_ranges,
                                 const int64_t* offsets, const int64_t* counts, int32_t blocks, void* stream) {
  return exchange_ranges("hg_nvls_allreduce_ranges_f32", mc_ptr, nullptr, flag_ptrs, rank, world, n_ranges, offsets,
                         counts, 0, 0, blocks, stream);
}

int hg_nvls_exchange_f32(void* mc_ptr, const float* local_ptr, const uint64_t* flag_ptrs, int32_t rank, int32_t world,
                         int32_t n_ranges, const int64_t* offsets, const int64_t* counts, int64_t gather_offset,
                         int64_t gather_count, int32_t blocks, void* stream) {
  return exchange_ranges("hg_nvls_exchange_f32", mc_ptr, local_ptr, flag_ptrs, rank, world, n_ranges, offsets, counts,
                         gather_offset, gather_count, blocks, stream);
}

}  // extern "C"
