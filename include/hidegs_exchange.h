/*
 * hidegs_exchange.h — C-ABI of the per-step gradient exchange of view-sharded training.
 *
 * The reference trains on one GPU: its step ends with `loss.backward()` filling the `.grad` of the five parameter
 * tensors (scene/gaussian_model.py:175-233 registers them with the optimiser, scene/OurAdam.py:106 consumes them).
 * With camera views sharded over the GPUs of one NVSwitch box the only coupling between ranks is the SUM of those
 * gradients.  They live in ONE flat fp32 arena per rank (59 floats per Gaussian); when that arena is allocated in
 * symmetric memory with a multicast (NVLS) mapping, the sum is ONE kernel: every rank owns 1/world of the arena, pulls
 * the switch-reduced value of its slice (`multimem.ld_reduce.add.v4.f32`: the NVSwitch reads all replicas and adds
 * them) and broadcasts it back to every replica (`multimem.st.v4.f32`).  Each GPU moves ~1x the arena in and out
 * over NVLink instead of the 2(world-1)/world x of a ring, and no staging buffers are involved.
 *
 * Plumbing (allocation of the symmetric arena, exchange of the multicast / peer pointers) is done by the caller —
 * `torch.distributed._symmetric_memory` in hidegs_b200/parallel.py.  Conventions as in hidegs_raster.h.
 */
#ifndef HIDEGS_EXCHANGE_H
#define HIDEGS_EXCHANGE_H

#include <stddef.h>
#include <stdint.h>
#include "hidegs_raster.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of uint32 flag words every rank must provide (zero-initialised, symmetric) for `max_blocks` CTAs. */
HG_API size_t hg_nvls_flag_words(int32_t world, int32_t max_blocks);

/* In-place all-reduce (sum) of `n_floats` fp32 values that every rank holds at the same offset of a symmetric
 * allocation.
 *   mc_ptr      multicast (NVLS) address of the first value — stores/reductions through it reach every replica
 *   local_ptr   this rank's own replica of the same values (used only for the < 4-float tail and checks)
 *   flag_ptrs   DEVICE array of `world` pointers: flag_ptrs[r] = rank r's flag words as mapped in THIS process
 *               (peer-to-peer addresses; hg_nvls_flag_words words each, all zero before the first call)
 *   blocks      CTAs to launch (0 = default); must be identical on every rank and <= the max_blocks the flags were
 *               sized for
 * The kernel begins with a cross-rank barrier per CTA (every rank's arena is complete: the call is stream-ordered after
 * the local producer) and ends with one (every replica holds the sums and nobody still reads this rank's memory), so
 * the caller may overwrite the arena right after it on the same stream.  mc_ptr and local_ptr must be 16-byte aligned.
 * world == 1 is a no-op. */
HG_API int hg_nvls_allreduce_f32(void *mc_ptr, float *local_ptr, const uint64_t *flag_ptrs, int32_t rank,
                                 int32_t world, int64_t n_floats, int32_t blocks, void *stream);

/* The same exchange over up to 8 disjoint ranges of the arena in ONE kernel (offsets / counts in floats relative to
 * mc_ptr, multiples of 4; HOST arrays, read before the call returns).  This is the call a chunk hook of
 * hg_raster_backward_chunked issues on a side stream: a slot range of the per-Gaussian backward is final in the five
 * parameter blocks of the SoA arena (xyz | sh | opacity | scale | rotation), i.e. five ranges, and travels while the
 * next slot range is still being computed. */
HG_API int hg_nvls_allreduce_ranges_f32(void *mc_ptr, const uint64_t *flag_ptrs, int32_t rank, int32_t world,
                                        int32_t n_ranges, const int64_t *offsets, const int64_t *counts, int32_t blocks,
                                        void *stream);

/* Factored exchange: all-reduce ranges AND one all-gather range in the same kernel (same barriers).  Rank r owns
 * `gather_count` floats at gather_offset + r * gather_count of its replica (`local_ptr` = this rank's replica of
 * mc_ptr[0]) and replicates them into every GPU with multimem.st; the ranges are summed as above.  Used for the step
 * whose SH gradient travels as factors (hg_raster_backward_chunked's `sh_factor`): the 11 non-SH floats per Gaussian
 * are summed, the 3 N + 4 factor floats of every rank's view are gathered, and every rank rebuilds the summed SH rows
 * locally with hg_sh_gradient_from_factors — (44 + 12 world) bytes per Gaussian through the fabric instead of 236. */
HG_API int hg_nvls_exchange_f32(void *mc_ptr, const float *local_ptr, const uint64_t *flag_ptrs, int32_t rank,
                                int32_t world, int32_t n_ranges, const int64_t *offsets, const int64_t *counts,
                                int64_t gather_offset, int64_t gather_count, int32_t blocks, void *stream);

/* dL_dsh[g][k][c] = beta * dL_dsh[g][k][c] + sum_v basis_k(normalize(means3D[g] - campos_v)) * factor_v[g][c] over
 * `n_views` factor blocks laid out `view_stride` floats apart: block v = [N, 3] factors followed by campos_v (3 floats),
 * exactly what hg_raster_backward_chunked's `sh_factor` writes.  The basis is that of the reference's SH backward
 * (cuda_rasterizer/backward.cu:23-142 with auxiliary.h:34-51), views are added in index order, so every rank that
 * holds the same factor blocks forms the same bits.  beta: 0 (overwrite) or 1 (accumulate). */
HG_API int hg_sh_gradient_from_factors(int32_t N, int32_t D, int32_t M, int32_t n_views, const float *means3D,
                                       const float *factors, int64_t view_stride, float *dL_dsh, float beta,
                                       void *stream);

#ifdef __cplusplus
}
#endif
#endif
