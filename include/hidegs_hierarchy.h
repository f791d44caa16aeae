/*
 * hidegs_hierarchy.h — C-ABI of the hierarchy file format and the static cut.
 *
 * Reference functions replaced (paths relative to submodules/gaussianhierarchy/ of the reference tree):
 *   hg_hier_probe / hg_hier_load     HierarchyLoader::load (hierarchy_loader.cpp:26-128), bound as
 *                                    gaussian_hierarchy._C.load_hierarchy (torch/torch_interface.cpp:18-50)
 *   hg_hier_write                    HierarchyWriter::write (hierarchy_writer.cpp:27-118), bound as
 *                                    gaussian_hierarchy._C.write_hierarchy (torch/torch_interface.cpp:52-75)
 *   hg_hier_decode_device            the element-by-element half -> float / HalfNode -> Node loops of
 *                                    HierarchyLoader::load (hierarchy_loader.cpp:87-126), run on the GPU over the raw
 *                                    payload so a compressed .hier goes disk -> pinned host -> HBM -> fp32 tensors
 *   hg_hier_encode_device            the float -> half / Node -> HalfNode loops of HierarchyWriter::write (:64-107)
 *   hg_expand_to_target              Traversal::expandToTarget (traversal.cpp:14-38), bound as
 *                                    gaussian_hierarchy._C.expand_to_target (torch/torch_interface.cpp:77-83)
 *
 * File layout (little endian), fp32 variant (first int P >= 0):
 *   int P | pos f32[P,3] | rot f32[P,4] | log-scale f32[P,3] | opacity f32[P] | sh f32[P,48]
 *   | int N | nodes i32[N,7] (depth, parent, start, count_leafs, count_merged, start_children, count_children;
 *   types.h:47-56) | boxes f32[N,2,4]
 * half variant (first int = -P):
 *   int -P | pos f32[P,3] | rot f16[P,4] | log-scale f16[P,3] | opacity f16[P] | sh f16[P,48]
 *   | int N | half nodes {i32 parent, start, start_children; i16 depth, count_children, count_leafs, count_merged}[N]
 *   (types.h:58-64) | boxes f16[N,2,4]
 * The writer refuses (as the reference does, hierarchy_writer.cpp:96-97) a node whose depth or counts exceed 32000.
 *
 * Host functions take HOST pointers; the *_device functions take DEVICE pointers and a stream.  int status returns
 * as in hidegs_raster.h (hg_last_error() holds the message, e.g. "File not found!").
 */
#ifndef HIDEGS_HIERARCHY_H
#define HIDEGS_HIERARCHY_H

#include <stddef.h>
#include <stdint.h>
#include "hidegs_raster.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Byte offsets of the sections of a .hier file, relative to the start of the file. */
typedef struct hg_hier_layout {
  int64_t P, N;
  int32_t compressed; /* 1 = half variant */
  int64_t pos, rot, scale, opacity, sh, nodes, boxes, file_bytes;
} hg_hier_layout;

/* Reads the two counts of `filename` and fills the section offsets. */
HG_API int hg_hier_probe(const char *filename, hg_hier_layout *out);

/* Loads either variant into caller-allocated HOST arrays sized from hg_hier_probe: pos [P,3], shs [P,48], alphas [P],
 * scales [P,3], rot [P,4], nodes [N,7] i32, boxes [N,2,4].  The half variant is widened to fp32 exactly. */
HG_API int hg_hier_load(const char *filename, float *pos, float *shs, float *alphas, float *scales, float *rot,
                        int32_t *nodes, float *boxes);

/* Reads the whole file into `raw` (HOST, at least layout.file_bytes) in one pass, for the device decode below. */
HG_API int hg_hier_read_raw(const char *filename, void *raw, int64_t capacity);

/* Writes either variant from HOST arrays (same shapes as hg_hier_load).  compressed != 0 rounds to half
 * (round-to-nearest-even, as half.hpp 2.2 does) and narrows the nodes. */
HG_API int hg_hier_write(const char *filename, int64_t P, int64_t N, const float *pos, const float *shs,
                         const float *opacities, const float *log_scales, const float *rotations,
                         const int32_t *nodes, const float *boxes, int32_t compressed);

/* Device decode of a raw file image resident in HBM (`raw_dev`, `layout` from hg_hier_probe of the same file) into
 * fp32 / int32 DEVICE arrays; works for both variants (the fp32 variant is a set of copies). */
HG_API int hg_hier_decode_device(const void *raw_dev, const hg_hier_layout *layout, float *pos, float *shs,
                                 float *alphas, float *scales, float *rot, int32_t *nodes, float *boxes, void *stream);

/* Device encode: builds the raw file image (either variant) in `raw_dev` from DEVICE arrays.  `overflow_flag`
 * (DEVICE int32, zeroed by the call) is set when a node would lose information in the half variant. */
HG_API int hg_hier_encode_device(void *raw_dev, const hg_hier_layout *layout, const float *pos, const float *shs,
                                 const float *opacities, const float *log_scales, const float *rotations,
                                 const int32_t *nodes, const float *boxes, int32_t *overflow_flag, void *stream);

/* Fills `layout` for P Gaussians / N nodes of the given variant (what hg_hier_probe would report for such a file). */
HG_API int hg_hier_layout_for(int64_t P, int64_t N, int32_t compressed, hg_hier_layout *out);

/* Static cut at depth `target` (HOST nodes [N,7]): depth-first from node 0, every node contributes its leaves; a node
 * with depth <= target contributes its merged Gaussians and is not expanded.  Returns the number of indices (also when
 * it exceeds `capacity`, in which case only `capacity` are written), or -1 on bad arguments. */
HG_API int64_t hg_expand_to_target(const int32_t *nodes, int64_t N, int32_t target, int32_t *out_indices,
                                   int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
