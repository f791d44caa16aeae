"""`_C` operator module of the drop-in rasterizer package.

Same operator names, argument order and return tuples as the reference's
pybind module (submodules/hierarchy-rasterizer/ext.cpp:15-18,
rasterize_points.h:18-80), implemented as a thin marshalling layer over the
C-ABI (include/hidegs_raster.h): tensors are only used for device memory and
the current CUDA stream.
"""
import ctypes

import torch

from .. import _lib


_ROUND = 32 << 20  # scratch size granularity (bytes)


def _ptr(t):
    """data_ptr of a tensor or NULL for an empty one (reference: empty == absent)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32(t):
    if t is None:
        return None
    if t.numel() and (t.dtype != torch.float32 or not t.is_cuda):
        raise RuntimeError("expected a float32 CUDA tensor, got %s on %s" % (t.dtype, t.device))
    return t.contiguous()


def _i32(t):
    if t is None:
        return None
    if t.numel() and (t.dtype != torch.int32 or not t.is_cuda):
        raise RuntimeError("expected an int32 CUDA tensor, got %s on %s" % (t.dtype, t.device))
    return t.contiguous()


_R_HINT = {}  # (device index, P, W, H) -> largest num_rendered seen: sizes the binning part of the workspace


def _up(n, a=256):
    return (int(n) + a - 1) // a * a


class _Workspace:
    """ONE allocation per forward call that holds all of the library's scratch:

        [ geometry | image | backward accumulator | binning (capacity from the last views' R) ]

    The reference makes three growable byte tensors per call (resizeFunctional,
    rasterize_points.cu:27-33, 91-97).  A single block of (nearly) constant size per view lets torch's
    caching allocator reach a steady state after one step; separate, differently sized blocks kept
    fragmenting it (10-100 ms of cudaMalloc per training step).  If a view needs more binning space
    than the hint, that part alone falls back to its own allocation and the hint is raised."""

    def __init__(self, device, P, W, H):
        self.device, self.key = device, (device.index, P, W, H)
        L0 = _lib.layout(P, W, H, 0)
        self.geom_bytes, self.image_bytes = L0.geom_bytes, L0.image_bytes
        self.accum_bytes = _lib.lib().hg_raster_backward_accum_bytes(P)
        self.off_image = _up(self.geom_bytes)
        self.off_accum = self.off_image + _up(self.image_bytes)
        self.off_binning = self.off_accum + _up(self.accum_bytes)
        hint = _R_HINT.get(self.key, 0)
        self.binning_cap = 0
        if hint > 0:
            r_cap = _up(int(hint * 1.125) + 4096, 1 << 18)
            self.binning_cap = _up(_lib.layout(P, W, H, r_cap).binning_bytes, _ROUND)
        self.tensor = torch.empty(self.off_binning + self.binning_cap, dtype=torch.uint8, device=device)
        self.binning = self.tensor[:0]
        base = self.tensor.data_ptr()
        self.cb_geom = _lib.ALLOC_FN(lambda ctx, n: base if n <= self.geom_bytes else None)
        self.cb_image = _lib.ALLOC_FN(lambda ctx, n: base + self.off_image if n <= self.image_bytes else None)
        self.cb_binning = _lib.ALLOC_FN(self._alloc_binning)

    def _alloc_binning(self, _ctx, nbytes):
        try:
            if nbytes <= self.binning_cap:
                self.binning = self.tensor[self.off_binning:self.off_binning + int(nbytes)]
            else:
                self.binning = torch.empty(_up(nbytes, _ROUND), dtype=torch.uint8, device=self.device)
            return self.binning.data_ptr()
        except Exception:  # surfaces as HG_ERR_ALLOC
            return None

    def finish(self, R):
        if R > _R_HINT.get(self.key, 0):
            _R_HINT[self.key] = int(R)
        image = self.tensor[self.off_image:self.off_image + self.image_bytes]
        out = (self.tensor, self.binning, image)
        # The callbacks close over `self`: drop them (and the tensors) so that this object is not a reference cycle
        # that keeps a few hundred MB alive until Python's cyclic collector happens to run.
        self.cb_geom = self.cb_image = self.cb_binning = None
        self.tensor = self.binning = None
        return out


def _accum_ptr(geomBuffer, P, W, H):
    """Address of the backward accumulator inside the forward's workspace (see _Workspace)."""
    L0 = _lib.layout(P, W, H, 0)
    off = _up(L0.geom_bytes) + _up(L0.image_bytes)
    need = off + _lib.lib().hg_raster_backward_accum_bytes(P)
    if geomBuffer.numel() < need:
        return None
    return geomBuffer.data_ptr() + off


def _inputs(P, N, degree, M, W, H, tan_fovx, tan_fovy, scale_modifier, prefiltered, render_geo, debug,
            bg, viewmatrix, projmatrix, campos, indices, parent_indices, ts, kids, means3D, sh, colors,
            all_map, opacity, scales, rotations, cov3D_precomp):
    s = _lib.RasterInputs()
    s.P, s.N, s.D, s.M, s.W, s.H = P, N, degree, M, W, H
    s.tan_fovx, s.tan_fovy, s.scale_modifier = tan_fovx, tan_fovy, scale_modifier
    s.prefiltered, s.render_geo, s.debug = int(prefiltered), int(render_geo), int(debug)
    s.background, s.viewmatrix, s.projmatrix, s.campos = _ptr(bg), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
    s.indices, s.parent_indices, s.ts, s.kids = _ptr(indices), _ptr(parent_indices), _ptr(ts), _ptr(kids)
    s.means3D, s.shs, s.colors_precomp, s.all_map = _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(all_map)
    s.opacities, s.scales, s.rotations, s.cov3D_precomp = _ptr(opacity), _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp)
    return s


def rasterize_gaussians(background, indices, parent_indices, ts, kids, means3D, colors, all_map, opacity,
                        scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                        tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, render_geo, debug,
                        do_depth):
    """RasterizeGaussiansCUDA (rasterize_points.cu:35-147)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("hidegs_b200 rasterizer needs CUDA tensors (there is no CPU path)")
    dev = means3D.device
    background, viewmatrix, projmatrix, campos = _f32(background), _f32(viewmatrix), _f32(projmatrix), _f32(campos)
    means3D, colors, all_map, opacity = _f32(means3D), _f32(colors), _f32(all_map), _f32(opacity)
    scales, rotations, cov3D_precomp, sh, ts = _f32(scales), _f32(rotations), _f32(cov3D_precomp), _f32(sh), _f32(ts)
    indices, parent_indices, kids = _i32(indices), _i32(parent_indices), _i32(kids)

    N = means3D.size(0)
    P = N if indices.numel() == 0 else indices.size(0)
    H, W = int(image_height), int(image_width)
    M = sh.size(1) if sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    rendered = ctypes.c_int32(0)
    if P != 0 and all_map.numel() != 0 and all_map.size(0) < P:
        raise RuntimeError("all_map must have one row per rendered slot")

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        ws = _Workspace(dev, P, W, H)  # largest block first: keeps the allocator from splitting it
        # The library writes every element of these outputs, so no zero fill is needed.  All float
        # images come from one allocation and both int vectors from another (two blocks of constant
        # size per view instead of six).
        HW = H * W
        nd = 1 if do_depth else 0
        out_f = torch.empty(((9 + nd) * HW,), **f32)
        out_color = out_f[:3 * HW].view(3, H, W)
        out_all_map = out_f[3 * HW:8 * HW].view(5, H, W)
        out_plane_depth = out_f[8 * HW:9 * HW].view(1, H, W)
        out_invdepth = out_f[9 * HW:].view(nd, H, W)
        out_i = torch.empty((2 * P,), **i32)
        radii, out_observe = out_i[:P], out_i[P:]
        s = _inputs(P, N, degree, M, W, H, tan_fovx, tan_fovy, scale_modifier, prefiltered, render_geo, debug,
                    background, viewmatrix, projmatrix, campos, indices, parent_indices, ts, kids, means3D, sh,
                    colors, all_map, opacity, scales, rotations, cov3D_precomp)
        rc = _lib.lib().hg_raster_forward(
            ctypes.byref(s), ws.cb_geom, None, ws.cb_binning, None, ws.cb_image, None,
            _ptr(out_color), _ptr(out_invdepth), _ptr(out_observe), _ptr(out_all_map), _ptr(out_plane_depth),
            _ptr(radii), ctypes.byref(rendered), stream)
    _lib.check(rc, "rasterize_gaussians")
    geomBuffer, binningBuffer, imgBuffer = ws.finish(rendered.value)
    return (rendered.value, out_color, radii, out_observe, out_all_map, out_plane_depth, geomBuffer,
            binningBuffer, imgBuffer, out_invdepth)


_gradient_arena_provider = None


def set_gradient_arena_provider(fn):
    """Opt-in: `fn(numel, device) -> flat fp32 tensor or None` supplies the arena the next backward calls write their
    gradients into (instead of a fresh allocation).  The caller owns the buffer: gradients returned by a backward are
    views of it and are overwritten by the next backward that receives the same buffer.  Used by the view-sharded
    data-parallel path so that the gradients are born in multicast symmetric memory (hidegs_b200/parallel.py)."""
    global _gradient_arena_provider
    _gradient_arena_provider = fn


_backward_chunk_hook = None


def set_backward_chunk_hook(n_chunks=0, fn=None):
    """Opt-in: issue the per-Gaussian part of the next backward calls in `n_chunks` slot ranges and call
    `fn(chunk, slot_begin, slot_end)` right after each range's kernel has been queued on the current stream
    (hg_raster_backward_chunked).  Rows slot_begin..slot_end-1 of every gradient are final once that kernel completes:
    the view-sharded data-parallel path starts the gradient exchange of those rows from the hook, on a side stream.
    `set_backward_chunk_hook()` removes the hook."""
    global _backward_chunk_hook
    _backward_chunk_hook = (int(n_chunks), fn) if fn is not None and n_chunks > 0 else None


def rasterize_gaussians_backward(background, all_map_pixels, indices, parent_indices, ts, kids, means3D, radii,
                                 colors, all_maps, opacities, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_all_map,
                                 dL_dout_plane_depth, dL_dout_invdepth, sh, degree, campos, geomBuffer, R,
                                 binningBuffer, imageBuffer, render_geo, debug, sh_sink=None):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:149-279).

    `sh_sink` (extension, optional): `(tensor [N, M, 3], beta)` — the SH gradient is accumulated into that tensor
    (`sink = beta * sink + grad`, hg_raster_backward_chunked) and the returned dL_dsh is None.  Returns False through
    `sh_sink_supported()` for calls that cannot use it (index remap, no SH)."""
    dev = means3D.device
    background, viewmatrix, projmatrix, campos = _f32(background), _f32(viewmatrix), _f32(projmatrix), _f32(campos)
    means3D, colors, all_maps, opacities = _f32(means3D), _f32(colors), _f32(all_maps), _f32(opacities)
    scales, rotations, cov3D_precomp, sh, ts = _f32(scales), _f32(rotations), _f32(cov3D_precomp), _f32(sh), _f32(ts)
    indices, parent_indices, kids, radii = _i32(indices), _i32(parent_indices), _i32(kids), _i32(radii)
    all_map_pixels = _f32(all_map_pixels)
    dL_dout_color, dL_dout_all_map = _f32(dL_dout_color), _f32(dL_dout_all_map)
    dL_dout_plane_depth, dL_dout_invdepth = _f32(dL_dout_plane_depth), _f32(dL_dout_invdepth)

    fullP = means3D.size(0)
    P = fullP if indices.numel() == 0 else indices.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if sh.numel() != 0 else 0
    # With an index remap or parents the library accumulates into pre-zeroed rows.
    prezero = indices.numel() != 0 or parent_indices.numel() != 0 or P == 0
    has_depth_grad = dL_dout_invdepth is not None and dL_dout_invdepth.numel() != 0
    # One flat fp32 arena per backward.  The trainable parameters come first
    # (xyz 3 | sh 3M | opacity 1 | scale 3 | rotation 4 = 59 floats per Gaussian at M = 16), so the
    # view-sharded trainer can all-reduce `arena[:59 N]` in place without a pack kernel.
    widths = (("means3D", 3), ("sh", 3 * M), ("opacity", 1), ("scales", 3), ("rotations", 4), ("means2D", 3),
              ("colors", 3), ("cov3D", 6), ("all_map", 5), ("invdepths", 1 if has_depth_grad else 0))
    # (every block starts on a multiple of 4 floats, so that the float4 paths of the backward — SH staging, the SH sink —
    # see 16-byte aligned rows for any Gaussian count; with fullP % 4 == 0 the blocks are back to back)
    total = sum(_up(fullP * w, 4) for _, w in widths)
    arena = _gradient_arena_provider(total, dev) if _gradient_arena_provider is not None else None
    if arena is None:
        arena = (torch.zeros if prezero else torch.empty)((total,), dtype=torch.float32, device=dev)
    else:  # caller-owned arena (e.g. multicast symmetric memory for the in-fabric gradient exchange)
        if arena.numel() < total or arena.dtype != torch.float32 or arena.device != dev or not arena.is_contiguous():
            raise ValueError("gradient arena provider returned an unusable buffer")
        arena = arena[:total]
        if prezero:
            arena.zero_()
    g, off = {}, 0
    for name, w in widths:
        g[name] = arena[off:off + fullP * w].view(fullP, w)
        off += _up(fullP * w, 4)
    dL_dmeans3D, dL_dmeans2D, dL_dcolors, dL_dall_map = g["means3D"], g["means2D"], g["colors"], g["all_map"]
    dL_dopacity, dL_dcov3D, dL_dscales, dL_drotations = g["opacity"], g["cov3D"], g["scales"], g["rotations"]
    dL_dsh = g["sh"].view(fullP, M, 3)
    dL_dinvdepths = g["invdepths"] if has_depth_grad else torch.zeros((0, 1), dtype=torch.float32, device=dev)

    if P != 0:
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            accum_ptr = _accum_ptr(geomBuffer, P, W, H)
            if accum_ptr is None:  # foreign geometry buffer: fall back to a separate accumulator
                accum = torch.empty(_lib.lib().hg_raster_backward_accum_bytes(P), dtype=torch.uint8, device=dev)
                accum_ptr = accum.data_ptr()
            s = _inputs(P, fullP, degree, M, W, H, tan_fovx, tan_fovy, scale_modifier, False, render_geo, debug,
                        background, viewmatrix, projmatrix, campos, indices, parent_indices, ts, kids, means3D,
                        sh, colors, all_maps, opacities, scales, rotations, cov3D_precomp)
            args = (ctypes.byref(s), int(R), _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                    _ptr(all_map_pixels), _ptr(dL_dout_color), _ptr(dL_dout_all_map), _ptr(dL_dout_plane_depth),
                    _ptr(dL_dout_invdepth) if has_depth_grad else None, accum_ptr,
                    _ptr(dL_dmeans2D), None, _ptr(dL_dopacity), _ptr(dL_dcolors),
                    _ptr(dL_dinvdepths) if has_depth_grad else None, _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh),
                    _ptr(dL_dscales), _ptr(dL_drotations), _ptr(dL_dall_map))
            hook = _backward_chunk_hook
            if sh_sink is not None:
                sink_t, beta = sh_sink
                if (sink_t.dtype != torch.float32 or not sink_t.is_contiguous() or sink_t.numel() != fullP * 3 * M
                        or sink_t.device != dev):
                    raise ValueError("sh_sink must be a contiguous fp32 tensor of the SH gradient's shape")
                n_chunks, cb = (hook[0], None) if (hook is not None and not prezero) else (1, None)
                failure = []
                if hook is not None and not prezero:
                    def _on_chunk(_ctx, chunk, p0, p1, _stream, fn=hook[1]):
                        try:
                            fn(chunk, p0, p1)
                        except BaseException as e:  # noqa: BLE001
                            failure.append(e)
                    cb = _lib.CHUNK_FN(_on_chunk)
                rc = _lib.lib().hg_raster_backward_chunked(*args, n_chunks, cb if cb is not None else _lib.CHUNK_FN(),
                                                           None, sink_t.data_ptr(), float(beta), stream)
                if failure:
                    raise failure[0]
                dL_dsh = None
            elif hook is not None and not prezero:
                failure = []

                def _on_chunk(_ctx, chunk, p0, p1, _stream, fn=hook[1]):
                    try:
                        fn(chunk, p0, p1)
                    except BaseException as e:  # noqa: BLE001 — must not unwind through the C frame
                        failure.append(e)
                cb = _lib.CHUNK_FN(_on_chunk)
                rc = _lib.lib().hg_raster_backward_chunked(*args, hook[0], cb, None, None, 0.0, stream)
                if failure:
                    raise failure[0]
            else:
                rc = _lib.lib().hg_raster_backward(*args, stream)
        _lib.check(rc, "rasterize_gaussians_backward")
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations,
            dL_dall_map)


def sh_sink_supported(sh, indices, parent_indices):
    """Whether a backward over these inputs can accumulate its SH gradient into a sink (contiguous rows, 16-byte rows)."""
    return (sh is not None and sh.numel() != 0 and (indices is None or indices.numel() == 0)
            and (parent_indices is None or parent_indices.numel() == 0) and (3 * sh.size(1)) % 4 == 0
            and sh.data_ptr() % 16 == 0)


def mark_visible(positions, viewmatrix, projmatrix):
    """markVisible (rasterizer_impl.cu:145-157).  The reference's Python calls
    `_C.mark_visible` (diff_gaussian_rasterization/__init__.py:187) but never
    binds it (ext.cpp:15-18); it is bound here."""
    positions, viewmatrix, projmatrix = _f32(positions), _f32(viewmatrix), _f32(projmatrix)
    P = positions.size(0)
    present = torch.empty((P,), dtype=torch.bool, device=positions.device)
    if P:
        with torch.cuda.device(positions.device):
            rc = _lib.lib().hg_mark_visible(P, _ptr(positions), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "mark_visible")
    return present
