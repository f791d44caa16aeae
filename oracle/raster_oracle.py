"""numpy/ctypes front-end of oracle/raster_oracle.c (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; nothing under hidegs_b200/ does.
It mirrors the reference's operator pair (rasterize_points.cu:35-279) on host
arrays: `forward(...)` returns the same outputs plus the intermediate state
(keys, sorted list, ranges, n_contrib, final_T), `backward(...)` the same nine
gradient arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_raster.so")
_lib = None

_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int32)
_up = ctypes.POINTER(ctypes.c_uint32)


class _Inputs(ctypes.Structure):
    _fields_ = [("P", ctypes.c_int32), ("N", ctypes.c_int32), ("D", ctypes.c_int32), ("M", ctypes.c_int32),
                ("W", ctypes.c_int32), ("H", ctypes.c_int32),
                ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float), ("scale_modifier", ctypes.c_float),
                ("render_geo", ctypes.c_int32), ("do_depth", ctypes.c_int32),
                ("bg", _fp), ("view", _fp), ("proj", _fp), ("campos", _fp),
                ("indices", _ip), ("parent_indices", _ip), ("ts", _fp), ("kids", _ip),
                ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("all_map", _fp), ("opacities", _fp),
                ("scales", _fp), ("rotations", _fp), ("cov3D_precomp", _fp)]


class _State(ctypes.Structure):
    _fields_ = [("depths", _fp), ("radii", _ip), ("rects", _ip), ("tiles_touched", _up), ("offsets", _up),
                ("means2D", _fp), ("cov3D", _fp), ("conic_opacity", _fp), ("rgb", _fp),
                ("clamped", ctypes.POINTER(ctypes.c_uint8))]


def build(force=False):
    """Compile the C restatement with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_raster.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.ora_preprocess.restype = ctypes.c_int64
    return _lib


def _arr(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    return a if a.size else None


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


class OracleRasterizer:
    """Holds one forward pass worth of state so that backward can follow."""

    def __init__(self, *, bg, viewmatrix, projmatrix, campos, means3D, opacities, image_height, image_width,
                 tanfovx, tanfovy, shs=None, colors_precomp=None, all_map=None, scales=None, rotations=None,
                 cov3D_precomp=None, scale_modifier=1.0, sh_degree=0, render_indices=None, parent_indices=None,
                 interpolation_weights=None, num_node_kids=None, render_geo=True, do_depth=True, nthreads=1):
        f32, i32 = np.float32, np.int32
        self.k = dict(bg=_arr(bg, f32), view=_arr(viewmatrix, f32), proj=_arr(projmatrix, f32),
                      campos=_arr(campos, f32), indices=_arr(render_indices, i32),
                      parent_indices=_arr(parent_indices, i32), ts=_arr(interpolation_weights, f32),
                      kids=_arr(num_node_kids, i32), means3D=_arr(means3D, f32), shs=_arr(shs, f32),
                      colors_precomp=_arr(colors_precomp, f32), all_map=_arr(all_map, f32),
                      opacities=_arr(opacities, f32), scales=_arr(scales, f32), rotations=_arr(rotations, f32),
                      cov3D_precomp=_arr(cov3D_precomp, f32))
        self.N = self.k["means3D"].shape[0]
        self.P = self.N if self.k["indices"] is None else self.k["indices"].shape[0]
        self.M = self.k["shs"].shape[1] if self.k["shs"] is not None else 0
        self.W, self.H = int(image_width), int(image_height)
        self.render_geo, self.do_depth = bool(render_geo), bool(do_depth)
        self.nthreads = int(nthreads)
        s = _Inputs()
        s.P, s.N, s.D, s.M, s.W, s.H = self.P, self.N, int(sh_degree), self.M, self.W, self.H
        s.tan_fovx, s.tan_fovy, s.scale_modifier = tanfovx, tanfovy, scale_modifier
        s.render_geo, s.do_depth = int(render_geo), int(do_depth)
        for name in ("bg", "view", "proj", "campos", "ts", "means3D", "shs", "colors_precomp", "all_map",
                     "opacities", "scales", "rotations", "cov3D_precomp"):
            setattr(s, name, _p(self.k[name], _fp))
        for name in ("indices", "parent_indices", "kids"):
            setattr(s, name, _p(self.k[name], _ip))
        self.inp = s
        self.gx, self.gy = (self.W + 15) // 16, (self.H + 15) // 16

    # ------------------------------------------------------------------ forward
    def forward(self, tile_step=1):
        P, W, H = self.P, self.W, self.H
        f32 = np.float32
        st = dict(depths=np.zeros(P, f32), radii=np.zeros(P, np.int32), rects=np.zeros((P, 2), np.int32),
                  tiles_touched=np.zeros(P, np.uint32), offsets=np.zeros(P, np.uint32),
                  means2D=np.zeros((P, 2), f32), cov3D=np.zeros((P, 6), f32), conic_opacity=np.zeros((P, 4), f32),
                  rgb=np.zeros((P, 3), f32), clamped=np.zeros((P, 3), np.uint8))
        cs = _State()
        cs.depths, cs.radii, cs.rects = _p(st["depths"], _fp), _p(st["radii"], _ip), _p(st["rects"], _ip)
        cs.tiles_touched, cs.offsets = _p(st["tiles_touched"], _up), _p(st["offsets"], _up)
        cs.means2D, cs.cov3D, cs.conic_opacity = _p(st["means2D"], _fp), _p(st["cov3D"], _fp), _p(st["conic_opacity"], _fp)
        cs.rgb, cs.clamped = _p(st["rgb"], _fp), st["clamped"].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        self.st, self.cs = st, cs
        L = lib()
        R = int(L.ora_preprocess(ctypes.byref(self.inp), ctypes.byref(cs), self.nthreads)) if P else 0
        self.R = R
        out = dict(num_rendered=R, radii=st["radii"],
                   color=np.zeros((3, H, W), f32), invdepth=np.zeros((1 if self.do_depth else 0, H, W), f32),
                   out_observe=np.zeros(P, np.int32), all_map=np.zeros((5, H, W), f32),
                   plane_depth=np.zeros((1, H, W), f32), final_T=np.zeros(H * W, f32),
                   n_contrib=np.zeros(H * W, np.uint32), ranges=np.zeros((self.gx * self.gy, 2), np.uint32),
                   keys_unsorted=np.zeros(R, np.uint64), vals_unsorted=np.zeros(R, np.uint32),
                   keys=np.zeros(R, np.uint64), point_list=np.zeros(R, np.uint32))
        if R > 0:  # rasterizer_impl.cu:332-333: nothing is rendered when R == 0
            u64p = ctypes.POINTER(ctypes.c_uint64)
            L.ora_binning(ctypes.byref(self.inp), ctypes.byref(cs), ctypes.c_int64(R),
                          out["keys_unsorted"].ctypes.data_as(u64p), _p(out["vals_unsorted"], _up),
                          out["keys"].ctypes.data_as(u64p), _p(out["point_list"], _up), _p(out["ranges"], _up))
            L.ora_render(ctypes.byref(self.inp), ctypes.byref(cs), _p(out["point_list"], _up), _p(out["ranges"], _up),
                         _p(out["color"], _fp), _p(out["invdepth"], _fp) if self.do_depth else _fp(),
                         _p(out["out_observe"], _ip), _p(out["all_map"], _fp), _p(out["plane_depth"], _fp),
                         _p(out["final_T"], _fp), _p(out["n_contrib"], _up), int(tile_step), self.nthreads)
        self.out = out
        return out

    # ----------------------------------------------------------------- backward
    def backward(self, dL_dcolor, dL_dall_map, dL_dplane_depth, dL_dinvdepth=None, tile_step=1):
        f32 = np.float32
        N, M = self.N, self.M
        g = dict(dL_dmeans2D=np.zeros((N, 3), f32), dL_dconic=np.zeros((N, 4), f32), dL_dopacity=np.zeros((N, 1), f32),
                 dL_dcolors=np.zeros((N, 3), f32), dL_dall_map=np.zeros((N, 5), f32), dL_dmeans3D=np.zeros((N, 3), f32),
                 dL_dcov3D=np.zeros((N, 6), f32), dL_dsh=np.zeros((N, M, 3), f32), dL_dscales=np.zeros((N, 3), f32),
                 dL_drotations=np.zeros((N, 4), f32))
        dcol = _arr(dL_dcolor, f32)
        dam = _arr(dL_dall_map, f32)
        dpd = _arr(dL_dplane_depth, f32)
        dinv = _arr(dL_dinvdepth, f32)
        g["dL_dinvdepths"] = np.zeros((N, 1), f32) if dinv is not None else None
        if self.P == 0 or self.R == 0:
            # R == 0 is undefined behaviour in the reference (ranges are never written); defined as zeros here.
            return g
        L = lib()
        o = self.out
        L.ora_render_backward(ctypes.byref(self.inp), ctypes.byref(self.cs), _p(o["point_list"], _up),
                              _p(o["ranges"], _up), _p(o["all_map"], _fp), _p(o["final_T"], _fp),
                              _p(o["n_contrib"], _up), _p(dcol, _fp), _p(dam, _fp), _p(dpd, _fp), _p(dinv, _fp),
                              _p(g["dL_dmeans2D"], _fp), _p(g["dL_dconic"], _fp), _p(g["dL_dopacity"], _fp),
                              _p(g["dL_dcolors"], _fp), _p(g["dL_dinvdepths"], _fp), _p(g["dL_dall_map"], _fp),
                              int(tile_step), self.nthreads)
        L.ora_preprocess_backward(ctypes.byref(self.inp), ctypes.byref(self.cs), _p(g["dL_dmeans2D"], _fp),
                                  _p(g["dL_dconic"], _fp), _p(g["dL_dopacity"], _fp), _p(g["dL_dcolors"], _fp),
                                  _p(g["dL_dinvdepths"], _fp), _p(g["dL_dmeans3D"], _fp), _p(g["dL_dcov3D"], _fp),
                                  _p(g["dL_dsh"], _fp), _p(g["dL_dscales"], _fp), _p(g["dL_drotations"], _fp))
        return g
