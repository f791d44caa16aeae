"""oracle/ply_oracle.py — restatement of how the reference writes / reads point_cloud.ply (TEST INFRASTRUCTURE; only
tests/ may import it).

Follows scene/gaussian_model.py:526-544 (save_ply: one structured row per Gaussian, `elements[:] = list(map(tuple,
attributes))`) and :322-355 (load_ply_file).  The container is written by the third-party `plyfile` package
(requirements.txt:158, plyfile==1.1), which is NOT installed in this image and cannot be fetched: its published output
for a single float32 `vertex` element on a little-endian host is restated here — header lines `ply`,
`format binary_little_endian 1.0`, `element vertex N`, one `property float <name>` per field, `end_header`, each
terminated by "\\n", followed by the packed structured array.

PARITY UNPINNED: no golden file produced by plyfile itself is available; parity of the .ply path rests on this
restatement, on the reference's own read code (load_ply_file, restated literally below) and on round trips.
"""
import numpy as np


def attribute_names(n_dc=3, n_rest=45, n_scale=3, n_rot=4):  # gaussian_model.py:472-484
    l = ["x", "y", "z", "nx", "ny", "nz"]
    for i in range(n_dc):
        l.append("f_dc_{}".format(i))
    for i in range(n_rest):
        l.append("f_rest_{}".format(i))
    l.append("opacity")
    for i in range(n_scale):
        l.append("scale_{}".format(i))
    for i in range(n_rot):
        l.append("rot_{}".format(i))
    return l


def file_bytes(xyz, features_dc, features_rest, opacity, scaling, rotation):
    """Arguments in the model's layout (numpy): features_dc [N,1,3], features_rest [N,15,3]."""
    normals = np.zeros_like(xyz)
    f_dc = np.ascontiguousarray(np.transpose(features_dc, (0, 2, 1))).reshape(len(xyz), -1)
    f_rest = np.ascontiguousarray(np.transpose(features_rest, (0, 2, 1))).reshape(len(xyz), -1)
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    dtype_full = [(a, "f4") for a in names]
    elements = np.empty(xyz.shape[0], dtype=dtype_full)
    attributes = np.concatenate((xyz, normals, f_dc, f_rest, opacity, scaling, rotation), axis=1)
    elements[:] = list(map(tuple, attributes))
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(xyz)
    header += "".join("property float %s\n" % a for a in names) + "end_header\n"
    return header.encode("ascii") + elements.tobytes()


def load_ply_file(path, degree):
    """Literal restatement of GaussianModel.load_ply_file over a minimal reader of the layout above."""
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode("ascii").split("\n")
    n = int([ln for ln in lines if ln.startswith("element vertex")][0].split()[-1])
    props = [ln.split()[-1] for ln in lines if ln.startswith("property float")]
    el = np.frombuffer(body, dtype=[(p, "<f4") for p in props], count=n)
    xyz = np.stack((np.asarray(el["x"]), np.asarray(el["y"]), np.asarray(el["z"])), axis=1)
    opacities = np.asarray(el["opacity"])[..., np.newaxis]
    features_dc = np.zeros((xyz.shape[0], 3, 1))
    features_dc[:, 0, 0] = np.asarray(el["f_dc_0"])
    features_dc[:, 1, 0] = np.asarray(el["f_dc_1"])
    features_dc[:, 2, 0] = np.asarray(el["f_dc_2"])
    extra_f_names = sorted([p for p in props if p.startswith("f_rest_")], key=lambda x: int(x.split("_")[-1]))
    assert len(extra_f_names) == 3 * (degree + 1) ** 2 - 3
    features_extra = np.zeros((xyz.shape[0], len(extra_f_names)))
    for idx, attr_name in enumerate(extra_f_names):
        features_extra[:, idx] = np.asarray(el[attr_name])
    features_extra = features_extra.reshape((features_extra.shape[0], 3, (degree + 1) ** 2 - 1))
    scale_names = sorted([p for p in props if p.startswith("scale_")], key=lambda x: int(x.split("_")[-1]))
    scales = np.zeros((xyz.shape[0], len(scale_names)))
    for idx, attr_name in enumerate(scale_names):
        scales[:, idx] = np.asarray(el[attr_name])
    rot_names = sorted([p for p in props if p.startswith("rot")], key=lambda x: int(x.split("_")[-1]))
    rots = np.zeros((xyz.shape[0], len(rot_names)))
    for idx, attr_name in enumerate(rot_names):
        rots[:, idx] = np.asarray(el[attr_name])
    return xyz, features_dc, features_extra, opacities, scales, rots
