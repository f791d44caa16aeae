"""hidegs_b200 — B200-native (sm_100a) implementation of HiDeGS's differentiable
Gaussian-splatting hot path behind the reference's own Python API.

Sub-packages / modules:
  diff_gaussian_rasterization  drop-in for submodules/hierarchy-rasterizer's
                               Python package (GaussianRasterizationSettings,
                               GaussianRasterizer, rasterize_gaussians, _C)
  _lib                         ctypes binding of the C-ABI in include/*.h
  build                        in-tree nvcc build of libhidegs_b200.so
"""
__version__ = "0.1.0"


def install(patch_reference_modules=True):
    """Route the reference's imports to this package (call once, before importing the reference's modules):

        sys.modules: diff_gaussian_rasterization(+._C), simple_knn(+._C)  ->  hidegs_b200's drop-ins
        if importable and `patch_reference_modules`: the reference's own `utils.loss_utils` (l1_loss, l2_loss, ssim,
        get_img_grad_weight, lncc), `scripts.frequency_regularization.frequency_regularization_pyramid_scale`,
        `gaussian_renderer` (render, render_post, render_normal) and `scene.OurAdam.Adam` get their hot functions
        replaced in place, so `from utils.loss_utils import ssim` etc. keep working unchanged.

    Returns the list of names that were redirected."""
    import importlib
    import sys
    from . import diff_gaussian_rasterization as dgr, simple_knn as knn
    done = []
    for name, mod in (("diff_gaussian_rasterization", dgr), ("diff_gaussian_rasterization._C", dgr._C),
                      ("simple_knn", knn), ("simple_knn._C", knn._C)):
        sys.modules[name] = mod
        done.append(name)
    if not patch_reference_modules:
        return done
    from . import loss_utils as lu, frequency_regularization as fr, gaussian_renderer as gr, optim
    targets = (("utils.loss_utils", lu, ("l1_loss", "l2_loss", "ssim", "get_img_grad_weight", "lncc")),
               ("scripts.frequency_regularization", fr, ("frequency_regularization_pyramid_scale",)),
               ("gaussian_renderer", gr, ("render", "render_post", "render_normal")),
               ("scene.OurAdam", optim, ("Adam",)))
    for modname, src, names in targets:
        try:
            mod = importlib.import_module(modname)
        except Exception:  # the reference is not on sys.path (or its own imports are unavailable): nothing to patch
            continue
        if getattr(mod, "__file__", "") and "hidegs_b200" in (mod.__file__ or ""):
            continue
        for n in names:
            setattr(mod, n, getattr(src, n))
            done.append(modname + "." + n)
    return done
