"""Top-level alias so that `from gaussian_renderer import render, render_post, render_coarse` (as the reference's training / eval
scripts do, eval.py:24) resolves to the B200 implementation when this repository is on sys.path."""
from hidegs_b200.gaussian_renderer import (  # noqa: F401
    render, render_post, render_coarse, render_normal, geometry_all_map, normal_consistency_loss, camera_intrinsics)
