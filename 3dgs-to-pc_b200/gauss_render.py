"""Colour stage — drop-in for the reference's gauss_render.py (get_renderer factory).  [placeholder: filled in next]"""


def get_renderer(renderer_type: str, xyz, opacities, colours, covariances, shs=None, visible_gaussian_threshold=0.0,
                 surface_distance_std=None, calculate_surface_distance=False):
    raise NotImplementedError("colour stage not built yet")
