"""Camera model — drop-in for the reference's camera_handler.py (same names and signatures).

Reference: /root/reference/camera_handler.py (fov/focal helpers :8-12, getProjectionMatrix :14-34, Camera :36-50,
get_camera :53-108).  The 4x4 matrix algebra is tiny host-side set-up (one camera per call) and stays in torch; the
matrices are handed to the colour kernels by value (g2pc_camera_t).
"""
import math

import torch


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def getProjectionMatrix(znear, zfar, fovX, fovY):
    """Perspective matrix of the reference (camera_handler.py:14-34): symmetric frustum, +z convention, so that
    w_clip = z_view.  Only five entries are non-zero."""
    half_w = math.tan(fovX / 2) * znear
    half_h = math.tan(fovY / 2) * znear
    depth = zfar - znear
    return torch.tensor([[znear / half_w, 0.0, 0.0, 0.0],
                         [0.0, znear / half_h, 0.0, 0.0],
                         [0.0, 0.0, zfar / depth, -(zfar * znear) / depth],
                         [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)


class Camera():
    """Pinhole camera of the python renderer (camera_handler.py:36-50): OpenGL c2w (looks down -z), row-vector
    view matrix world_view_transform = inv(c2w)^T, znear 10, zfar 100.

    The 4x4 algebra runs on the host in float32 (the kernels take the matrices by value); the tensor attributes of the
    reference (world_view_transform, projection_matrix, camera_center, full_proj_transform) are exposed on c2w's
    device on first access."""

    def __init__(self, width, height, focal_x, focal_y, c2w, znear=10, zfar=100):
        self.znear = znear
        self.zfar = zfar
        self.focal_x = focal_x
        self.focal_y = focal_y
        self.FoVx = focal2fov(self.focal_x, width)
        self.FoVy = focal2fov(self.focal_y, height)
        self.image_width = int(width)
        self.image_height = int(height)
        self.c2w = c2w
        host = c2w.detach().to("cpu", torch.float32)
        self._host = {}
        self._host["world_view_transform"] = torch.linalg.inv(host).permute(1, 0).contiguous()
        self._host["projection_matrix"] = getProjectionMatrix(znear=self.znear, zfar=self.zfar, fovX=self.FoVx,
                                                               fovY=self.FoVy).transpose(0, 1).contiguous()
        self._host["camera_center"] = self._host["world_view_transform"].inverse()[3, :3].contiguous()
        self._host["full_proj_transform"] = self._host["world_view_transform"] @ self._host["projection_matrix"]
        self._dev = {}

    def host(self, name):
        return self._host[name]

    def __getattr__(self, name):
        h = self.__dict__.get("_host", {})
        if name in h:
            d = self.__dict__["_dev"]
            if name not in d:
                d[name] = h[name].to(self.__dict__["c2w"].device)
            return d[name]
        raise AttributeError(name)


def get_camera(renderer_type, transform, cam_intrinsic, colour_resolution=None, sh_degree=3, white_bkgd=True, mask=None):
    """Per-camera object for `renderer(camera)` (camera_handler.py:53-108).

    cam_intrinsic = [w, h, fx, fy].  Without a mask the image is rescaled to `colour_resolution` pixels wide (focal
    lengths scale along); with a mask the native size is kept and the mask must match it (it is returned flattened on
    the camera as `.mask`)."""
    native_w, native_h = int(cam_intrinsic[0]), int(cam_intrinsic[1])
    scale = 1 if (colour_resolution is None or mask is not None) else colour_resolution / native_w
    if mask is not None:
        if tuple(mask.shape[:2]) != (native_h, native_w):
            raise Exception("Size of mask must match size of input image")
        mask = mask.flatten()
    if renderer_type != "python":
        if renderer_type == "cuda":
            raise NotImplementedError("renderer_type='cuda' cameras (z-forward convention, GaussianRasterizationSettings) "
                                      "belong to the not-yet-built 16x16-tile back-end; use renderer_type='python'")
        raise Exception(f"Renderer of type {renderer_type} is not supported")
    cam = Camera(int(native_w * scale), int(native_h * scale), float(cam_intrinsic[2]) * scale,
                 float(cam_intrinsic[3]) * scale, transform)
    cam.sh_degree = sh_degree
    cam.white_bkgd = white_bkgd
    cam.mask = mask
    return cam
