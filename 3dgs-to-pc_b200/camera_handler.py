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
    """OpenGL-style perspective matrix with z_sign = +1 (camera_handler.py:14-34)."""
    tan_y = math.tan(fovY / 2)
    tan_x = math.tan(fovX / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right

    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


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
    """Build the per-camera object the renderer consumes (camera_handler.py:53-108).

    cam_intrinsic = [w, h, fx, fy]; the image is scaled so that its width equals colour_resolution unless a mask is
    given (then the native size is kept and the mask is flattened)."""
    diff = 1 if (colour_resolution is None or mask is not None) else colour_resolution / int(cam_intrinsic[0])

    if mask is not None:
        if mask.shape[1] != int(cam_intrinsic[0]) or mask.shape[0] != int(cam_intrinsic[1]):
            raise Exception("Size of mask must match size of input image")
        mask = mask.flatten()

    img_width = int(int(cam_intrinsic[0]) * diff)
    img_height = int(int(cam_intrinsic[1]) * diff)

    focal_x = float(cam_intrinsic[2]) * diff
    focal_y = float(cam_intrinsic[3]) * diff

    if renderer_type == "python":
        cam = Camera(img_width, img_height, focal_x, focal_y, transform)
        cam.sh_degree = sh_degree
        return cam

    elif renderer_type == "cuda":
        from gaussian_pointcloud_rasterization import GaussianRasterizationSettings

        transform[:, 1:3] = -transform[:, 1:3]  # OpenGL -> z-forward, in place like the reference (:75)

        fovX = focal2fov(focal_x, img_width)
        fovY = focal2fov(focal_y, img_height)

        projmatrix = getProjectionMatrix(znear=10, zfar=100, fovX=fovX, fovY=fovY).transpose(0, 1).to(transform.device)
        viewmatrix = torch.linalg.inv(transform).permute(1, 0)
        campos = viewmatrix.inverse()[3, :3]
        bg = torch.ones(3, device=transform.device) if white_bkgd else torch.zeros(3, device=transform.device)

        return GaussianRasterizationSettings(
            image_height=int(img_height),
            image_width=int(img_width),
            tanfovx=math.tan(fovX * 0.5),
            tanfovy=math.tan(fovY * 0.5),
            bg=bg,
            scale_modifier=1.0,
            campos=campos,
            viewmatrix=viewmatrix,
            projmatrix=viewmatrix @ projmatrix,
            sh_degree=sh_degree,
            prefiltered=False,
            mask=mask,
            debug=True,
            antialiasing=False,
        )
