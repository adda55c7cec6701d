"""Colour stage — drop-in for the reference's gauss_render.py.

Reference: /root/reference/gauss_render.py.  `get_renderer` (:467-493) is the plugin boundary: it returns a callable
`renderer(camera) -> (image, radii|None, invdepth|None, depth|None)` that, as a side effect, keeps for every Gaussian
the largest contribution alpha*T it made to any pixel of any camera and the blended colour of that pixel, plus the
getters used by the pipeline (gauss_to_pc.py:481-513).

`renderer_type="python"` reproduces GaussPythonRenderer (:210-465): quadtree tiles, every Gaussian of a tile blended
into every pixel of the tile — but as sm_100a kernels behind the C ABI (csrc/s3_preprocess.cu, s4_tree.cu,
s5_blend.cu), one camera = 9 entry-point calls + one 32-byte header read.  The tile parameters the reference derives from free
GPU memory at call time (:440-444) are pinned (g2pc.config.MAX_TILE_SIZE / MAX_GAUSSIANS_PER_TILE).
"""
import ctypes
import math

import numpy as np
import torch

from g2pc import capi, config, quadtree

# SH constants kept for API parity with the reference module (gauss_render.py:9-38)
C0 = 0.28209479177387814
C1 = 0.4886025119029199

homogeneous = lambda points: torch.cat([points, torch.ones_like(points[..., :1])], dim=-1)


def strip_lowerdiag(L):
    idx = torch.tensor([0, 1, 2, 4, 5, 8], device=L.device)
    return L.reshape(L.shape[0], 9).index_select(1, idx).to(torch.float)


def strip_symmetric(sym):
    return strip_lowerdiag(sym)


class GaussPythonRenderer():
    """B200 implementation of the reference's pure-torch tile renderer (same constructor arguments, attributes and
    getters as gauss_render.py:210-264)."""

    def __init__(self, means3D, opacity, colour, cov3d, white_bkgd=True, visible_gaussian_threshold=0.0, shs=None,
                 sh_degree=None):
        capi.require_cuda(means3D, opacity, colour, cov3d, shs)
        self.lib = capi.load()
        self.white_bkgd = white_bkgd
        self.device = means3D.device
        n = means3D.shape[0]

        self.gaussian_max_contribution = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.gaussian_total_contribution = torch.zeros(n, device=self.device, dtype=torch.float32)
        # blended colour of each Gaussian's best pixel (f32; the reference keeps f64)
        self.gaussian_colours = torch.zeros((n, 3), device=self.device, dtype=torch.float32)
        self.visible_gaussian_threshold = visible_gaussian_threshold

        self.means3D = means3D.to(torch.float32).contiguous()
        self.opacity = opacity.to(torch.float32).reshape(-1).contiguous()
        self.cov3d = cov3d.to(torch.float32).contiguous()
        self.colour = colour
        self._colour_f32 = None if colour is None else colour.to(torch.float32).contiguous()
        self.shs = None
        self.sh_degree = 0
        if shs is not None:
            self.shs = shs.to(torch.float32).contiguous()
            ncoef = self.shs.shape[-1]
            deg = int(round(math.sqrt(ncoef))) - 1 if sh_degree is None else int(sh_degree)
            self.sh_degree = min(deg, 3)

        self.max_tile_size = config.MAX_TILE_SIZE
        self.max_gaussians_per_tile = config.MAX_GAUSSIANS_PER_TILE
        self.compose_image = True
        self._tables = {}
        self._extra_levels = 0
        self._n = n
        # per-camera scratch, allocated once
        self._proj = torch.empty((n, 12), dtype=torch.float32, device=self.device)
        self._cam_best = torch.zeros((n,), dtype=torch.int64, device=self.device)
        self._depth_key = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._touched = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._order = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._incl = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._depth_ws = None
        self._inst_leaf = self._inst_leaf_alt = self._inst_gid = self._inst_gid_alt = None
        self._leaf_colour = None
        self._sort_ws = None
        self._hdr_host = torch.zeros((capi.HDR_WORDS,), dtype=torch.int32).pin_memory()
        self.last_stats = {}

    # ---- getters (gauss_render.py:237-264) -----------------------------------------------------------------
    def get_gaussian_colours(self):
        return self.gaussian_colours * 255

    def get_gaussians_above_contribution_threshold(self, contribution_threshold):
        return self.gaussian_max_contribution > contribution_threshold

    def get_visible_gaussians(self):
        return self.get_gaussians_above_contribution_threshold(self.visible_gaussian_threshold)

    def get_surface_gaussians(self):
        return self.get_gaussians_above_contribution_threshold(torch.mean(self.gaussian_max_contribution))

    def get_total_gaussian_contributions(self):
        # the python back-end of the reference reports the MAX contribution here (gauss_render.py:261-264)
        return self.gaussian_max_contribution

    # ---- per-resolution tables ---------------------------------------------------------------------------------
    def _get_tables(self, W, H):
        key = (W, H, self.max_tile_size, self.max_gaussians_per_tile, self._extra_levels)
        t = self._tables.get(key)
        if t is None:
            qt = quadtree.QuadtreeTables(W, H, self.max_tile_size, self.max_gaussians_per_tile,
                                         extra_levels=self._extra_levels)
            flat = np.concatenate(qt.flat()).astype(np.int32)
            dev = self.device
            t = dict(qt=qt, tables=torch.from_numpy(flat).to(dev),
                     node_cnt=torch.zeros((qt.nodes_2d + 1,), dtype=torch.int32, device=dev),
                     node_state=torch.zeros((qt.nodes_2d,), dtype=torch.uint8, device=dev),
                     leaf_of_node=torch.full((qt.nodes_2d,), -1, dtype=torch.int32, device=dev),
                     leaves=torch.zeros((qt.nodes_2d, capi.LEAF_WORDS), dtype=torch.int32, device=dev),
                     seg_begin=torch.zeros((qt.nodes_2d + 1,), dtype=torch.int32, device=dev),
                     leaf_order=torch.zeros((qt.nodes_2d,), dtype=torch.int32, device=dev),
                     header=torch.zeros((capi.HDR_WORDS,), dtype=torch.int32, device=dev),
                     owner=torch.zeros((W * H,), dtype=torch.int32, device=dev),
                     image=torch.ones((H, W, 3), dtype=torch.float32, device=dev),
                     max_quads=int(max(((int(w) + 3) // 4) * int(h)
                                       for w in [min(self.max_tile_size, W)] for h in [min(self.max_tile_size, H)])))
            t["work_counter"] = t["node_cnt"][qt.nodes_2d:]  # zeroed together with the node counts
            self._tables[key] = t
        return t

    @staticmethod
    def _camera_struct(camera):
        c = capi.Camera()
        if hasattr(camera, "host"):  # matrices already on the host (camera_handler.Camera)
            get = camera.host
        else:
            get = lambda k: getattr(camera, k).detach().to("cpu", torch.float32)
        V = get("world_view_transform").contiguous().reshape(-1).tolist()
        P = get("projection_matrix").contiguous().reshape(-1).tolist()
        pos = get("camera_center").reshape(-1).tolist()
        for i in range(16):
            c.view[i] = V[i]
            c.proj[i] = P[i]
        for i in range(3):
            c.campos[i] = pos[i]
        c.tan_fovx = math.tan(camera.FoVx * 0.5)
        c.tan_fovy = math.tan(camera.FoVy * 0.5)
        c.focal_x = camera.focal_x
        c.focal_y = camera.focal_y
        c.width = camera.image_width
        c.height = camera.image_height
        return c

    def _grow(self, name, numel, dtype):
        buf = getattr(self, name)
        if buf is None or buf.numel() < numel:
            buf = torch.empty((int(numel * 1.25) + 1024,), dtype=dtype, device=self.device)
            setattr(self, name, buf)
        return buf

    def __call__(self, camera, **kwargs):
        """Render one camera and update the per-Gaussian accumulators (gauss_render.py:404-465).
        Returns (image (H,W,3) f32 flipped left-right | None, None, None, None)."""
        lib = self.lib
        st = capi.stream_ptr(self.device)
        W, H = int(camera.image_width), int(camera.image_height)
        cam = self._camera_struct(camera)
        n = self._n
        if self._depth_ws is None:
            nbytes = lib.g2pc_depth_order_workspace_bytes(n)
            self._depth_ws = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=self.device)
        while True:
            t = self._get_tables(W, H)
            qt = t["qt"]
            t["node_cnt"].zero_()  # (work_counter is the last word of this buffer)
            capi.call("g2pc_preprocess", capi.ptr(self.means3D), capi.ptr(self.cov3d), capi.ptr(self.opacity),
                      capi.ptr(self._colour_f32) if self.shs is None else None, capi.ptr(self.shs),
                      int(self.shs.shape[-1]) if self.shs is not None else 0, self.sh_degree, n, ctypes.byref(cam),
                      capi.ptr(t["tables"]), qt.num_levels, qt.max_gaussians_per_tile, capi.ptr(self._proj),
                      capi.ptr(t["node_cnt"]), capi.ptr(self._depth_key), capi.ptr(self._touched), st)
            capi.call("g2pc_depth_order", capi.ptr(self._depth_key), capi.ptr(self._touched), n,
                      capi.ptr(self._order), capi.ptr(self._incl), capi.ptr(self._depth_ws), self._depth_ws.numel(), st)
            capi.call("g2pc_build_tree", capi.ptr(t["tables"]), qt.num_levels, qt.max_gaussians_per_tile,
                      capi.ptr(t["node_cnt"]), capi.ptr(self._incl), n, capi.ptr(t["node_state"]),
                      capi.ptr(t["leaf_of_node"]), capi.ptr(t["leaves"]), capi.ptr(t["seg_begin"]),
                      capi.ptr(t["leaf_order"]), qt.nodes_2d, capi.ptr(t["header"]), st)
            self._hdr_host.copy_(t["header"], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()  # the one host read per camera (32 bytes)
            hdr = self._hdr_host.tolist()
            if hdr[capi.HDR_NEED_DEEPER]:
                # a tile at the deepest tabulated level holds more than max_gaussians_per_tile Gaussians: tabulate one
                # more level and redo this camera (rare; the reference keeps splitting in its host BFS)
                self._extra_levels += 1
                if self._get_tables(W, H)["qt"].num_levels <= qt.num_levels:
                    raise capi.G2pcError(
                        f"a tile still holds more than max_gaussians_per_tile={self.max_gaussians_per_tile} Gaussians at "
                        f"quadtree level {qt.num_levels - 1} (tiles of a few pixels): deeper than the tabulated levels")
                continue
            if hdr[capi.HDR_LEAF_OVERFLOW]:
                raise capi.G2pcError("leaf table overflow")
            break
        num_leaves, total_inst, total_pix = hdr[capi.HDR_NUM_LEAVES], hdr[capi.HDR_TOTAL_INST], hdr[capi.HDR_TOTAL_PIX]
        total_upper = hdr[capi.HDR_TOTAL_UPPER]
        self.last_stats = dict(num_leaves=num_leaves, total_instances=total_inst, total_leaf_pixels=total_pix,
                               levels=qt.num_levels, instance_slots=total_upper)
        bg = 1.0 if self.white_bkgd else 0.0
        if num_leaves > 0 and total_inst > 0:
            for name in ("_inst_leaf", "_inst_leaf_alt", "_inst_gid", "_inst_gid_alt"):
                self._grow(name, total_upper, torch.int32)
            leaf_colour = self._grow("_leaf_colour", total_pix * 3, torch.float32)
            capi.call("g2pc_emit_instances", capi.ptr(self._proj), capi.ptr(self._order), capi.ptr(self._incl),
                      capi.ptr(self._touched), n, W, H, capi.ptr(t["tables"]), qt.num_levels, qt.candidate_level_mask(),
                      capi.ptr(t["node_state"]), capi.ptr(t["leaf_of_node"]), capi.ptr(self._inst_leaf),
                      capi.ptr(self._inst_gid), st)
            ws_bytes = lib.g2pc_sort_instances_workspace_bytes(total_upper)
            if ws_bytes < 0:
                raise capi.G2pcError("cub workspace query failed")
            ws = self._grow("_sort_ws", max(ws_bytes, 1), torch.uint8)
            in_alt = ctypes.c_int32(0)
            # padding entries carry leaf id 0xFFFFFFFF: sorting on bit_length(num_leaves) bits puts them last
            leaf_bits = max(1, int(num_leaves).bit_length())
            capi.call("g2pc_sort_instances", capi.ptr(self._inst_leaf), capi.ptr(self._inst_leaf_alt),
                      capi.ptr(self._inst_gid), capi.ptr(self._inst_gid_alt), total_upper, leaf_bits, capi.ptr(ws),
                      ws.numel(), ctypes.byref(in_alt), st)
            sorted_gid = self._inst_gid_alt if in_alt.value else self._inst_gid
            self._last_sorted_gid = sorted_gid
            capi.call("g2pc_blend", capi.ptr(t["leaves"]), capi.ptr(t["leaf_order"]), num_leaves, t["max_quads"],
                      capi.ptr(sorted_gid), capi.ptr(self._proj), capi.ptr(self._cam_best),
                      capi.ptr(self.gaussian_max_contribution), capi.ptr(leaf_colour), capi.ptr(t["owner"]), W, H, bg,
                      capi.ptr(t["work_counter"]), st)
            capi.call("g2pc_accumulate", capi.ptr(self._cam_best), capi.ptr(leaf_colour), n,
                      capi.ptr(self.gaussian_max_contribution), capi.ptr(self.gaussian_colours), st)
            if self.compose_image:
                capi.call("g2pc_compose_image", capi.ptr(t["owner"]), capi.ptr(leaf_colour), W, H, bg,
                          capi.ptr(t["image"]), st)
            else:
                t["owner"].zero_()
        elif self.compose_image:
            t["image"].fill_(bg)
        self._last_tables = t
        return (t["image"] if self.compose_image else None), None, None, None


    # ---- introspection for the parity tests ---------------------------------------------------------------------
    def debug_last_camera(self):
        """Per-Gaussian projection records and per-leaf sorted Gaussian ids of the most recent camera (host copies)."""
        t = self._last_tables
        nl = self.last_stats["num_leaves"]
        leaves = t["leaves"][:nl].cpu().numpy()
        gids = (self._last_sorted_gid[: self.last_stats["total_instances"]].cpu().numpy().astype(np.int64)
                if nl else np.zeros(0, np.int64))
        out = []
        for (r0, c0, w, h, beg, cnt, pix, node) in leaves:
            out.append((int(r0), int(c0), int(w), int(h), gids[beg:beg + cnt]))
        return self._proj.cpu().numpy(), out


def get_renderer(renderer_type: str, xyz, opacities, colours, covariances, shs=None, visible_gaussian_threshold=0.0,
                 surface_distance_std=None, calculate_surface_distance=False):
    """Factory with the reference's signature (gauss_render.py:467-493)."""
    if renderer_type == "python":
        return GaussPythonRenderer(xyz, opacities.type(torch.float), colours if shs is None else None, covariances,
                                   visible_gaussian_threshold=visible_gaussian_threshold, shs=shs)
    if renderer_type == "cuda":
        raise NotImplementedError(
            "renderer_type='cuda' (16x16-tile semantics of the reference's CUDA extension, incl. surface distances) is "
            "the next row of the scope table (SURVEY.md §8f N2) and is not built yet; use renderer_type='python', "
            "which runs the python renderer's semantics as B200 kernels")
    raise Exception(f"Renderer of type {renderer_type} is not supported")
