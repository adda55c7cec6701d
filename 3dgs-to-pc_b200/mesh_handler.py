"""Open3D post-processing hooks (reference: mesh_handler.py).  OUT OF SCOPE of this build (third-
party CPU library); the names exist so that `--clean_pointcloud` / `--generate_mesh` fail with a clear message."""


def _need_open3d():
    try:
        import open3d  # noqa: F401
    except ImportError as e:
        raise ImportError("Open3D is required for point-cloud cleaning / meshing and is not part of g2pc") from e
    raise NotImplementedError("Open3D cleaning / meshing is outside the scope of the g2pc hot path")


def clean_point_cloud(points, colours, normals, device="cuda:0"):
    _need_open3d()


def generate_mesh(points, colours, normals, output_path, depth=10, laplacian_iters=10):
    _need_open3d()
