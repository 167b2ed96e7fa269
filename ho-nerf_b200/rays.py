"""Ray generation on the device: the reference's `_xy_to_ray_bundle` (utils/utils.py:31-115) over
pytorch3d `PerspectiveCameras`, restated from the NDC pinhole model (pytorch3d itself is not part of the reference tree)."""
import torch

from . import ops


class PerspectiveCameras:
    """The four arguments the reference passes to pytorch3d's class of this name (exp_runner.py:201-202,336-337)."""

    def __init__(self, R, T, focal_length, principal_point):
        self.record = ops.pack_cameras(R, T, focal_length, principal_point)

    def to(self, device):
        self.record = self.record.to(device)
        return self


class RayBundle:
    def __init__(self, origins, directions, lengths, xys):
        self.origins, self.directions, self.lengths, self.xys = origins, directions, lengths, xys


def _xy_to_ray_bundle(cameras, xy_grid, min_depth, max_depth, n_pts_per_ray, unit_directions=True,
                      stratified_sampling=False):
    """utils/utils.py:31-115.  Only the reference's own use is supported: unit directions, no stratified jitter."""
    if not unit_directions or stratified_sampling:
        raise NotImplementedError("the reference calls _xy_to_ray_bundle with unit directions and no jitter only")
    o, d = ops.rays_from_ndc(xy_grid, cameras.record)
    lengths = xy_grid.new_empty((0,))
    if n_pts_per_ray > 0:
        depths = torch.linspace(min_depth, max_depth, n_pts_per_ray, dtype=xy_grid.dtype, device=xy_grid.device)
        lengths = depths.expand(*xy_grid.shape[:-1], n_pts_per_ray)
    return RayBundle(o, d, lengths, xy_grid)


def image_ray_chunks(cameras, H, W, batch_size):
    """Full-image rays of exp_runner.py:338-355 as a generator of (rays_o, rays_d) chunks of `batch_size` pixels; the
    [H*W, 3] ray list is never materialised."""
    xs, ys = ops.ndc_grid_axes(H, W, cameras.record.device)
    for first in range(0, H * W, batch_size):
        yield ops.rays_ndc_grid(xs, ys, cameras.record[0], first, min(batch_size, H * W - first))
