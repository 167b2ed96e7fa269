"""Drop-in for utils/renderer_batch.py of the reference: the frame-batched hand + object
``NeuSRenderer_fitting`` used by fitting_video.py.  Rays are [F, P, 3], ``bt_inv`` [F,21,4,4],
``T_pose_21`` [F,21,3], ``Ro`` [F,3,3], ``To`` [F,3].  All arithmetic is in libhonerf_b200.so."""
import torch

from . import ops
from .renderer import _FittingBase


class NeuSRenderer_fitting(_FittingBase):
    """utils/renderer_batch.py:41-371."""

    def convert_obj_to_local(self, rays_o, rays_d, Ro, To):
        """utils/renderer_batch.py:176-182."""
        rays_o = rays_o - To.unsqueeze(1)
        rays_o = torch.matmul(Ro.unsqueeze(1), rays_o.unsqueeze(-1))[..., -1]
        rays_d = torch.matmul(Ro.unsqueeze(1), rays_d.unsqueeze(-1))[..., -1]
        return rays_o, rays_d

    def render(self, rays_o, rays_d, near, far, bt_inv, T_pose_21, verts, Ro, To, get_SDF=False):
        """utils/renderer_batch.py:184-281."""
        self.batch_size, self.pixel_sample, _ = rays_o.shape
        return self._render(rays_o, rays_d, near, far, bt_inv, T_pose_21, Ro, To)

    def _grid_hand_pts(self, pts):
        return pts.reshape(1, -1, 3)

    def _grid_obj_pts(self, pts, Ro, To):
        obj_pts = pts - To
        return torch.matmul(Ro, obj_pts.unsqueeze(-1))[..., 0]

    def get_stable_loss_cross(self, pts, bt_inv, T_pose_21, Ro, To, fixed=False):
        """utils/renderer_batch.py:318-371: temporal contact-stability loss.  The hand-SDF query runs through the
        fused field; the frame filter, in/out sets, nearest-neighbour selection (hn_nn_select instead of a scipy
        cKDTree per frame) and the sums stay on the device (ops.stable_loss_from_sdf) -- no host round trip.
        Returns a 0-d tensor (0.0 where upstream returns the int 0: fewer than two penetrating frames)."""
        pts = pts[:, ::10, :]
        batch_size, p_num, _ = pts.shape
        pts_world = (Ro.unsqueeze(1) @ pts.unsqueeze(-1))[..., 0] + To.unsqueeze(1)
        hand_sdf = self.sdf_network_hand.sdf(pts_world, bt_inv, T_pose_21).reshape(batch_size, p_num)
        return ops.stable_loss_from_sdf(hand_sdf, pts[0], fixed=fixed)
