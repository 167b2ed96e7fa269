"""Drop-in for utils/renderer_batch.py of the reference: the frame-batched hand + object
``NeuSRenderer_fitting`` used by fitting_video.py.  Rays are [F, P, 3], ``bt_inv`` [F,21,4,4],
``T_pose_21`` [F,21,3], ``Ro`` [F,3,3], ``To`` [F,3].  All arithmetic is in libhonerf_b200.so."""
import numpy as np
import torch

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

    def get_stable_loss_cross(self, pts, bt_inv, T_pose_21, Ro, To):
        """utils/renderer_batch.py:318-371: temporal contact-stability loss.  The hand-SDF query runs on
        the device through the fused field; the nearest-neighbour bookkeeping stays on the host with
        scipy's cKDTree exactly as upstream."""
        from scipy import spatial
        pts = pts[:, ::10, :]
        batch_size, p_num, _ = pts.shape
        pts_world = (Ro.unsqueeze(1) @ pts.unsqueeze(-1))[..., 0] + To.unsqueeze(1)
        vert_id_all = range(p_num)
        hand_sdf = self.sdf_network_hand.sdf(pts_world, bt_inv, T_pose_21).reshape(batch_size, p_num, 1)
        hand_sdf_list, in_id_list = [], []
        for batch_id in range(batch_size):
            cur_hand_sdf = hand_sdf[batch_id].reshape(-1)
            penet_id = cur_hand_sdf < 0
            if penet_id.float().sum() > 0:
                in_id_list.append(penet_id)
                hand_sdf_list.append(cur_hand_sdf)
        stable_loss = 0
        if len(in_id_list) > 1:
            hand_sdf_list = torch.stack(hand_sdf_list, 0)
            in_time = hand_sdf_list.shape[0]
            for cid in range(in_time):
                cur_in_id = in_id_list[cid].clone().cpu()
                # upstream passes the BOOLEAN mask to setdiff1d (utils/renderer_batch.py:349), which removes
                # only the ids {0, 1} that the mask's values compare equal to; kept for identical results
                cur_out_id = np.setdiff1d(vert_id_all, cur_in_id.numpy())
                in_points = pts[0, cur_in_id.to(pts.device)].detach().cpu()
                out_points = pts[0, cur_out_id].detach().cpu()
                in_points_num = in_points.shape[0]
                nn_index = spatial.cKDTree(out_points.numpy())
                _, near_out_id = nn_index.query(in_points.numpy(), k=1)
                near_out_id = np.unique(near_out_id.reshape(-1))
                in_mask = cur_in_id.to(hand_sdf_list.device)
                in_err = hand_sdf_list[:, in_mask].clip(0, 1e7).sum() / ((in_time - 1) * in_points_num)
                hand_sdf_select = hand_sdf_list[:, cur_out_id]
                out_err = torch.abs(hand_sdf_select[:, near_out_id].clip(-1e7, 0)).sum() / ((in_time - 1) * in_points_num)
                stable_loss = stable_loss + in_err + 0.05 * out_err
            stable_loss /= in_time
        return stable_loss
