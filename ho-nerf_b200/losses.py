"""Caller-side losses of the reference as fused kernels.  Same arithmetic as the drivers'
inline code: exp_runner.py:206-227 (training), fitting_single.py:253-283, fitting_video.py:286-309 (fitting)."""
from . import ops


def training_loss(render_out, true_rgb, true_mask, igr_weight=1.0, mask_weight=1.0, return_stats=False):
    """exp_runner.py:206-227 without the VGG term; true_mask is thresholded like :205.  stats = [color_fine_loss,
    mask_loss, psnr] stay on the device (no host sync)."""
    mask = (true_mask > 0.5).float()
    loss, stats = ops.render_loss(render_out["color_fine"], render_out["weight_sum"], true_rgb, mask,
                                  render_out["gradient_error"], 0.0, 1.0, mask_weight, igr_weight)
    return (loss, stats) if return_stats else loss


def fitting_render_loss(render_out, true_rgb, true_mask, scale=1.0, return_stats=False):
    """fitting_single.py:253-256: color L1 / n_rays + 0.5 BCE  (scale = 0.5 gives fitting_video.py:287-291, whose
    F * P divisor equals the flattened ray count)."""
    n = render_out["weight_sum"].numel()
    loss, stats = ops.render_loss(render_out["color_fine"], render_out["weight_sum"], true_rgb, true_mask, None,
                                  float(n), scale, 0.5 * scale, 0.0)
    return (loss, stats) if return_stats else loss


def interaction_loss(render_out, w_contact=30.0, w_penet=20.0, return_stats=False):
    """fitting_single.py:268-283: 30 * contact_loss + 20 * penet_loss on the per-sample SDFs of both fields."""
    loss, stats = ops.interaction_loss(render_out["sdf_hand"], render_out["sdf_obj"], 1e-2, w_contact, w_penet)
    return (loss, stats) if return_stats else loss
