"""Network / renderer hyper-parameters of the reference configs (confs/wmask_realobj_bean.conf:40-77,
confs/wmask_realhand_hand1.conf:40-77, fit_confs/fit_12_8views.conf:26-91), as plain dicts."""
OBJ_SDF_CONF = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], v_multires=10,
                    r_multires=4, bias=0.5, scale=1.0, geometric_init=True, weight_norm=True)
OBJ_COLOR_CONF = dict(d_feature=256, d_in=3, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                      v_multires=10, r_multires=4, grad_multires=4, squeeze_out=True,
                      use_gradients=True)
HAND_SDF_CONF = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], v_multires=10,
                     r_multires=7, bias=0.5, scale=1.0, geometric_init=True, weight_norm=True)
HAND_COLOR_CONF = dict(d_feature=256, d_in=3, d_out=3, d_hidden=256, n_layers=4,
                       weight_norm=True, v_multires=10, r_multires=7, grad_multires=4,
                       squeeze_out=True, use_gradients=True)
RENDERER_CONF = dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0)
VARIANCE_INIT = 0.3
