"""GPU: parity against the oracle at BASELINE configs[2] SIZE -- 2048 pixels x 64 rays, training mode, 128 SGs, forward + loss
+ backward -- on the rough scene (SDF perturbation 0.08), at the library's shipped defaults.  Both sides run on the same
device with identical weights, rays and random numbers (tests/parity_util.py).

north_star tolerances and what is asserted here (measured values in the comments; `tools/diag_gpu.py fullsize` prints them):
  * hit masks: bit-exact is only defined for an analytic SDF (tests/test_tracer_gpu.py).  With the MLP in the loop the two
    sides evaluate the SDF with different arithmetic (tcgen05 fp16-split products vs cuBLAS SGEMM, ~4e-7 apart; torch's own
    fp32 result is 1.5e-7 from the f64 value); measured 0 mismatching rays of 32768 and 0 mismatching pixels of 2048 --
    asserted as such.
  * depth abs 1e-4: asserted on >= 99 % of the agreeing hits (measured 99.8 - 100 %; median 1.2e-7).  The rest are rays that
    graze a bump, where the first of two nearby sign changes is a different 1/100-sample bracket on the two sides.
  * gradients rel 1e-3 (lgtSGs, material MLP, radiance MLP): asserted (measured <= 3.5e-4).
  * shading rel 1e-4 (abs floor 1e-6): reached on 98.4 - 99.2 % of the per-pixel lanes (the figure moves by +-0.4 % with
    any change of the rounding order, e.g. the accuracy tier: a pixel is the mean of 64 primary x 3 secondary rays, and one
    grazing secondary ray that brackets differently moves it), p95 1.7e-5, p99 4e-4.  Asserted at >= 97.5 % with p95 <= 5e-5.
    Round 1 / early round 2 (bf16-split planes, SDF error 1.9e-6): 96 %, p95 4.3e-5.
"""
import pytest
import torch

from tests.parity_util import fullsize_compare

pytestmark = pytest.mark.gpu


def test_training_step_at_configs2_size(cuda_device):
    r = fullsize_compare(cuda_device, bumps=0.08, n_px=2048, n_rays=64, training=True, grads=True, verbose=True, ref64=True)
    assert r['pixels'] == 2048 and r['hits'] > 300
    assert r['mask_mismatch'] == 0, r['mask_mismatch']                 # per pixel: all 64 rays of the pixel agree
    assert r['depth'][0] < 5e-7 and r['depth'][1] < 3e-6, r['depth']   # median / p95 of |d points| on hits (pixel means)
    assert r['depth_frac_1e4'] >= 0.99, r['depth_frac_1e4']            # north_star: depth abs 1e-4 (measured: 0 - 2 of 484 pixels beyond)
    assert r['sdf_output_hit'][2] < 3e-6, r['sdf_output_hit']
    k = r['keys']
    assert k['sg_rgb_values']['frac_1e4'] >= 0.975, k['sg_rgb_values']        # measured 0.984 - 0.992
    assert k['sg_rgb_values']['q'][1] <= 5e-5, k['sg_rgb_values']      # p95 well within the north_star tolerance
    assert k['sg_diffuse_rgb_values']['frac_1e4'] >= 0.975                    # measured 0.985
    assert k['sg_roughness_values']['frac_1e4'] >= 0.999 and k['sg_diffuse_albedo_values']['frac_1e4'] >= 0.999
    assert k['normal_values']['absq'][2] < 5e-5, k['normal_values']    # unit vectors: absolute, p99 (measured 1.4e-5)
    assert r['background_rel'][3] < 1e-5                               # environment lookup on miss rays: max rel
    assert r['secondary_mismatch'] is not None and r['secondary_mismatch'] <= 1e-4 * r['secondary_rays'], r['secondary_mismatch']
    # SURVEY 8(d) noise floor: against the SAME oracle run in float64, next to the fp32 oracle's own distance from it.  Measured:
    # sg_rgb lanes within 1e-4 of the float64 result: ours 98.5 %, the fp32 oracle 99.2 % (p95 1.6e-5 / 6.4e-6); idr_rgb (the
    # radiance MLP behind a 2^9-frequency encoding of the hit point) 71 % / 91 %; hit masks: 0 mismatches for both.
    f = r['f64']
    assert f['mask_mismatch_ours'] == f['mask_mismatch_ref32'] == 0, f
    assert f['keys']['sg_rgb_values']['ours'][0] >= f['keys']['sg_rgb_values']['ref32'][0] - 0.02, f['keys']['sg_rgb_values']
    assert f['keys']['sg_rgb_values']['ours'][2] <= 5e-5, f['keys']['sg_rgb_values']        # p95 vs float64
    assert f['keys']['points']['ours'][0] >= 0.995, f['keys']['points']
    # north_star: gradients within rel 1e-3
    assert r['g_lgt'] < 1e-3, r['g_lgt']
    assert max(r['g_mat']) < 1e-3, r['g_mat']
    assert max(r['g_rad']) < 1e-3, r['g_rad']


def test_per_ray_lanes_at_32768_rays(cuda_device):
    """the same rays as single-ray pixels (no averaging over the 64 rays of a pixel): masks ray by ray"""
    r = fullsize_compare(cuda_device, bumps=0.08, n_px=32768, n_rays=0, training=True, grads=False, verbose=True)
    assert r['hits'] > 5000
    assert r['mask_mismatch'] == 0, r['mask_mismatch']
    assert r['depth_frac_1e4'] >= 0.999, r['depth_frac_1e4']            # measured 0.99988
    assert r['depth'][0] < 5e-7 and r['depth'][2] < 1e-5, r['depth']
    assert r['keys']['sg_rgb_values']['frac_1e4'] >= 0.95, r['keys']['sg_rgb_values']      # measured 0.964 (single rays)
    assert r['keys']['sg_rgb_values']['q'][0] < 1e-5


def test_eval_render_at_8192_rays(cuda_device):
    r = fullsize_compare(cuda_device, bumps=0.08, n_px=8192, n_rays=0, training=False, grads=False, verbose=True)
    assert r['mask_mismatch'] == 0
    assert r['depth_frac_1e4'] >= 0.999
    assert r['keys']['sg_rgb_values']['frac_1e4'] >= 0.985, r['keys']['sg_rgb_values']     # measured 0.993
    assert r['keys']['sg_rgb_values']['q'][1] < 5e-5


def test_fitted_concave_scene_at_configs2_size(cuda_device):
    """SURVEY 8d cfg 3 geometry: the SDF network L1-fitted (seeded, 400 Adam steps of the step-1 path on the trainable tcgen05 stack)
    to the analytic robot-scale union of boxes and spheres -- concave, thin limbs, many sampler / bisection rays -- then frozen;
    the same full-size training-mode comparison against the oracle holding exactly the fitted weights."""
    import bench
    fit, _, l1 = bench.fit_geometry(cuda_device, 400)
    assert l1 < 0.05, l1
    fit.freeze_geometry()
    fit.train()
    r = fullsize_compare(cuda_device, bumps=0.0, n_px=2048, n_rays=64, training=True, grads=True, verbose=True, model=fit)
    # measured: 920 hit pixels, 0 mask mismatches (2 of 177 741 secondary rays), depth median 1.3e-7 / 100 % within 1e-4,
    # sg_rgb 98.6 % of lanes within 1e-4 (p95 1.2e-5), gradients 2e-5 / 4e-5 / 1.4e-4.  The fit runs on the GPU and is not
    # bit-reproducible (atomic column sums): thresholds leave room for a slightly different network.
    assert r['hits'] > 400
    assert r['mask_mismatch'] <= 1, r['mask_mismatch']
    assert r['depth'][0] < 5e-7 and r['depth_frac_1e4'] >= 0.99, (r['depth'], r['depth_frac_1e4'])
    assert r['keys']['sg_rgb_values']['q'][1] <= 5e-5, r['keys']['sg_rgb_values']
    assert r['keys']['sg_rgb_values']['frac_1e4'] >= 0.96, r['keys']['sg_rgb_values']
    assert r['secondary_mismatch'] <= 1e-4 * r['secondary_rays']
    assert r['g_lgt'] < 1e-3 and max(r['g_mat']) < 1e-3 and max(r['g_rad']) < 1e-3, (r['g_lgt'], r['g_mat'], r['g_rad'])
