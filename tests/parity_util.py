"""Full-size parity harness (TEST INFRASTRUCTURE): IDRNetwork.forward_with_uv of nefii_b200 against oracle/pipeline.py on the
same device, with identical weights, rays and random numbers, at BASELINE configs[2] size.  Used by
tests/test_parity_fullsize_gpu.py (assertions) and tools/diag_gpu.py fullsize (printout)."""
import time

import torch


def _chunked(fn, chunk=1 << 21):
    """evaluate an oracle callable in row chunks (bounds the torch-eager activation memory at full size)"""
    def g(x):
        if x.shape[0] <= chunk:
            return fn(x)
        return torch.cat([fn(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)], 0)
    return g


def _quant(x, qs=(0.5, 0.95, 0.99, 1.0)):
    x = x.flatten().float()
    if x.numel() == 0:
        return tuple(float("nan") for _ in qs)
    srt = x.sort()[0]
    return tuple(srt[min(srt.numel() - 1, int(q * (srt.numel() - 1) + 0.5))].item() for q in qs)


def fullsize_compare(dev, bumps, n_px, n_rays, training, tiers=None, grads=True, seed=0, verbose=True, ref64=False, model=None):
    """BASELINE configs[2]-sized parity run: IDRNetwork.forward_with_uv against oracle/pipeline.py on the same device, same
    weights, same uniforms.  Returns a dict of statistics (used by tests/test_parity_fullsize_gpu.py and printed here)."""
    import bench
    from oracle import pipeline, ref_harness as rh
    from nefii_b200 import _lib
    from nefii_b200.model.implicit_differentiable_renderer import IDRNetwork
    from nefii_b200.utils.conf import default_model_conf
    if tiers is not None:
        _lib.check(_lib.raw().nefii_trace_set_tiers(int(tiers[0]), int(tiers[1])))
    if model is None:
        om = rh.small_model(seed=seed, bumps=bumps)
        torch.manual_seed(0)
        net = IDRNetwork(default_model_conf()).to(dev)
        rh.load_oracle_weights(net, om)
    else:          # a prepared IDRNetwork (e.g. a fitted geometry): the oracle gets exactly its weights
        net = model
        om = bench.oracle_from_net(net)
    om = om.to(dev)
    om.sdf_fn = _chunked(om.sdf_fn)
    net.train(training)
    uv, obj, rgb = bench.make_batch(seed + 11, num_pixels=n_px, num_rays=max(n_rays, 1))
    if n_rays == 0:
        uv = uv[:, :, 0]
    pose, K = bench.make_camera()
    obj = obj.clone()
    obj[0, ::9] = False
    g = torch.Generator().manual_seed(seed + 100)
    U = torch.rand(n_px * max(n_rays, 1), 7, generator=g).to(dev)
    vecs = [torch.rand(100, generator=g) for _ in range(2)]
    inp = dict(uv=uv.to(dev), pose=pose.to(dev), intrinsics=K.to(dev), object_mask=obj.to(dev))
    gt = rgb.to(dev)
    res = {}

    def loss_of(out):
        m = out['network_object_mask'] & out['object_mask']
        bg = (~out['network_object_mask']) & (~out['object_mask'])
        l = (out['sg_rgb_values'][m] - gt[m]).abs().mean() + (out['idr_rgb_values'][m] - gt[m]).abs().mean()
        if bool(bg.any()):
            l = l + ((out['sg_rgb_values'][bg] - gt[bg]) ** 2).mean()
        return l

    for p in net.parameters():
        p.grad = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.set_grad_enabled(grads):
        mine = net.forward_with_uv(inp, uniforms=U, trace_uniforms=vecs[0])
        if grads:
            loss_of(mine).backward()
    torch.cuda.synchronize()
    res['ours_s'] = time.perf_counter() - t0
    if grads:
        om.lgtSGs.requires_grad_(True)
        om.material.requires_grad_(True)
        om.radiance.requires_grad_(True)
    t0 = time.perf_counter()
    with torch.set_grad_enabled(grads):
        ref = pipeline.forward_with_uv(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda n: U[:n], training,
                                       vecs[0], vecs[1])
        if grads:
            loss_of(ref).backward()
    torch.cuda.synchronize()
    res['oracle_s'] = time.perf_counter() - t0

    a, b = mine['network_object_mask'], ref['network_object_mask']
    agree = a == b
    hit = agree & a
    res['pixels'] = int(a.numel())
    res['mask_mismatch'] = int((~agree).sum())
    res['hits'] = int(hit.sum())
    dp = (mine['points'] - ref['points'])[hit].abs().amax(-1)
    res['depth'] = _quant(dp)
    res['depth_frac_1e4'] = (dp < 1e-4).float().mean().item() if dp.numel() else 1.0
    miss = agree & ~a
    res['sdf_output_hit'] = _quant((mine['sdf_output'] - ref['sdf_output'])[hit].abs())
    res['sdf_output_miss'] = _quant((mine['sdf_output'] - ref['sdf_output'])[miss].abs()) if bool(miss.any()) else None
    res['keys'] = {}
    for k in ('normal_values', 'idr_rgb_values', 'sg_rgb_values', 'sg_diffuse_rgb_values', 'sg_specular_rgb_values',
              'sg_roughness_values', 'sg_diffuse_albedo_values'):
        x, y = mine[k][hit].float(), ref[k][hit].float()
        rel = ((x - y).abs() / (y.abs() + 1e-6)).flatten()
        res['keys'][k] = dict(q=_quant(rel), frac_1e4=(rel <= 1e-4).float().mean().item(), absq=_quant((x - y).abs()))
    bgm = agree & ~a & ~mine['object_mask']
    if bool(bgm.any()):
        x, y = mine['sg_rgb_values'][bgm], ref['sg_rgb_values'][bgm]
        res['background_rel'] = _quant(((x - y).abs() / (y.abs() + 1e-6)))
    if mine['secondary_mask'] is not None and ref['secondary_mask'] is not None and mine['secondary_mask'].shape == ref['secondary_mask'].shape:
        sm, sr = mine['secondary_mask'], ref['secondary_mask']
        res['secondary_rays'] = int(sm.numel())
        res['secondary_mismatch'] = int((sm != sr).sum())
        both = (sm & sr).reshape(3, -1)
        d2 = (mine['secondary_points'] - ref['secondary_points']).abs().amax(-1)[both]
        res['secondary_depth'] = _quant(d2)
    else:
        res['secondary_mismatch'] = None
    if ref64:
        # SURVEY 8(d) "parity noise floor": the same oracle in float64 as the yardstick -- |ours - ref64| next to |ref32 - ref64|
        om64 = om.to(torch.float64)
        om64.sdf_fn = _chunked(om64.sdf_fn, chunk=1 << 19)
        with torch.no_grad():
            r64 = pipeline.forward_with_uv(om64, inp['uv'].double(), inp['pose'].double(), inp['intrinsics'].double(), inp['object_mask'],
                                           lambda n: U[:n].double(), training, vecs[0].double(), vecs[1].double())
        h3 = hit & r64['network_object_mask']
        res['f64'] = dict(hits=int(h3.sum()), mask_mismatch_ours=int((a != r64['network_object_mask']).sum()),
                          mask_mismatch_ref32=int((b != r64['network_object_mask']).sum()), keys={})
        for k in ('sg_rgb_values', 'idr_rgb_values', 'normal_values', 'points'):
            y = r64[k][h3]
            e_ours = ((mine[k][h3].double() - y).abs() / (y.abs() + 1e-6)).flatten()
            e_ref = ((ref[k][h3].double() - y).abs() / (y.abs() + 1e-6)).flatten()
            res['f64']['keys'][k] = dict(ours=((e_ours <= 1e-4).double().mean().item(),) + _quant(e_ours),
                                         ref32=((e_ref <= 1e-4).double().mean().item(),) + _quant(e_ref))
        if verbose:
            print("   vs the float64 oracle (hits common to all three: %d; mask mismatches ours %d, fp32 oracle %d):" % (
                res['f64']['hits'], res['f64']['mask_mismatch_ours'], res['f64']['mask_mismatch_ref32']))
            for k, v in res['f64']['keys'].items():
                print("      %-16s lanes<=1e-4 / median / p95 / p99 / max:  ours %.4f %.2e %.2e %.2e %.2e | fp32 oracle %.4f %.2e %.2e %.2e %.2e" % (
                    (k,) + v['ours'] + v['ref32']))
    if grads:
        def rel(x, y):
            return (x - y).norm().item() / (y.norm().item() + 1e-20)
        res['g_lgt'] = rel(net.envmap_material_network.lgtSGs.grad, om.lgtSGs.grad)
        g_mat = [l.weight.grad for l in net.envmap_material_network.diffuse_albedo_layers if hasattr(l, "weight")]
        res['g_mat'] = [rel(x, w.grad) for x, w in zip(g_mat, om.material.W)]
        g_rad = []
        for i, w in enumerate(om.radiance.W):
            gv_mine = getattr(net.rendering_network, "lin%d" % i).weight_v.grad
            gW, v = w.grad, w.detach()
            nrm = v.norm(dim=1, keepdim=True)
            gv = gW - (gW * v).sum(1, keepdim=True) * v / (nrm * nrm)
            g_rad.append(rel(gv_mine, gv))
        res['g_rad'] = g_rad
    if verbose:
        print("FULL bumps=%.2f px=%d rays/px=%d train=%d tiers=%s: ours %.3fs oracle %.3fs | mask mismatches %d of %d (hits %d)" % (
            bumps, n_px, max(n_rays, 1), training, tiers, res['ours_s'], res['oracle_s'], res['mask_mismatch'], res['pixels'], res['hits']))
        print("   depth |dp| on hits   median %.2e p95 %.2e p99 %.2e max %.2e | within 1e-4: %.5f" % (res['depth'] + (res['depth_frac_1e4'],)))
        print("   |d sdf_output| hits  median %.2e p95 %.2e p99 %.2e max %.2e" % res['sdf_output_hit'])
        if res['sdf_output_miss']:
            print("   |d sdf_output| miss  median %.2e p95 %.2e p99 %.2e max %.2e" % res['sdf_output_miss'])
        for k, v in res['keys'].items():
            print("   %-26s rel median %.2e p95 %.2e p99 %.2e max %.2e | lanes<=1e-4: %.4f | abs p99 %.2e max %.2e" % (
                (k,) + v['q'] + (v['frac_1e4'], v['absq'][2], v['absq'][3])))
        if 'background_rel' in res:
            print("   background sg_rgb rel  median %.2e p95 %.2e p99 %.2e max %.2e" % res['background_rel'])
        if res['secondary_mismatch'] is not None:
            print("   secondary mask mismatches %d of %d; depth on common hits median %.2e p95 %.2e p99 %.2e max %.2e" % (
                (res['secondary_mismatch'], res['secondary_rays']) + res['secondary_depth']))
        if grads:
            print("   grad rel (Frobenius): lgtSGs %.2e | material %s | radiance %s" % (
                res['g_lgt'], " ".join("%.1e" % x for x in res['g_mat']), " ".join("%.1e" % x for x in res['g_rad'])))
    if tiers is not None:
        _lib.check(_lib.raw().nefii_trace_set_tiers(0, 0))
    res['_outputs'] = (mine, ref)      # for diagnostics (tools/diag_tail.py)
    return res
