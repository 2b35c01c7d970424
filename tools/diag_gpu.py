"""GPU diagnostics printed during development trips.  TEST INFRASTRUCTURE, like tests/: it compares the CUDA path with the
oracle (and therefore imports oracle/), times kernels with CUDA events and drives the ablation masks; nothing under
nefii_b200/ imports it."""
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import inputs, sg  # noqa: E402
from tests.util import load_golden, rel_stats  # noqa: E402


def ev_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def diag_sg():
    from nefii_b200.model.sg_render import render_with_sg
    dev = torch.device("cuda:0")
    g = load_golden("sg_render_cfg1.npz")
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    for lights in ("sunrise", "synthetic", "envmap1"):
        for rough in inputs.ROUGHNESS_SWEEP:
            args = (t("lgt_" + lights), t("spec"), torch.tensor([[rough]], device=dev), t("albedo"), t("normal"), t("view"))
            got = render_with_sg(*args)
            ref = sg.render_with_sg(*args)
            for k in ("sg_specular_rgb", "sg_diffuse_rgb"):
                frac, p99, mx = rel_stats(got[k], ref[k])
                r64 = torch.from_numpy(g["%s_r%g_%s_f64" % (lights, rough, k)])
                f_ref = rel_stats(ref[k], r64)
                f_got = rel_stats(got[k], r64)
                print("SG %-9s r=%-5g %-16s vs-oracle: frac<=1e-4 %.4f p99 %.2e max %.2e | vs f64: oracle p99 %.2e ours p99 %.2e  bitexact %.3f"
                      % (lights, rough, k, frac, p99, mx, f_ref[1], f_got[1], (got[k] == ref[k]).float().mean().item()))
    for n in (1024, 131072, 1 << 20):
        normal, view, albedo = [x.to(dev) for x in inputs.shading_inputs(n, seed=0)]
        lgt = t("lgt_sunrise")
        spec, rough = t("spec"), torch.tensor([[0.3]], device=dev)
        ms = ev_time(lambda: render_with_sg(lgt, spec, rough, albedo, normal, view))
        ms_ref = ev_time(lambda: sg.render_with_sg(lgt, spec, rough, albedo, normal, view), iters=3, warm=1) if n <= 131072 else float("nan")
        print("SG time n=%d ours %.4f ms (%.1f Mray/s)  torch-oracle-on-gpu %.3f ms" % (n, ms, n / ms / 1e3, ms_ref))
        lg, rg, ag = lgt.clone().requires_grad_(True), rough.clone().requires_grad_(True), albedo.clone().requires_grad_(True)

        def fb():
            out = render_with_sg(lg, spec, rg, ag, normal, view)
            out['sg_rgb'].sum().backward()
        ms_fb = ev_time(fb, iters=5, warm=2)
        print("SG fwd+bwd n=%d ours %.4f ms (bwd ~ %.4f ms, %.1f Mray/s)" % (n, ms_fb, ms_fb - ms, n / max(ms_fb - ms, 1e-9) / 1e3))


def diag_gemm():
    from nefii_b200 import ops
    dev = torch.device("cuda:0")
    for rows in (4096, 131072, 1 << 20):
        k = n = 512
        x = torch.randn(rows, k, device=dev) * 0.3
        w = torch.randn(n, k, device=dev) / k ** 0.5
        bias = torch.zeros(n, device=dev)
        a = ops.split_to_planes(x)
        b = ops.split_to_planes(w)
        dst = (torch.empty(rows, n, device=dev, dtype=torch.bfloat16), torch.empty(rows, n, device=dev, dtype=torch.bfloat16))
        fn = lambda: ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n)
        ms = ev_time(fn)
        fl = 2.0 * rows * n * k
        print("GEMM rows=%d: %.4f ms  %.1f TFLOP/s algorithmic (x3 MMA = %.1f bf16 TFLOP/s)" % (rows, ms, fl / ms / 1e9, 3 * fl / ms / 1e9))
        if rows <= 131072:
            z = torch.nn.functional.softplus(x @ w.t() + bias, beta=100)
            got = dst[0].float() + dst[1].float()
            z64 = torch.nn.functional.softplus(x.double() @ w.double().t() + bias.double(), beta=100)
            print("   err vs f64: ours max %.2e mean %.2e | torch fp32 max %.2e mean %.2e" % (
                (got.double() - z64).abs().max().item(), (got.double() - z64).abs().mean().item(),
                (z.double() - z64).abs().max().item(), (z.double() - z64).abs().mean().item()))
            ms_t = ev_time(lambda: torch.nn.functional.softplus(x @ w.t() + bias, beta=100))
            print("   torch fp32 linear+softplus: %.4f ms" % ms_t)


def diag_sdf():
    from nefii_b200 import ops
    from oracle import mlp
    dev = torch.device("cuda:0")
    params = mlp.sdf_init(seed=1, bumps=0.3)
    fmt = int(os.environ.get("DIAG_FMT", "-1"))
    net = ops.SdfMlp(device=dev, fmt=None if fmt < 0 else fmt)
    print("SDF plane format %d" % net.format)
    net.set_weights([w.to(dev) for w in params.W], [b.to(dev) for b in params.b])
    p32 = params.to(dev)
    p64 = params.to(dev, torch.float64)
    for n in (4096, 131072, 1 << 20):
        x = (torch.rand(n, 3, device=dev) * 1.8 - 0.9)
        ms_f = ev_time(lambda: net.eval(x), iters=10)
        ms_g = ev_time(lambda: net.eval(x, want_feat=True, want_grad=True), iters=10)
        print("SDF n=%d fwd %.3f ms (%.1f TFLOP/s alg) fwd+grad %.3f ms (%.1f TFLOP/s alg)" % (
            n, ms_f, n * 3.671e6 / ms_f / 1e9, ms_g, n * 7.342e6 / ms_g / 1e9))
        if n <= 131072:
            ms_t = ev_time(lambda: mlp.sdf_forward(p32, x), iters=5)
            sdf, feat, grad = net.eval(x, want_feat=True, want_grad=True)
            r64 = mlp.sdf_forward(p64, x.double())
            r32 = mlp.sdf_forward(p32, x)
            g64 = mlp.sdf_gradient(p64, x.double())
            g32 = mlp.sdf_gradient(p32, x)
            print("   torch fp32 fwd %.3f ms | sdf err vs f64: ours max %.2e mean %.2e ; torch fp32 max %.2e mean %.2e" % (
                ms_t, (sdf.double() - r64[:, 0]).abs().max().item(), (sdf.double() - r64[:, 0]).abs().mean().item(),
                (r32[:, 0].double() - r64[:, 0]).abs().max().item(), (r32[:, 0].double() - r64[:, 0]).abs().mean().item()))
            print("   grad rel err vs f64: ours max %.2e ; torch fp32 max %.2e" % (
                ((grad.double() - g64).norm(dim=-1) / g64.norm(dim=-1)).max().item(),
                ((g32.double() - g64).norm(dim=-1) / g64.norm(dim=-1)).max().item()))


def diag_cluster():
    from nefii_b200 import ops, _lib
    dev = torch.device("cuda:0")
    rows, k, n = 262144, 512, 512
    x = torch.randn(rows, k, device=dev) * 0.3
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.zeros(n, device=dev)
    a = ops.split_to_planes(x)
    b = ops.split_to_planes(w)
    dst = (torch.empty(rows, n, device=dev, dtype=torch.bfloat16), torch.empty(rows, n, device=dev, dtype=torch.bfloat16))
    ref = None
    for cl in (1, 2):
        _lib.check(_lib.raw().nefii_gemm_set_cluster(cl))
        fn = lambda: ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n)
        ms = ev_time(fn)
        got = (dst[0].float() + dst[1].float()).clone()
        if ref is None:
            ref = got
        print("CLUSTER %d: %.4f ms %.1f TFLOP/s alg ; identical to cluster 1: %s" % (cl, ms, 2.0 * rows * n * k / ms / 1e9, torch.equal(got, ref)))
    _lib.check(_lib.raw().nefii_gemm_set_cluster(2))


def diag_kflush():
    from nefii_b200 import ops, _lib
    from oracle import mlp
    dev = torch.device("cuda:0")
    rows, k, n = 262144, 512, 512
    x = torch.randn(rows, k, device=dev) * 0.3
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.zeros(n, device=dev)
    a = ops.split_to_planes(x); b = ops.split_to_planes(w)
    dst = (torch.empty(rows, n, device=dev, dtype=torch.bfloat16), torch.empty(rows, n, device=dev, dtype=torch.bfloat16))
    xp = (torch.rand(4096, 512, device=dev) + 0.5).bfloat16().float(); wp = (torch.rand(256, 512, device=dev) + 0.5).bfloat16().float()
    ap = ops.split_to_planes(xp); bp = ops.split_to_planes(wp)
    refp = xp.double() @ wp.double().t()
    params = mlp.sdf_init(seed=1, bumps=0.3)
    net = ops.SdfMlp(device=dev)
    net.set_weights([t.to(dev) for t in params.W], [t.to(dev) for t in params.b])
    pts = torch.rand(131072, 3, device=dev) * 1.8 - 0.9
    r64 = mlp.sdf_forward(params.to(dev, torch.float64), pts.double())[:, 0]
    for kf, head in ((1, 1), (2, 2), (2, 3), (1, 3), (2, 4), (4, 4), (8, 8)):
        _lib.check(_lib.raw().nefii_gemm_set_k_flush(kf))
        _lib.check(_lib.raw().nefii_gemm_set_k_flush_head(head))
        ms = ev_time(lambda: ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n), iters=10)
        out = torch.zeros(4096, 256, device=dev)
        ops.gemm_split_bf16(ap, bp, 512, 256, dst_f32=out, f32_begin=0, f32_end=256)
        err = (out.double() - refp) / refp
        sdf, _, _ = net.eval(pts)
        e = sdf.double() - r64
        print("KFLUSH %d head %d: %.4f ms %.1f TFLOP/s alg | all-positive K=512 bias %.2f ulp | SDF err mean %.2e (signed %.2e) max %.2e" % (
            kf, head, ms, 2.0 * rows * n * k / ms / 1e9, err.mean().item() / 2 ** -24, e.abs().mean().item(), e.mean().item(), e.abs().max().item()))
    _lib.check(_lib.raw().nefii_gemm_set_k_flush(4))


def diag_ablate():
    from nefii_b200 import ops, _lib
    dev = torch.device("cuda:0")
    rows, k, n = 262144, 512, 512
    x = torch.randn(rows, k, device=dev) * 0.3
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.zeros(n, device=dev)
    a = ops.split_to_planes(x)
    b = ops.split_to_planes(w)
    dst = (torch.empty(rows, n, device=dev, dtype=torch.bfloat16), torch.empty(rows, n, device=dev, dtype=torch.bfloat16))
    names = {0: "full", 1: "no final math/stores", 8: "no flush", 9: "no flush, no final", 2: "no TMA", 4: "no MMA", 6: "no TMA, no MMA",
             11: "no TMA, no flush, no final (MMA only)", 13: "TMA only (no MMA, flush, final)", 15: "barriers only", 3: "no TMA no final", 5: "no MMA no final",
             16: "final math only (no staging, no stores)", 32: "no global stores", 22: "epilogue alone, math only", 38: "epilogue alone, no global stores"}
    for cl in (1, 2):
        _lib.check(_lib.raw().nefii_gemm_set_cluster(cl))
        for mask in (0, 1, 8, 9, 2, 3, 4, 5, 6, 11, 13, 15, 16, 32, 22, 38):
            _lib.check(_lib.raw().nefii_gemm_set_debug(mask))
            ms = ev_time(lambda: ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n), iters=10)
            print("ABLATE cl=%d mask=%2d %-42s %.4f ms" % (cl, mask, names[mask], ms))
    _lib.check(_lib.raw().nefii_gemm_set_debug(0))
    _lib.check(_lib.raw().nefii_gemm_set_cluster(2))


def diag_gemmprof():
    from nefii_b200 import ops
    dev = torch.device("cuda:0")
    rows, k, n = 131072, 512, 512
    fmt = int(os.environ.get("DIAG_FMT", "1"))       # the SDF network's default plane format (fp16 split)
    x = torch.randn(rows, k, device=dev) * 0.3
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.zeros(n, device=dev)
    a = ops.split_to_planes(x, fmt=fmt)
    b = ops.split_to_planes(w, fmt=fmt)
    pdt = torch.float16 if fmt else torch.bfloat16
    dst = (torch.empty(rows, n, device=dev, dtype=pdt), torch.empty(rows, n, device=dev, dtype=pdt))
    from nefii_b200 import _lib
    _lib.check(_lib.raw().nefii_gemm_set_debug(int(os.environ.get("GEMM_DEBUG", "0"))))     # ablation mask for A/B captures
    for _ in range(8):
        ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n)
    torch.cuda.synchronize()
    _lib.check(_lib.raw().nefii_gemm_set_debug(0))


def diag_trunc():
    """Is the fp32 accumulation inside tcgen05 biased (truncation)?  All-positive operands make a bias visible."""
    from nefii_b200 import ops
    dev = torch.device("cuda:0")
    rows, n = 4096, 256
    for k in (64, 512, 2048):
        x = torch.rand(rows, k, device=dev) + 0.5
        w = torch.rand(n, k, device=dev) + 0.5
        # make operands exactly representable in bf16 so only the accumulation can err
        x = x.bfloat16().float(); w = w.bfloat16().float()
        a = ops.split_to_planes(x); b = ops.split_to_planes(w)
        out = torch.zeros(rows, n, device=dev)
        ops.gemm_split_bf16(a, b, k, n, dst_f32=out, f32_begin=0, f32_end=n)
        ref = x.double() @ w.double().t()
        err = (out.double() - ref) / ref
        t32 = (x @ w.t()).double()
        e32 = (t32 - ref) / ref
        print("TRUNC k=%d: ours signed mean rel err %.3e (in ulps of 2^-24: %.2f) max |rel| %.2e | torch fp32 mean %.3e max %.2e" % (
            k, err.mean().item(), err.mean().item() / 2 ** -24, err.abs().max().item(), e32.mean().item(), e32.abs().max().item()))


def diag_trace():
    from nefii_b200 import ops
    from nefii_b200.model.ray_tracing import RayTracing
    from oracle import mlp, tracer as otr
    dev = torch.device("cuda:0")
    for bumps in (0.03,):
        params = mlp.sdf_init(seed=1, bumps=bumps)
        net = ops.SdfMlp(device=dev)
        net.set_weights([w.to(dev) for w in params.W], [b.to(dev) for b in params.b])

        class Src:
            def nefii_sdf_source(self):
                return 0, net.handle.value, 0, net
        p32 = params.to(dev)
        oracle_sdf = lambda x: mlp.sdf_forward(p32, x)[:, 0]
        cfg = otr.TraceConfig()
        for n_side, training in ((64, False), (64, True), (362, False), (362, True)):
            K = torch.eye(4); K[0, 0] = K[1, 1] = n_side * 2.4; K[0, 2] = K[1, 2] = n_side / 2
            pose = torch.eye(4); pose[:3, 3] = torch.tensor([0., 0., -3.])
            ii, jj = torch.meshgrid(torch.arange(n_side).float(), torch.arange(n_side).float(), indexing="xy")
            uv = torch.stack([ii, jj], -1).reshape(1, -1, 2) + 0.5
            dirs, loc = otr.camera_rays(uv, pose[None], K[None])
            dirs, loc = dirs.to(dev), loc.to(dev)
            obj = torch.ones(n_side * n_side, dtype=torch.bool, device=dev)
            rt = RayTracing(**cfg.as_kwargs()); rt.train(training); rt.collect_stats = True
            u = torch.rand(100)
            pts, mask, dist = rt(Src(), loc, obj, dirs, uniforms=u)
            rt.collect_stats = False
            ms = ev_time(lambda: rt(Src(), loc, obj, dirs, uniforms=u), iters=3, warm=1)
            line = "TRACE bumps=%.2f n=%d train=%d: %.2f ms, hits %.3f, sampler %d, rootfind %d, minsdf %d, evals/ray %.1f -> %.1f TFLOP/s alg" % (
                bumps, n_side * n_side, training, ms, mask.float().mean().item(), rt.last_stats["n_sampler"], rt.last_stats["n_rootfind"],
                rt.last_stats["n_min_sdf"], rt.last_stats["n_evals"] / (n_side * n_side), rt.last_stats["n_evals"] * 3.671e6 / ms / 1e9)
            if n_side <= 64:
                t0 = time.time()
                o_pts, o_mask, o_dist, st = otr.ray_trace(oracle_sdf, loc, obj, dirs, cfg, training=training, uniforms=u)
                torch.cuda.synchronize()
                both = mask & o_mask
                line += " | oracle(torch gpu) %.0f ms evals/ray %.1f, mask agree %.4f, depth err med %.1e max %.1e" % (
                    (time.time() - t0) * 1e3, st["n_evals"] / (n_side * n_side), (mask == o_mask).float().mean().item(),
                    (dist - o_dist)[both].abs().median().item() if both.any() else -1, (dist - o_dist)[both].abs().max().item() if both.any() else -1)
            print(line)


def diag_mis():
    from nefii_b200 import integrator
    from oracle import mis
    dev = torch.device("cuda:0")
    n = 20000
    normal, view, albedo = [x.to(dev) for x in inputs.shading_inputs(n, seed=0)]
    g = torch.Generator().manual_seed(1)
    rough = (torch.rand(n, 1, generator=g) * 0.9 + 0.089).to(dev)
    lgt = inputs.synthetic_light_sgs(128, seed=2).to(dev)
    u = torch.rand(n, 7, generator=g).to(dev)
    wi, pdf, weight, mat = integrator.mis_sample(lgt, rough, normal, view, u, want_matrix=True)
    o_wi, o_pdf, o_mat = mis.sample_directions(lgt, rough, normal, view, u)
    for s in range(3):
        dw = (wi[s] - o_wi[s]).abs().amax(-1)
        rp = ((pdf[s] - o_pdf[s, :, 0]).abs() / o_pdf[s, :, 0].abs().clamp_min(1e-6))
        print("MIS sample %d: wi maxdiff %.2e frac<2e-4 %.5f bitexact %.3f | pdf rel p99 %.2e max %.2e bitexact %.3f" % (
            s, dw.max().item(), (dw < 2e-4).float().mean().item(), (wi[s] == o_wi[s]).all(-1).float().mean().item(),
            rp.kthvalue(int(0.99 * n))[0].item(), rp.max().item(), (pdf[s] == o_pdf[s, :, 0]).float().mean().item()))
        for j in range(3):
            rm = ((mat[s, j] - o_mat[s, j, :, 0]).abs() / o_mat[s, j, :, 0].abs().clamp_min(1e-6))
            print("     mat[%d][%d] rel p99 %.2e max %.2e" % (s, j, rm.kthvalue(int(0.99 * n))[0].item(), rm.max().item()))


def diag_dense():
    from nefii_b200 import mlp, ops
    from oracle import mlp as omlp
    dev = torch.device("cuda:0")
    for n in (500, 6000):
        g = torch.Generator().manual_seed(n)
        pts = (torch.rand(n, 3, generator=g) - 0.5).to(dev)
        nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
        view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
        feat = torch.rand(n, 512, generator=g).to(dev)
        p = omlp.radiance_init(seed=2).to(dev)
        p.requires_grad_(True)
        gy = torch.rand(n, 3, generator=g).to(dev)
        ref = omlp.radiance_forward(p, pts, nrm, view, feat)
        (ref * gy).sum().backward()
        ref_grads = [t.grad.clone() for t in p.tensors()]
        for t in p.tensors():
            t.grad = None
        raw = mlp.dense_mlp([(pts, 10), (view, 4), (nrm, -1), (feat, -1)], p.W, p.b, ops.ACT_RELU)
        out = raw ** 2
        print("DENSE n=%d fwd max abs err %.2e (ref max %.2e)" % (n, (out - ref).abs().max().item(), ref.abs().max().item()))
        (out * gy).sum().backward()
        for i, (t, want) in enumerate(zip(p.tensors(), ref_grads)):
            print("   grad %d shape %s: max err %.2e scale %.2e" % (i, tuple(want.shape), (t.grad - want).abs().max().item(), want.abs().max().item()))


def diag_dotorder():
    """Which association does torch use for 3-element sum / norm / cross on this device?"""
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    a = torch.randn(200000, 3, device=dev); b = torch.randn(200000, 3, device=dev)
    p = a * b
    s = torch.sum(p, dim=-1)
    cands = {"(0+1)+2": (p[:, 0] + p[:, 1]) + p[:, 2], "(0+2)+1": (p[:, 0] + p[:, 2]) + p[:, 1], "0+(1+2)": p[:, 0] + (p[:, 1] + p[:, 2])}
    for k, v in cands.items():
        print("SUM3 %s: bitexact %.4f" % (k, (v == s).float().mean().item()))
    s4 = torch.sum(p.reshape(-1, 1, 1, 3).expand(-1, 4, 1, 3), dim=-1)[:, 0, 0]
    for k, v in cands.items():
        print("SUM3(expanded) %s: bitexact %.4f" % (k, (v == s4).float().mean().item()))
    sk = torch.sum(p, dim=-1, keepdim=True)[:, 0]
    print("keepdim same as not:", (sk == s).all().item())
    nrm = torch.norm(a, dim=-1)
    q = a * a
    ncands = {"sqrt((0+1)+2)": torch.sqrt((q[:, 0] + q[:, 1]) + q[:, 2]), "sqrt((0+2)+1)": torch.sqrt((q[:, 0] + q[:, 2]) + q[:, 1]),
              "sqrt(0+(1+2))": torch.sqrt(q[:, 0] + (q[:, 1] + q[:, 2])),
              "fma chain": torch.sqrt(torch.addcmul(torch.addcmul(q[:, 0], a[:, 1], a[:, 1]), a[:, 2], a[:, 2]))}
    for k, v in ncands.items():
        print("NORM3 %s: bitexact %.4f" % (k, (v == nrm).float().mean().item()))
    nk = torch.norm(a, dim=-1, keepdim=True)[:, 0]
    print("norm keepdim same:", (nk == nrm).all().item())
    nrm2 = a.norm(2, 1)
    print("a.norm(2,1) same:", (nrm2 == nrm).all().item())
    c = torch.cross(a, b, dim=-1)
    c_plain = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    c_fma = torch.addcmul(-(a[:, 2] * b[:, 1]), a[:, 1], b[:, 2])
    c_fma2 = -torch.addcmul(-(a[:, 1] * b[:, 2]), a[:, 2], b[:, 1])
    print("CROSS plain %.4f fma(a1*b2 fused) %.4f fma(a2*b1 fused) %.4f" % ((c[:, 0] == c_plain).float().mean().item(),
          (c[:, 0] == c_fma).float().mean().item(), (c[:, 0] == c_fma2).float().mean().item()))
    # sum over 128 (dim=-2) order is not reproducible by a sequential loop; report how far a sequential sum is
    x = torch.rand(5000, 128, 3, device=dev)
    seq = torch.zeros(5000, 3, device=dev)
    for m in range(128):
        seq = seq + x[:, m]
    ts = x.sum(-2)
    print("SUM128 sequential vs torch: bitexact %.4f max rel %.2e" % ((seq == ts).float().mean().item(), ((seq - ts).abs() / ts).max().item()))
    # F.normalize and pow
    r = torch.rand(100000, device=dev) + 0.05
    print("pow4: powf %.4f (r*r)*(r*r) %.4f ((r*r)*r)*r %.4f" % ((torch.pow(r, 4) == torch.pow(r, torch.tensor(4.0, device=dev))).float().mean().item(),
          (torch.pow(r, 4) == (r * r) * (r * r)).float().mean().item(), (torch.pow(r, 4) == ((r * r) * r) * r).float().mean().item()))
    print("div by python scalar == mul by reciprocal: %.4f ; == true division %.4f" % (
        ((r / 3.141592653589793) == r * torch.tensor(1.0 / 3.14159274, device=dev).float()).float().mean().item(),
        ((r / 3.141592653589793) == r / torch.tensor(3.141592653589793, device=dev)).float().mean().item()))


def diag_pipeline():
    from oracle import pipeline, ref_harness as rh
    from nefii_b200.model.implicit_differentiable_renderer import IDRNetwork
    from nefii_b200.utils.conf import default_model_conf
    dev = torch.device("cuda:0")
    om = rh.small_model(seed=0, bumps=float(os.environ.get("PIPE_BUMPS", "0.03")))     # larger: rougher SDF, more sampler / bisection rays
    torch.manual_seed(0)
    net = IDRNetwork(default_model_conf()).to(dev)
    rh.load_oracle_weights(net, om)
    om = om.to(dev)
    scale = int(os.environ.get("PIPE_SCALE", "1"))      # 2: 16 384 eval rays / 36 864 training rays
    for training, rays, n_side in ((False, 0, 64 * scale), (True, 4, 48 * scale)):
        net.train(training)
        uv, pose, K = rh.camera_batch(n_side, rays, seed=3)
        S = uv.shape[1]
        obj = torch.ones(1, S, dtype=torch.bool)
        g = torch.Generator().manual_seed(103)
        U = torch.rand(S * max(rays, 1), 7, generator=g).to(dev)
        vecs = [torch.rand(100, generator=g) for _ in range(2)]
        inp = dict(uv=uv.to(dev), pose=pose.to(dev), intrinsics=K.to(dev), object_mask=obj.to(dev))
        with torch.no_grad():
            mine = net.forward_with_uv(inp, uniforms=U, trace_uniforms=vecs[0])
            ref = pipeline.forward_with_uv(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda n: U[:n], training, vecs[0], vecs[1])
        a, b = mine['network_object_mask'], ref['network_object_mask']
        agree = a == b
        hit = agree & a
        print("PIPE train=%d rays/px=%d pixels=%d: mask agree %.5f (%d mismatches), hits %d" % (training, max(rays, 1), S, agree.float().mean().item(), int((~agree).sum()), int(hit.sum())))
        print("   depth |dp| on hits: median %.2e p99 %.2e max %.2e" % tuple(((mine['points'] - ref['points'])[hit].abs().amax(-1)).quantile(torch.tensor([0.5, 0.99, 1.0], device=dev)).tolist()))
        for k in ('normal_values', 'idr_rgb_values', 'sg_rgb_values', 'sg_diffuse_rgb_values', 'sg_specular_rgb_values', 'sg_roughness_values', 'sg_diffuse_albedo_values'):
            x, y = mine[k][hit].float(), ref[k][hit].float()
            rel = ((x - y).abs() / (y.abs() + 1e-3)).flatten()
            print("   %-26s rel err median %.2e p95 %.2e p99 %.2e max %.2e" % ((k,) + tuple(rel.quantile(torch.tensor([0.5, 0.95, 0.99, 1.0], device=dev)).tolist())))
        if mine['secondary_mask'] is not None and mine['secondary_mask'].shape == ref['secondary_mask'].shape:
            print("   secondary mask agree %.5f" % (mine['secondary_mask'] == ref['secondary_mask']).float().mean().item())


from tests.parity_util import fullsize_compare  # noqa: E402


def diag_fullsize():
    """PARITY at BASELINE configs[2] size.  Env: FULL_BUMPS (0.08), FULL_PX (2048), FULL_RAYS (64), FULL_TIERS ("1,0;4,4;...")."""
    dev = torch.device("cuda:0")
    bumps = float(os.environ.get("FULL_BUMPS", "0.08"))
    n_px = int(os.environ.get("FULL_PX", "2048"))
    n_rays = int(os.environ.get("FULL_RAYS", "64"))
    tiers = [tuple(int(v) for v in t.split(",")) for t in os.environ.get("FULL_TIERS", "0,0").split(";")]
    for t in tiers:
        fullsize_compare(dev, bumps, n_px, n_rays, True, tiers=t, grads=True, ref64=os.environ.get("FULL_REF64", "0") == "1")
    fullsize_compare(dev, bumps, n_px * 4, 0, False, tiers=tiers[0], grads=False)
    # per-ray lanes (no averaging over the rays of a pixel): the same rays as single-ray pixels
    fullsize_compare(dev, bumps, min(n_px * max(n_rays, 1), 32768), 0, True, tiers=tiers[0], grads=False)


def diag_trunccomp():
    """Calibrates and checks the first-order compensation of tcgen05's round-toward-zero accumulation
    (nefii_gemm_set_trunc_comp): slope of the accumulation error against the exact product sum, per partial length."""
    from nefii_b200 import ops, _lib
    from oracle import mlp
    dev = torch.device("cuda:0")
    lib = _lib.raw()
    fmt = int(os.environ.get("DIAG_FMT", "0"))      # 0 bf16 split, 1 fp16 split
    print("TRUNCCOMP plane format %d" % fmt)
    rows, k, n = 16384, 512, 512
    g = torch.Generator(device="cpu").manual_seed(0)
    slopes = {}
    for data in ("positive", "gauss"):
        x = torch.randn(rows, k, generator=g) * 0.3
        if data == "positive":
            x = x.abs()
        w = torch.randn(n, k, generator=g) / k ** 0.5
        x, w = x.to(dev), w.to(dev)
        a = ops.split_to_planes(x, fmt=fmt); b = ops.split_to_planes(w, fmt=fmt)
        ah, al, bh, bl = [t.double() for t in (a[0], a[1], b[0], b[1])]
        ref3 = ah @ bh.t() + ah @ bl.t() + al @ bh.t()        # what the three MMAs sum, exactly
        for L in (1, 2, 4, 8):
            _lib.check(lib.nefii_gemm_set_trunc_comp_fmt(fmt, L, 0.0))
            out = torch.zeros(rows, n, device=dev)
            ops.gemm_split_bf16(a, b, k, n, dst_f32=out, f32_begin=0, f32_end=n, k_flush=L)
            err = out.double() - ref3
            slope = ((err * ref3).sum() / (ref3 * ref3).sum()).item()
            resid = (err - slope * ref3)
            slopes[(data, L)] = slope
            print("TRUNCCOMP %-8s L=%d: slope %.3e (%.2f ulp of 2^-24) | err rms %.3e -> after removing the slope %.3e | mean|ref| %.3f" % (
                data, L, slope, slope / 2 ** -24, err.pow(2).mean().sqrt().item(), resid.pow(2).mean().sqrt().item(), ref3.abs().mean().item()))
    params = mlp.sdf_init(seed=1, bumps=0.3)
    net = ops.SdfMlp(device=dev, fmt=fmt)
    net.set_weights([t.to(dev) for t in params.W], [t.to(dev) for t in params.b])
    pts = torch.rand(131072, 3, device=dev) * 1.8 - 0.9
    p64 = params.to(dev, torch.float64)
    r64 = mlp.sdf_forward(p64, pts.double())[:, 0]
    g64 = mlp.sdf_gradient(p64, pts[:16384].double())

    def report(tag):
        for L in (1, 2, 4, 8):
            sdf, _, grad = net.eval(pts[:16384], want_grad=True, k_flush=L)
            sdf_all, _, _ = net.eval(pts, k_flush=L)
            e = sdf_all.double() - r64
            ge = ((grad.double() - g64).norm(dim=-1) / g64.norm(dim=-1))
            print("   %s SDF k_flush=%d: err mean|.| %.2e signed %.2e rms-about-mean %.2e max %.2e | grad rel median %.2e p99 %.2e" % (
                tag, L, e.abs().mean().item(), e.mean().item(), (e - e.mean()).pow(2).mean().sqrt().item(), e.abs().max().item(),
                ge.median().item(), ge.kthvalue(int(0.99 * ge.numel()))[0].item()))
    report("comp off")
    for src in ("positive", "gauss"):
        for L in (1, 2, 4, 8):
            _lib.check(lib.nefii_gemm_set_trunc_comp_fmt(fmt, L, -slopes[(src, L)]))
        report("comp from %-8s" % src)
    for L in (1, 2, 4, 8):
        _lib.check(lib.nefii_gemm_set_trunc_comp_fmt(fmt, L, 0.0))
    r32 = mlp.sdf_forward(params.to(dev), pts)[:, 0]
    e = r32.double() - r64
    print("   torch fp32 for scale: err mean|.| %.2e signed %.2e max %.2e" % (e.abs().mean().item(), e.mean().item(), e.abs().max().item()))


def diag_l2fit():
    """Is the layer GEMM faster per row tile when its activations stay L2-resident (small row counts, same buffers reused)?"""
    from nefii_b200 import ops
    dev = torch.device("cuda:0")
    k = n = 512
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.zeros(n, device=dev)
    b = ops.split_to_planes(w)
    for rows in (18944, 37888, 75776, 151552, 262144, 1 << 20):
        x = torch.randn(rows, k, device=dev) * 0.3
        a = ops.split_to_planes(x)
        dst = (torch.empty(rows, n, device=dev, dtype=torch.bfloat16), torch.empty(rows, n, device=dev, dtype=torch.bfloat16))
        ms = ev_time(lambda: ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n), iters=20)
        waves = rows / 128 / 148
        print("L2FIT rows %8d (%.1f MB in + out): %.4f ms  %.1f TFLOP/s alg  %.2f us per wave of 148 tiles" % (
            rows, rows * 4096 / 1e6, ms, 2.0 * rows * n * k / ms / 1e9, ms * 1e3 / waves))




def diag_sdfbias():
    """Where does the signed SDF error come from?  Feature vector (= last hidden activation) against f64, signed, per tier."""
    from nefii_b200 import ops
    from oracle import mlp
    dev = torch.device("cuda:0")
    fmt = int(os.environ.get("DIAG_FMT", "-1"))
    params = mlp.sdf_init(seed=1, bumps=0.3)
    net = ops.SdfMlp(device=dev, fmt=None if fmt < 0 else fmt)
    net.set_weights([t.to(dev) for t in params.W], [t.to(dev) for t in params.b])
    pts = torch.rand(16384, 3, device=dev) * 1.8 - 0.9
    p64 = params.to(dev, torch.float64)
    out64 = mlp.sdf_forward(p64, pts.double())
    r64, f64 = out64[:, 0], out64[:, 1:]
    w_last = p64.W[-1][0]
    for L in (1, 4, 8):
        sdf, feat, _ = net.eval(pts, want_feat=True, k_flush=L)
        e = sdf.double() - r64
        fe = feat.double() - f64
        big = f64 > 1e-3
        relf = (fe / f64)[big]
        # what the feature error alone explains of the SDF error (the output layer is fp32 FMAs)
        via_feat = fe @ w_last
        print("SDFBIAS fmt %d k_flush=%d: sdf err signed %.2e rms %.2e | feat rel err signed %.2e rms %.2e | sdf err explained by feat: signed %.2e, rest signed %.2e rms %.2e" % (
            net.format, L, e.mean().item(), e.pow(2).mean().sqrt().item(), relf.mean().item(), relf.pow(2).mean().sqrt().item(),
            via_feat.mean().item(), (e - via_feat).mean().item(), (e - via_feat).pow(2).mean().sqrt().item()))
    r32 = mlp.sdf_forward(params.to(dev), pts)
    fe = (r32[:, 1:].double() - f64)
    print("SDFBIAS torch fp32: sdf err signed %.2e | feat rel signed %.2e rms %.2e" % (
        (r32[:, 0].double() - r64).mean().item(), (fe / f64)[f64 > 1e-3].mean().item(), (fe / f64)[f64 > 1e-3].pow(2).mean().sqrt().item()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["sg", "gemm"]
    print(torch.cuda.get_device_name(0))
    for w in which:
        globals()["diag_" + w]()
