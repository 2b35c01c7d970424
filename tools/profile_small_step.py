"""Where does a small (one-rank-of-eight sized) training step spend its time?  Wall clock per phase with synchronisation
around each phase, plus the torch profiler's CPU-side table.  TEST/diagnostic infrastructure."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    pose, K = [t.to(dev) for t in bench.make_camera()]
    flat = bench.FlatGrads(model.parameters())
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=5e-4)
    batches = [[t.to(dev) for t in bench.make_batch(i)] for i in range(12)]

    def step(b):
        uv, obj, rgb = b
        flat.zero()
        out = model({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
        loss = bench.idr_loss(out, rgb)
        loss.backward()
        opt.step()
        return loss

    for b in batches[:4]:
        step(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in batches[4:]:
        step(b)
    torch.cuda.synchronize()
    print("STEP wall %.2f ms per step (%d pixels x %d rays)" % ((time.perf_counter() - t0) / 8 * 1e3, bench.NUM_PIXELS, bench.NUM_RAYS))
    # phases with a sync around each
    uv, obj, rgb = batches[0]
    inp = {'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K}

    def timed(name, fn, n=5):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize()
        print("   %-34s %.2f ms" % (name, (time.perf_counter() - t) / n * 1e3))
        return r
    from nefii_b200.utils import rend_util
    B, S, R, _ = uv.shape
    uv2 = uv.reshape(B, S * R, 2)
    dirs, cam = timed("get_camera_params", lambda: rend_util.get_camera_params(uv2, pose, K))
    om = obj.reshape(B, S, 1).expand(B, S, R).reshape(-1)
    with torch.no_grad():
        pts, hit, dist = timed("primary trace (train)", lambda: model.ray_tracer(sdf=model.implicit_network, cam_loc=cam, object_mask=om, ray_directions=dirs))
        timed("sdf eval all rays", lambda: model.implicit_network.evaluate(pts))
        idx = torch.nonzero(hit).squeeze(1)
        xs = pts.index_select(0, idx)
        timed("sdf eval feat+grad on hits (%d)" % xs.shape[0], lambda: model.implicit_network.evaluate(xs, want_feat=True, want_grad=True))
    view = -dirs.reshape(-1, 3).index_select(0, idx)
    timed("get_rbg_value fwd (incl. secondary trace)", lambda: model.get_rbg_value(xs, view))
    timed("forward", lambda: model(inp))
    def fb():
        flat.zero()
        out = model(inp)
        bench.idr_loss(out, rgb).backward()
    timed("forward+loss+backward", fb)
    timed("adam", lambda: opt.step())
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for b in batches[4:8]:
            step(b)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=50))


if __name__ == "__main__":
    main()
