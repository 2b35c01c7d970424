"""How much of the prefetched primary trace really overlaps the step of the batch before it?  Diagnostic (events + torch.profiler)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def ev(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    from torch.profiler import profile, ProfilerActivity
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    pose, K = [t.to(dev) for t in bench.make_camera()]
    flat = bench.FlatGrads(model.parameters())
    opt_idr = torch.optim.Adam([p for p in model.rendering_network.parameters() if p.requires_grad], lr=5e-4)
    opt_sg = torch.optim.Adam([p for p in model.envmap_material_network.parameters() if p.requires_grad], lr=5e-4)
    n_px = int(os.environ.get("PX", "256"))
    batches = [[t.to(dev) for t in bench.make_batch(50 + i, num_pixels=n_px)] for i in range(8)]

    def inp(b):
        return {'uv': b[0], 'object_mask': b[1], 'pose': pose, 'intrinsics': K}

    def step(b):
        flat.zero()
        out = model(inp(b))
        loss = bench.idr_loss(out, b[2])
        loss.backward()
        opt_idr.step()
        opt_sg.step()

    def plain():
        for b in batches:
            step(b)

    def piped():
        model.prefetch_trace(inp(batches[0]))
        for i, b in enumerate(batches):
            model.prefetch_trace(inp(batches[(i + 1) % len(batches)]))
            step(b)
        model.prefetch_join()

    def trace_only():
        for b in batches:
            uv = b[0].reshape(1, -1, 2)
            obj = b[1].reshape(1, -1, 1).expand(1, b[0].shape[1], b[0].shape[2]).reshape(-1)
            with torch.no_grad():
                model._primary_trace(uv, pose, K, obj, None, model.ray_tracer, True)

    def rest_only():
        # every step finds its trace ready: prefetch all first (two slots only -> one at a time, joined)
        for i, b in enumerate(batches):
            model.prefetch_trace(inp(b))
            model.prefetch_join()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(b)
            e.record()
            torch.cuda.synchronize()
            rest_only.t.append(a.elapsed_time(e))
    rest_only.t = []

    for _ in range(2):
        plain()
        piped()
    n = len(batches)
    print("px %d: plain step %.2f ms | pipelined %.2f ms | primary trace alone %.2f ms" % (n_px, ev(plain, 3) / n, ev(piped, 3) / n, ev(trace_only, 3) / n))
    rest_only()
    rest_only.t = []
    rest_only()
    print("step with its trace ready (no overlap): %.2f ms" % (sum(rest_only.t) / len(rest_only.t)))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        piped()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    by_stream = {}
    for e in evs:
        sid = getattr(e, "stream", None)
        if sid is None:
            sid = getattr(e, "device_resource_id", -1)
        by_stream.setdefault(sid, []).append((e.time_range.start, e.time_range.end, e.name))

    def union(iv):
        iv = sorted((s, t) for s, t, _ in iv)
        tot, cs, ce = 0.0, None, None
        out = []
        for s, t in iv:
            if ce is None or s > ce:
                if ce is not None:
                    out.append((cs, ce))
                cs, ce = s, t
            else:
                ce = max(ce, t)
        if ce is not None:
            out.append((cs, ce))
        return out

    unions = {k: union(v) for k, v in by_stream.items()}
    for k, u in unions.items():
        print("stream %s: %d activities, busy %.2f ms per step, span %.2f ms" % (k, len(by_stream[k]), sum(t - s for s, t in u) / n / 1000.0,
                                                                                (u[-1][1] - u[0][0]) / 1000.0))
    ks = sorted(unions, key=lambda k: -len(by_stream[k]))
    if len(ks) >= 2:
        a, b = unions[ks[0]], unions[ks[1]]
        i = j = 0
        ov = 0.0
        while i < len(a) and j < len(b):
            lo, hi = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
            if hi > lo:
                ov += hi - lo
            if a[i][1] < b[j][1]:
                i += 1
            else:
                j += 1
        print("overlap of the two busiest streams: %.2f ms per step" % (ov / n / 1000.0))
        # timeline of the side stream (fewer activities): busy windows merged at 200 us, relative to the first activity
        side = unions[ks[1]]
        t0 = min(unions[ks[0]][0][0], side[0][0])
        merged = []
        for s, t in side:
            if merged and s - merged[-1][1] < 200:
                merged[-1][1] = t
            else:
                merged.append([s, t])
        print("side-stream windows (ms from start): " + " ".join("[%.1f-%.1f]" % ((s - t0) / 1000.0, (t - t0) / 1000.0) for s, t in merged[:40]))
        main_ = unions[ks[0]]
        merged = []
        for s, t in main_:
            if merged and s - merged[-1][1] < 200:
                merged[-1][1] = t
            else:
                merged.append([s, t])
        print("main-stream windows (ms from start): " + " ".join("[%.1f-%.1f]" % ((s - t0) / 1000.0, (t - t0) / 1000.0) for s, t in merged[:60]))


if __name__ == "__main__":
    main()
