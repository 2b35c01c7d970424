"""GPU busy time against wall time of one bench step (torch.profiler, warm): how much of a small step is launch gaps?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from torch.profiler import profile, ProfilerActivity
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    pose, K = [t.to(dev) for t in bench.make_camera()]
    flat = bench.FlatGrads(model.parameters())
    idr_params = [p for p in model.rendering_network.parameters() if p.requires_grad]
    sg_params = [p for p in model.envmap_material_network.parameters() if p.requires_grad]
    opt_idr = torch.optim.Adam(idr_params, lr=5e-4)
    opt_sg = torch.optim.Adam(sg_params, lr=5e-4)
    for n_px in (256, 2048):
        batches = [[t.to(dev) for t in bench.make_batch(50 + i, num_pixels=n_px)] for i in range(6)]

        def step(uv, obj, rgb):
            flat.zero()
            out = model({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
            loss = bench.idr_loss(out, rgb)
            loss.backward()
            opt_idr.step()
            opt_sg.step()
        for b in batches[:3]:
            step(*b)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for b in batches[3:]:
            step(*b)
        e.record()
        torch.cuda.synchronize()
        wall = a.elapsed_time(e) / 3
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for b in batches[3:]:
                step(*b)
            torch.cuda.synchronize()
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        busy = sum(ev.device_time for ev in evs) / 3 / 1000.0 if hasattr(evs[0], "device_time") else sum(ev.cuda_time for ev in evs) / 3 / 1000.0
        # union of the kernel intervals (kernels of different streams / PDL overlap)
        iv = sorted((ev.time_range.start, ev.time_range.end) for ev in evs)
        union, cur_s, cur_e = 0.0, None, None
        for s, t in iv:
            if cur_e is None or s > cur_e:
                if cur_e is not None:
                    union += cur_e - cur_s
                cur_s, cur_e = s, t
            else:
                cur_e = max(cur_e, t)
        union += (cur_e - cur_s) if cur_e is not None else 0
        # where the device waits: the largest gaps between consecutive device activities, and the gap time by what follows
        named = sorted(((ev.time_range.start, ev.time_range.end, ev.name) for ev in evs), key=lambda x: x[0])
        gaps, end, last = [], None, ""
        for s, t, name in named:
            if end is not None and s > end:
                gaps.append((s - end, last, name))
            if end is None or t > end:
                end, last = t, name
        tot = sum(g[0] for g in gaps) / 3 / 1000.0
        big = sum(g[0] for g in gaps if g[0] > 20) / 3 / 1000.0
        print("px %d: gaps %.2f ms per step in %d gaps (%.2f ms in gaps > 20 us)" % (n_px, tot, len(gaps) // 3, big))
        by_next = {}
        for g, a_, b_ in gaps:
            k = b_[:60]
            by_next[k] = by_next.get(k, 0.0) + g
        for k, v in sorted(by_next.items(), key=lambda kv: -kv[1])[:14]:
            print("    wait before %-60s %.3f ms per step" % (k, v / 3 / 1000.0))
        for g, a_, b_ in sorted(gaps, key=lambda x: -x[0])[:12]:
            print("    gap %7.1f us  after %-44s before %s" % (g, a_[:44], b_[:44]))
        print("px %d: step %.2f ms (events) | %d device activities per step, summed %.2f ms, union of intervals %.2f ms per step" % (
            n_px, wall, len(evs) // 3, busy, union / 3 / 1000.0))


if __name__ == "__main__":
    main()
