"""Timeline of ONE primary trace (fixed launch schedule so that the profiler sees every kernel): per SDF evaluation its wall time,
kernel time and the gap before it.  Diagnostic for the small-batch step.  Run with NEFII_TRACE_GRAPH=0."""
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
    n_px = int(os.environ.get("PX", "256"))
    batches = [[t.to(dev) for t in bench.make_batch(50 + i, num_pixels=n_px)] for i in range(4)]

    def trace(b):
        uv = b[0].reshape(1, -1, 2)
        obj = b[1].reshape(1, -1, 1).expand(1, b[0].shape[1], b[0].shape[2]).reshape(-1)
        with torch.no_grad():
            return model._primary_trace(uv, pose, K, obj, None, model.ray_tracer, True)

    for b in batches:
        trace(b)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for b in batches:
        trace(b)
    e.record()
    torch.cuda.synchronize()
    print("px %d: primary trace %.2f ms (graph mode %d)" % (n_px, a.elapsed_time(e) / len(batches),
                                                            int(__import__("nefii_b200")._lib.raw().nefii_trace_graph_mode())))
    model.ray_tracer.collect_stats = True
    trace(batches[0])
    print("stats:", model.ray_tracer.last_stats)
    model.ray_tracer.collect_stats = False
    whole_step = os.environ.get("PHASE", "trace") == "step"
    if whole_step:
        flat = bench.FlatGrads(model.parameters())
        opt_idr = torch.optim.Adam([p for p in model.rendering_network.parameters() if p.requires_grad], lr=5e-4)
        opt_sg = torch.optim.Adam([p for p in model.envmap_material_network.parameters() if p.requires_grad], lr=5e-4)

        def step(b):
            flat.zero()
            out = model({'uv': b[0], 'object_mask': b[1], 'pose': pose, 'intrinsics': K})
            loss = bench.idr_loss(out, b[2])
            loss.backward()
            opt_idr.step()
            opt_sg.step()
        for b in batches:
            step(b)
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if whole_step:
            step(batches[0])
            step(batches[1])
        else:
            trace(batches[0])
        torch.cuda.synchronize()
    evs = sorted(((ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events()
                  if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda x: x[0])
    t0 = evs[0][0]
    groups = []          # consecutive kernels of one kind: 'E' = encode + gemm chain, else the kernel name
    for s, t, name in evs:
        short = name.split("::")[-1].split("(")[0].split("<")[0]
        kind = "eval" if ("gemm_split" in name or "encode_kernel" in name) else short
        if whole_step and kind != "eval":
            kind = "nefii-other" if "nefii" in name else "torch"
        if groups and groups[-1]["kind"] == kind and not (kind == "eval" and "encode_kernel" in name):
            g = groups[-1]
            g["end"] = max(g["end"], t); g["busy"] += t - s; g["n"] += 1
        else:
            groups.append(dict(kind=kind, start=s, end=t, busy=t - s, n=1))
    prev_end = t0
    tot_eval = tot_gap = tot_other = 0.0
    lines = []
    for g in groups:
        gap = g["start"] - prev_end
        prev_end = max(prev_end, g["end"])
        wall = g["end"] - g["start"]
        if g["kind"] == "eval":
            tot_eval += wall
        else:
            tot_other += wall
        tot_gap += max(gap, 0)
        lines.append("%8.2f ms  gap %6.0f us  %-28s n=%3d wall %7.0f us busy %7.0f us" % ((g["start"] - t0) / 1000.0, gap, g["kind"][:28], g["n"],
                                                                                        wall, g["busy"]))
    print("total span %.2f ms: evaluations %.2f ms, other kernels %.2f ms, gaps %.2f ms, %d kernels" % (
        (prev_end - t0) / 1000.0, tot_eval / 1000.0, tot_other / 1000.0, tot_gap / 1000.0, len(evs)))
    big = [l for l, g in zip(lines, groups) if (g["end"] - g["start"]) > 60 or g["kind"] != "eval"]
    print("\n".join(lines if len(lines) < 260 else big[:260]))


if __name__ == "__main__":
    main()
