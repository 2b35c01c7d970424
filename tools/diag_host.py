"""Host (enqueue) time against device time of a bench step, and the top host-side functions (cProfile).  Diagnostic."""
import cProfile
import os
import pstats
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
    idr_params = [p for p in model.rendering_network.parameters() if p.requires_grad]
    sg_params = [p for p in model.envmap_material_network.parameters() if p.requires_grad]
    opt_idr = torch.optim.Adam(idr_params, lr=5e-4)
    opt_sg = torch.optim.Adam(sg_params, lr=5e-4)
    n_px = int(os.environ.get("PX", "256"))
    batches = [[t.to(dev) for t in bench.make_batch(50 + i, num_pixels=n_px)] for i in range(8)]

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
    t0 = time.perf_counter()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for b in batches[3:]:
        step(*b)
    t1 = time.perf_counter()
    e.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = len(batches) - 3
    print("px %d: host enqueue %.2f ms per step | device %.2f ms per step | wall incl. final sync %.2f ms per step" % (
        n_px, (t1 - t0) * 1e3 / n, a.elapsed_time(e) / n, (t2 - t0) * 1e3 / n))
    pr = cProfile.Profile()
    pr.enable()
    for b in batches[3:]:
        step(*b)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(18)
    st.sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
