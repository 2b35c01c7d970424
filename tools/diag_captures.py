"""When are trace graphs captured?  Runs bench-like steps and prints the capture count and the secondary-ray bucket per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from nefii_b200 import _lib
    lib = _lib.raw()
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    pose, K = [t.to(dev) for t in bench.make_camera()]
    flat = bench.FlatGrads(model.parameters())
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=5e-4)
    for i in range(16):
        uv, obj, rgb = [t.to(dev) for t in bench.make_batch(1000 + i)]
        c0 = int(lib.nefii_trace_graph_captures())
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        flat.zero()
        out = model({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
        bench.idr_loss(out, rgb).backward()
        opt.step()
        e.record()
        torch.cuda.synchronize()
        n_sec = 3 * out['secondary_mask'].shape[1]
        print("step %2d: %.1f ms, captures in this step %d, secondary rays %d (bucket %d), shapes held %d" % (
            i, a.elapsed_time(e), int(lib.nefii_trace_graph_captures()) - c0, n_sec, (n_sec + 8191) // 8192, len(model.ray_tracer._shape_bufs)))


if __name__ == "__main__":
    main()
