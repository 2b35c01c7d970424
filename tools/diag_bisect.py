"""Trace statistics and time per bisection depth (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from nefii_b200 import _lib
    from nefii_b200.utils import rend_util
    lib = _lib.raw()
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    model.train()
    pose, K = [t.to(dev) for t in bench.make_camera()]
    for px in (256, 2048):
        uv, obj, rgb = bench.make_batch(1000, num_pixels=px)
        uv = uv.to(dev).reshape(1, -1, 2)
        objm = obj.to(dev).reshape(1, -1, 1).expand(1, px, bench.NUM_RAYS).reshape(-1)
        dirs, cam = rend_util.get_camera_params(uv, pose, K)
        rt = model.ray_tracer
        rt.collect_stats = True
        ref = None
        for depth in (1, 2, 3, 4):
            _lib.check(lib.nefii_trace_set_bisect_depth(depth))
            with torch.no_grad():
                out = rt(sdf=model.implicit_network, cam_loc=cam, object_mask=objm, ray_directions=dirs)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    out = rt(sdf=model.implicit_network, cam_loc=cam, object_mask=objm, ray_directions=dirs)
                b.record()
                torch.cuda.synchronize()
            same = True if ref is None else all(torch.equal(x, y) for x, y in zip(out, ref))
            ref = ref or out
            print("px %d depth %d: %.3f ms per trace | stats %s | identical %s" % (px, depth, a.elapsed_time(b) / 5, rt.last_stats, same))


def forward_times():
    from nefii_b200 import _lib
    lib = _lib.raw()
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    model.train()
    pose, K = [t.to(dev) for t in bench.make_camera()]
    model.ray_tracer.collect_stats = True
    for px in (256, 2048):
        uv, obj, rgb = [t.to(dev) for t in bench.make_batch(1000, num_pixels=px)]
        g = torch.Generator().manual_seed(1)
        U = torch.rand(px * bench.NUM_RAYS, 7, generator=g).to(dev)
        tu = torch.rand(100, generator=g)
        ref = None
        for depth in (1, 2, 3, 4):
            _lib.check(lib.nefii_trace_set_bisect_depth(depth))
            inp = {'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K}
            with torch.no_grad():
                out = model.forward_with_uv(inp, uniforms=U, trace_uniforms=tu)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    out = model.forward_with_uv(inp, uniforms=U, trace_uniforms=tu)
                b.record()
                torch.cuda.synchronize()
            keys = ('points', 'sg_rgb_values', 'network_object_mask', 'secondary_points', 'secondary_mask')
            same = True if ref is None else all(torch.equal(out[k], ref[k]) for k in keys)
            ref = ref or out
            print("fwd px %d depth %d: %.3f ms | last (secondary) trace stats %s | identical to depth 1: %s" % (
                px, depth, a.elapsed_time(b) / 5, model.ray_tracer.last_stats, same))


if __name__ == "__main__":
    forward_times()
    main()
