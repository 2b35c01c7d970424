"""One bench step bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    pose, K = [t.to(dev) for t in bench.make_camera()]
    flat = bench.FlatGrads(model.parameters())
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=5e-4)
    n_px = int(os.environ.get("PROFILE_PIXELS", bench.NUM_PIXELS))

    def step(seed):
        uv, obj, rgb = [t.to(dev) for t in bench.make_batch(seed, num_pixels=n_px)]
        flat.zero()
        out = model({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
        loss = bench.idr_loss(out, rgb)
        loss.backward()
        opt.step()
        return loss

    step(0)
    step(1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step(2)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
