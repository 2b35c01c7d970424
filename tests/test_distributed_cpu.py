"""CPU, world_size 2 over gloo: the host-side multi-GPU logic of bench.py -- ONE global pixel batch whose 2x2 patches are
dealt to the ranks (reference datasets/scene_dataset.py:268-279, training/idr_train.py:653-662) and the single flat gradient
all-reduce that replaces DDP's buckets."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                      # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    net[0].bias.requires_grad_(False)         # frozen parameters stay out of the flat buffer
    flat = bench.FlatGrads(net.parameters())
    uv, obj, rgb = bench.shard_batch(bench.make_batch(1000, num_pixels=64, num_rays=2), rank, world)      # this rank's patches of the global batch
    x = torch.cat([uv[0, :, 0], uv[0, :, 1], uv[0, :, 0] * 1e-3], dim=-1)
    flat.zero()
    loss = (net(x / 800.0) - rgb).abs().mean()
    loss.backward()
    local = flat.flat.clone()
    flat.all_reduce(world)
    # grads are views into the flat buffer: the optimizer sees the averaged values without any copy
    assert net[2].weight.grad.data_ptr() >= flat.flat.data_ptr()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    out[rank] = (bool(torch.allclose(flat.flat, expect, atol=1e-7)), flat.flat.numel(), float(uv.sum()))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == 6 * 16 + 16 * 3 + 3          # trainable parameters only (frozen bias excluded)
    assert out[0][2] != out[1][2]                    # each rank rendered different pixels
    whole = bench.make_batch(1000, num_pixels=64, num_rays=2)[0]
    assert abs(out[0][2] + out[1][2] - float(whole.sum())) < 1e-2 * abs(float(whole.sum())) * 1e-3 + 1.0      # together: the global batch


def test_shard_batch_deals_whole_patches():
    batch = bench.make_batch(5, num_pixels=96, num_rays=4)
    world = 4
    parts = [bench.shard_batch(batch, r, world) for r in range(world)]
    assert all(p[0].shape == (1, 24, 4, 2) and p[1].shape == (1, 24) and p[2].shape == (24, 3) for p in parts)
    # patch p of the global batch is patch p // world of rank p % world, pixels in order
    for p in range(24):
        r, j = p % world, p // world
        assert torch.equal(parts[r][0][0, 4 * j:4 * j + 4], batch[0][0, 4 * p:4 * p + 4])
        assert torch.equal(parts[r][2][4 * j:4 * j + 4], batch[2][4 * p:4 * p + 4])
    assert bench.shard_batch(batch, 0, 1) is batch


def test_batch_layout_matches_reference_dataset_contract():
    uv, obj, rgb = bench.make_batch(3, num_pixels=256, num_rays=8)
    assert uv.shape == (1, 256, 8, 2) and obj.shape == (1, 256) and rgb.shape == (256, 3)
    base = torch.floor(uv[0].mean(1)).reshape(64, 4, 2)
    # 2x2 patches (the normal-smoothness loss needs them intact, loss.py:255-264)
    assert torch.equal(base[:, 1] - base[:, 0], torch.tensor([1.0, 0.0]).expand(64, 2))
    assert torch.equal(base[:, 2] - base[:, 0], torch.tensor([0.0, 1.0]).expand(64, 2))
    # one shared jitter set for all pixels (scene_dataset.py:212-216)
    d = uv[0, 1:] - uv[0, :1]
    assert (d - d[:, :1]).abs().max().item() < 1e-4
