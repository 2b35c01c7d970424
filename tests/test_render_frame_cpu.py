"""CPU: the chunked / sharded frame loop (nefii_b200/utils/general.py, reference utils/general.py:24-82 and
scripts/render.py:283-360) with a stand-in model: split/merge round trip, ragged last chunk, and the world_size-2 gloo
gather (pixels dealt to the ranks p % world, an odd pixel count so that the ranks' shares differ) giving rank 0 exactly the
single-process frame."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nefii_b200.utils import general      # host logic only: importing it does not load the CUDA library


class FakeModel:
    """every plane is a fixed function of the pixel coordinate, so any mis-ordered chunk shows up"""

    def __call__(self, inp):
        uv = inp['uv'].reshape(-1, 2)
        n = uv.shape[0]
        key = uv[:, :1] * 0.001 + uv[:, 1:2] * 0.01
        out = {}
        for j, (name, c) in enumerate(general.FRAME_PLANES):
            v = (key + j).expand(n, c).clone()
            if name in ('network_object_mask', 'object_mask'):
                out[name] = (uv[:, 0].long() + j) % 3 == 0
            elif c == 1:
                out[name] = v
            else:
                out[name] = v + torch.arange(c).float()
        return out


def _frame(n=1000):
    g = torch.Generator().manual_seed(0)
    uv = torch.rand(1, n, 2, generator=g) * 100
    return {'uv': uv, 'object_mask': torch.ones(1, n, dtype=torch.bool), 'pose': torch.eye(4)[None], 'intrinsics': torch.eye(4)[None]}


def test_split_merge_roundtrip_and_ragged_chunks():
    inp = _frame(1000)
    split = general.split_input(inp, 1000, num_rays=1, memory_capacity_level=8)     # 256-pixel chunks, last one 232
    assert [s['uv'].shape[1] for s in split] == [256, 256, 256, 232]
    whole = FakeModel()(inp)
    got = general.render_frame(FakeModel(), inp, 1000, memory_capacity_level=8)
    for name, c in general.FRAME_PLANES:
        ref = whole[name]
        assert got[name].shape[0] == 1000
        assert torch.equal(got[name].reshape(ref.shape), ref) if ref.dtype == torch.bool else torch.allclose(got[name].reshape(ref.shape), ref), name
    assert general.split_input(inp, 1000, num_rays=4, memory_capacity_level=8)[0]['uv'].shape[1] == 64   # general.py:29-30


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inp = _frame(1001)
        out = general.render_frame(FakeModel(), inp, 1001, memory_capacity_level=8)     # 501 / 500 pixels per rank, 256-pixel forwards
        if rank == 0:
            q.put({k: v.clone() for k, v in out.items()})
        else:
            assert out is None
            q.put(None)
    finally:
        dist.destroy_process_group()


def test_world2_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = next(r for r in res if r is not None)
    whole = FakeModel()(_frame(1001))
    for name, c in general.FRAME_PLANES:
        ref = whole[name]
        assert got[name].shape[0] == 1001
        if ref.dtype == torch.bool:
            assert torch.equal(got[name].reshape(ref.shape), ref), name
        else:
            assert torch.allclose(got[name].reshape(ref.shape), ref), name
