"""Full-frame rendering around the hot path: the reference's split_input / merge_output helpers
(code/utils/general.py:24-37,68-82, same names and argument meaning) and the multi-GPU frame loop of
scripts/render.py:283-360 with its pickle-through-CPU `gather_object` replaced by ONE fixed-shape NCCL gather.

Work split: pixel p belongs to rank p % world (the reference deals whole 2**level-ray chunks round robin, render.py:288-295;
dealing single pixels gives every rank the same pixel count and the same hit density, so no rank waits for another).  A rank
renders its pixels in forwards of at most 2**memory_capacity_level rays (one forward when they fit), packs the 11 output
planes the renderer keeps (render.py:311-323) into one [pixels_per_rank, 27] float tensor; rank 0 receives all of them with a
single `dist.gather` and un-interleaves them on the device (a transpose).
"""
import torch

# the planes scripts/render.py keeps per chunk (render.py:311-323) and their channel counts
FRAME_PLANES = (('points', 3), ('idr_rgb_values', 3), ('sg_rgb_values', 3), ('network_object_mask', 1), ('object_mask', 1),
                ('normal_values', 3), ('sg_diffuse_albedo_values', 3), ('sg_diffuse_rgb_values', 3), ('sg_specular_rgb_values', 3),
                ('sg_roughness_values', 1), ('sg_specular_reflection_values', 3))
FRAME_CHANNELS = sum(c for _, c in FRAME_PLANES)
_BOOL_PLANES = ('network_object_mask', 'object_mask')


def split_input(model_input, total_pixels, num_rays=1, memory_capacity_level=18):
    """general.py:24-37: list of inputs whose uv / object_mask hold at most 2**level / num_rays pixels each
    (tensors stay on the device of model_input['uv'] instead of the reference's hard-coded .cuda())."""
    max_num = 2 ** memory_capacity_level
    n_pixels = max_num // num_rays if num_rays > 0 else max_num
    split = []
    idx = torch.arange(total_pixels, device=model_input['uv'].device)
    for indx in torch.split(idx, int(n_pixels), dim=0):
        data = model_input.copy()
        data['uv'] = torch.index_select(model_input['uv'], 1, indx)
        data['object_mask'] = torch.index_select(model_input['object_mask'], 1, indx)
        split.append(data)
    return split


def merge_output(res, total_pixels, batch_size):
    """general.py:68-82: concatenate the per-chunk output dicts back into whole-frame tensors."""
    model_outputs = {}
    for entry in res[0]:
        if res[0][entry] is None:
            continue
        if len(res[0][entry].shape) == 1:
            model_outputs[entry] = torch.cat([r[entry].reshape(batch_size, -1, 1) for r in res], 1).reshape(batch_size * total_pixels)
        else:
            model_outputs[entry] = torch.cat([r[entry].reshape(batch_size, -1, r[entry].shape[-1]) for r in res],
                                             1).reshape(batch_size * total_pixels, -1)
    return model_outputs


def pack_planes(out, n_pixels, pad_to):
    """one chunk's 11 planes -> [pad_to, 27] float32 (masks as 0/1; rows past n_pixels are zero)"""
    dev = out['points'].device
    buf = torch.zeros(pad_to, FRAME_CHANNELS, device=dev)
    c0 = 0
    for name, c in FRAME_PLANES:
        buf[:n_pixels, c0:c0 + c] = out[name].detach().reshape(n_pixels, c).to(torch.float32)
        c0 += c
    return buf


def unpack_planes(buf, n_pixels):
    out, c0 = {}, 0
    for name, c in FRAME_PLANES:
        v = buf[:n_pixels, c0:c0 + c]
        out[name] = (v[:, 0] > 0.5) if name in _BOOL_PLANES else (v.contiguous() if c > 1 or name == 'sg_roughness_values' else v[:, 0])
        c0 += c
    return out


def render_frame(model, model_input, total_pixels, num_rays=1, memory_capacity_level=18, group=None):
    """Renders every pixel of model_input['uv'] ([1, total_pixels, (R,) 2]) through `model`, sharded over the ranks of `group`
    (None / uninitialised torch.distributed = single process).  Returns the merged dict of FRAME_PLANES on rank 0
    ([total_pixels, c] tensors, masks bool [total_pixels]) and None on the other ranks."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    per_rank = (total_pixels + world - 1) // world
    dev = model_input['uv'].device
    mine = dict(model_input)
    mine['uv'] = model_input['uv'][:, rank::world]
    mine['object_mask'] = model_input['object_mask'][:, rank::world]
    n_mine = mine['uv'].shape[1]
    local = torch.zeros(per_rank, FRAME_CHANNELS, device=dev)
    with torch.no_grad():
        done = 0
        for s in (split_input(mine, n_mine, num_rays, memory_capacity_level) if n_mine > 0 else []):
            n = s['uv'].shape[1]
            out = model(s)
            local[done:done + n] = pack_planes(out, n, n)
            done += n
    if multi:
        gathered = torch.empty(world, per_rank, FRAME_CHANNELS, device=dev) if rank == 0 else None
        dist.gather(local, list(gathered.unbind(0)) if rank == 0 else None, dst=0, group=group)
        if rank != 0:
            return None
        frame = gathered.permute(1, 0, 2).reshape(world * per_rank, FRAME_CHANNELS)      # pixel p = rank + world * i
    else:
        frame = local
    return unpack_planes(frame, total_pixels)
