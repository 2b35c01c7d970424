"""Chunked full-frame rendering around the hot path: the reference's split_input / merge_output helpers
(code/utils/general.py:24-37,68-82, same names and argument meaning) and the multi-GPU frame loop of
scripts/render.py:283-360 with its pickle-through-CPU `gather_object` replaced by ONE fixed-shape NCCL gather.

Work split (render.py:288-295): the frame is cut into chunks of 2**memory_capacity_level rays; chunk c belongs to rank
c % world (round robin, balances hit density over the image).  Every rank renders its chunks, packs the 11 output
planes the renderer keeps (render.py:311-323) into one [n_chunks_local, chunk_pixels, 27] float tensor, and rank 0
receives all of them with a single `dist.gather`, un-interleaves the chunks and returns the merged dict.
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
        out[name] = (v[:, 0] > 0.5) if name in _BOOL_PLANES else (v if c > 1 or name == 'sg_roughness_values' else v[:, 0])
        c0 += c
    return out


def render_frame(model, model_input, total_pixels, num_rays=1, memory_capacity_level=18, group=None):
    """Renders every pixel of model_input['uv'] ([1, total_pixels, (R,) 2]) through `model` in chunks, sharded over the
    ranks of `group` (None / uninitialised torch.distributed = single process).  Returns the merged dict of FRAME_PLANES
    on rank 0 ([total_pixels, c] tensors, masks bool [total_pixels]) and None on the other ranks."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    level = memory_capacity_level
    if multi:
        w, shift = world, 0
        while w > 1:                      # render.py:287: level - floor(log2(world))
            w >>= 1
            shift += 1
        level -= shift
    split = split_input(model_input, total_pixels, num_rays, level)
    chunk_pixels = split[0]['uv'].shape[1]
    mine = split[rank::world]
    per_rank = (len(split) + world - 1) // world
    dev = model_input['uv'].device
    local = torch.zeros(per_rank, chunk_pixels, FRAME_CHANNELS, device=dev)
    with torch.no_grad():
        for i, s in enumerate(mine):
            out = model(s)
            local[i] = pack_planes(out, s['uv'].shape[1], chunk_pixels)
    if multi:
        gathered = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
        dist.gather(local, gathered, dst=0, group=group)
        if rank != 0:
            return None
    else:
        gathered = [local]
    res = []
    for c, s in enumerate(split):                     # chunk c was rendered by rank c % world as its (c // world)-th chunk
        res.append(unpack_planes(gathered[c % world][c // world], s['uv'].shape[1]))
    return merge_output(res, total_pixels, 1)
