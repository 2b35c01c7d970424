"""Ray set-up math of the hot path (reference code/utils/rend_util.py:90-142).

`get_camera_params` keeps the reference's signature.  On CUDA tensors the per-ray part (lift, pose transform, normalise) is
ONE kernel through the C ABI (nefii_camera_rays, csrc/sample_network.cu) instead of ~15 tensor ops and a materialised
[B,4,S] array; the [B,7] quaternion form of the pose is turned into a matrix first (a handful of [B]-sized ops).  CPU
tensors and poses that require grad take the reference's tensor-op path below (`get_camera_params_torch`: tests, tooling,
--train_cameras).  The ray/sphere intersection (rend_util.py:200-221) lives inside the CUDA tracer."""
import torch
import torch.nn.functional as F

# summation order of the 4-term products of torch.bmm(pose, cam_points) the kernel reproduces (csrc/sample_network.cu)
BMM_ORDER = 0


def lift(x, y, z, intrinsics):
    fx = intrinsics[:, 0, 0].unsqueeze(-1)
    fy = intrinsics[:, 1, 1].unsqueeze(-1)
    cx = intrinsics[:, 0, 2].unsqueeze(-1)
    cy = intrinsics[:, 1, 2].unsqueeze(-1)
    sk = intrinsics[:, 0, 1].unsqueeze(-1)
    x_lift = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_lift = (y - cy) / fy * z
    return torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1)


def quat_to_rot(q):
    batch_size, _ = q.shape
    q = F.normalize(q, dim=1)
    R = torch.ones((batch_size, 3, 3), device=q.device, dtype=q.dtype)
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (qj ** 2 + qk ** 2)
    R[:, 0, 1] = 2 * (qj * qi - qk * qr)
    R[:, 0, 2] = 2 * (qi * qk + qr * qj)
    R[:, 1, 0] = 2 * (qj * qi + qk * qr)
    R[:, 1, 1] = 1 - 2 * (qi ** 2 + qk ** 2)
    R[:, 1, 2] = 2 * (qj * qk - qi * qr)
    R[:, 2, 0] = 2 * (qk * qi - qj * qr)
    R[:, 2, 1] = 2 * (qj * qk + qi * qr)
    R[:, 2, 2] = 1 - 2 * (qi ** 2 + qj ** 2)
    return R


def pose_matrix(pose):
    """[B,4,4] as is, or [B,7] quaternion + translation -> [B,4,4] (rend_util.py:91-96)"""
    if pose.shape[1] == 7:
        p = torch.eye(4, device=pose.device, dtype=pose.dtype).repeat(pose.shape[0], 1, 1)
        p[:, :3, :3] = quat_to_rot(pose[:, :4])
        p[:, :3, 3] = pose[:, 4:]
        return p
    return pose


def get_camera_params(uv, pose, intrinsics, out_dirs=None):
    """uv [B,S,2], pose [B,4,4] (or [B,7] quaternion + translation), intrinsics [B,4,4]
    -> ray_dirs [B,S,3] (unit), cam_loc [B,3].  out_dirs: optional [B,S,3] float32 CUDA buffer to write into."""
    if not uv.is_cuda or pose.requires_grad or intrinsics.requires_grad or uv.requires_grad:
        return get_camera_params_torch(uv, pose, intrinsics)
    from .. import _lib
    p = _lib.f32c(pose_matrix(pose))
    k = _lib.f32c(intrinsics)
    u = _lib.f32c(uv)
    B, S, _ = u.shape
    dirs = out_dirs if out_dirs is not None else torch.empty(B, S, 3, device=u.device, dtype=torch.float32)
    cam = torch.empty(B, 3, device=u.device, dtype=torch.float32)
    with torch.cuda.device(u.device):
        _lib.check(_lib.raw().nefii_camera_rays(_lib.stream_ptr(u.device), B, S, u.data_ptr(), p.data_ptr(), k.data_ptr(), BMM_ORDER,
                                                dirs.data_ptr(), cam.data_ptr()))
    return dirs, cam


def get_camera_params_torch(uv, pose, intrinsics):
    """the reference's tensor-op form (rend_util.py:90-117)"""
    if pose.shape[1] == 7:
        cam_loc = pose[:, 4:]
        R = quat_to_rot(pose[:, :4])
        p = torch.eye(4, device=pose.device, dtype=pose.dtype).repeat(pose.shape[0], 1, 1)
        p[:, :3, :3] = R
        p[:, :3, 3] = cam_loc
    else:
        cam_loc = pose[:, :3, 3]
        p = pose
    batch_size, num_samples, _ = uv.shape
    depth = torch.ones((batch_size, num_samples), device=uv.device, dtype=uv.dtype)
    x_cam = uv[:, :, 0].view(batch_size, -1)
    y_cam = uv[:, :, 1].view(batch_size, -1)
    z_cam = depth.view(batch_size, -1)
    pixel_points_cam = lift(x_cam, y_cam, z_cam, intrinsics=intrinsics).permute(0, 2, 1)
    world_coords = torch.bmm(p, pixel_points_cam).permute(0, 2, 1)[:, :, :3]
    ray_dirs = world_coords - cam_loc[:, None, :]
    ray_dirs = F.normalize(ray_dirs, dim=2)
    return ray_dirs, cam_loc
