"""Oracle (test infrastructure): IDR sphere tracing of the reference, restated per ray.

Restates
  * get_sphere_intersection   -- reference code/utils/rend_util.py:200-221
  * RayTracing.sphere_tracing -- code/model/ray_tracing.py:104-193
  * RayTracing.ray_sampler    -- :195-257     * RayTracing.rootfind -- :259-280 (bisection)
  * RayTracing.minimal_sdf_points -- :309-337 * RayTracing.forward  -- :29-101
  * get_camera_params / lift  -- code/utils/rend_util.py:90-142

The reference drives all rays in lock step with boolean-mask gathers; here every ray carries its own
state and masks are applied with torch.where -- the results are identical because a finished ray's
update is a no-op there too (SURVEY.md section 7 "global, data-dependent loop exits").  Two places are
*batch-coupled* in the reference and are kept so: the bisection stops for everybody once no ray has
work left (ray_tracing.py:264,276) and min-SDF sampling shares one draw of 100 uniforms (:316).

"Fixing the march order" (BASELINE.json): 3-vector dot products that the reference evaluates with
torch.bmm are fixed here to ((a0*b0 + a1*b1) + a2*b2) without fused multiply-add; every quantity
that feeds a comparison is an explicit mul followed by an add.  The CUDA kernels follow the same
order, which is what makes hit masks bit-comparable for an analytic SDF.

Parity status: PINNED -- tests/test_oracle_hotpath.py runs this against the real RayTracing module
(with analytic and MLP SDF callables) when /root/reference is present, and against
tests/golden/tracer_*.npz generated from it.
"""
import torch
import torch.nn.functional as F


class TraceConfig:
    """conf.conf:85-94 (model.ray_tracer)."""

    def __init__(self, object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=3,
                 sphere_tracing_iters=10, n_steps=100, n_rootfind_steps=32):
        self.object_bounding_sphere = object_bounding_sphere
        self.sdf_threshold = sdf_threshold
        self.line_search_step = line_search_step
        self.line_step_iters = line_step_iters
        self.sphere_tracing_iters = sphere_tracing_iters
        self.n_steps = n_steps
        self.n_rootfind_steps = n_rootfind_steps

    def as_kwargs(self):
        return dict(self.__dict__)


def dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def along(o, t, d):
    """o + t * d with a rounded multiply then a rounded add (ray_tracing.py:160-161)."""
    return o + t.unsqueeze(-1) * d


# ------------------------------------------------------------------------------------------------
def camera_rays(uv, pose, intrinsics):
    """uv [B,S,2], pose [B,4,4], intrinsics [B,4,4] -> unit ray dirs [B,S,3], cam_loc [B,3]
    (rend_util.py:90-142, pose-matrix branch; depth plane z = 1)."""
    cam_loc = pose[:, :3, 3]
    fx, fy = intrinsics[:, 0, 0].unsqueeze(-1), intrinsics[:, 1, 1].unsqueeze(-1)
    cx, cy = intrinsics[:, 0, 2].unsqueeze(-1), intrinsics[:, 1, 2].unsqueeze(-1)
    sk = intrinsics[:, 0, 1].unsqueeze(-1)
    x, y = uv[:, :, 0], uv[:, :, 1]
    z = torch.ones_like(x)
    x_lift = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_lift = (y - cy) / fy * z
    cam_pts = torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1)          # [B,S,4]
    world = torch.bmm(pose, cam_pts.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return dirs, cam_loc


def sphere_intersection(origin, direction, radius):
    """origin, direction [R,3] -> t [R,2] (near, far; clamp_min 0.01), hits [R]."""
    b = dot3(direction, origin)
    onorm = torch.sqrt((origin[:, 0] * origin[:, 0] + origin[:, 1] * origin[:, 1]) + origin[:, 2] * origin[:, 2])
    under = b ** 2 - (onorm ** 2 - radius ** 2)
    hits = under > 0
    root = torch.sqrt(torch.where(hits, under, torch.zeros_like(under)))
    near = torch.where(hits, root * -1.0 - b, torch.zeros_like(b))
    far = torch.where(hits, root * 1.0 - b, torch.zeros_like(b))
    t = torch.stack([near, far], dim=-1).clamp_min(0.01)
    return t, hits


def _eval_on(sdf, pts, mask):
    out = torch.zeros(pts.shape[0], dtype=pts.dtype, device=pts.device)
    if mask.any():
        out[mask] = sdf(pts[mask]).to(pts.dtype)
    return out


def sphere_tracing(sdf, o, d, hits, t_sphere, cfg):
    """Two-sided sphere tracing.  Returns dict(points, unfinished, acc_start, acc_end, min_dis, max_dis)."""
    thr = cfg.sdf_threshold
    zero = torch.zeros(o.shape[0], dtype=o.dtype, device=o.device)
    unf_s, unf_e = hits.clone(), hits.clone()
    acc_s = torch.where(hits, t_sphere[:, 0], zero)
    acc_e = torch.where(hits, t_sphere[:, 1], zero)
    min_dis, max_dis = acc_s.clone(), acc_e.clone()
    pts_s = torch.where(hits.unsqueeze(-1), along(o, t_sphere[:, 0], d), torch.zeros_like(o))
    pts_e = torch.where(hits.unsqueeze(-1), along(o, t_sphere[:, 1], d), torch.zeros_like(o))
    nxt_s = _eval_on(sdf, pts_s, unf_s)
    nxt_e = _eval_on(sdf, pts_e, unf_e)
    n_evals = int(unf_s.sum() + unf_e.sum())
    it = 0
    while True:
        cur_s = torch.where(unf_s, nxt_s, zero)
        cur_s = torch.where(cur_s <= thr, zero, cur_s)
        cur_e = torch.where(unf_e, nxt_e, zero)
        cur_e = torch.where(cur_e <= thr, zero, cur_e)
        unf_s = unf_s & (cur_s > thr)
        unf_e = unf_e & (cur_e > thr)
        if (not bool(unf_s.any()) and not bool(unf_e.any())) or it == cfg.sphere_tracing_iters:
            break
        it += 1
        acc_s = acc_s + cur_s
        acc_e = acc_e - cur_e
        pts_s = along(o, acc_s, d)
        pts_e = along(o, acc_e, d)
        nxt_s = _eval_on(sdf, pts_s, unf_s)
        nxt_e = _eval_on(sdf, pts_e, unf_e)
        n_evals += int(unf_s.sum() + unf_e.sum())
        bad_s, bad_e = nxt_s < 0, nxt_e < 0
        ls = 0
        while (bool(bad_s.any()) or bool(bad_e.any())) and ls < cfg.line_step_iters:
            back = (1 - cfg.line_search_step) / (2 ** ls)
            acc_s = torch.where(bad_s, acc_s - back * cur_s, acc_s)
            pts_s = torch.where(bad_s.unsqueeze(-1), along(o, acc_s, d), pts_s)
            acc_e = torch.where(bad_e, acc_e + back * cur_e, acc_e)
            pts_e = torch.where(bad_e.unsqueeze(-1), along(o, acc_e, d), pts_e)
            nxt_s = torch.where(bad_s, _eval_on(sdf, pts_s, bad_s), nxt_s)
            nxt_e = torch.where(bad_e, _eval_on(sdf, pts_e, bad_e), nxt_e)
            n_evals += int(bad_s.sum() + bad_e.sum())
            bad_s, bad_e = nxt_s < 0, nxt_e < 0
            ls += 1
        crossed = acc_s < acc_e
        unf_s = unf_s & crossed
        unf_e = unf_e & crossed
    return dict(points=pts_s, unfinished=unf_s, acc_start=acc_s, acc_end=acc_e, min_dis=min_dis, max_dis=max_dis,
                n_evals=n_evals, entered_loop=it > 0)


def bisection(sdf, s_lo, s_hi, z_lo, z_hi, o, d, cfg):
    """RayTracing.rootfind: bisection on every passed ray while ANY ray still has work (batch-coupled)."""
    work = (s_lo > 0) & (s_hi < 0) & (z_hi > z_lo)
    z_mid = (z_lo + z_hi) / 2.
    i, n_evals = 0, 0
    while bool(work.any()) and i < cfg.n_rootfind_steps:
        s_mid = sdf(along(o, z_mid, d)).to(z_mid.dtype)
        n_evals += z_mid.shape[0]
        pos = s_mid > 0
        neg = s_mid <= 0
        z_lo = torch.where(pos, z_mid, z_lo)
        s_lo = torch.where(pos, s_mid, s_lo)
        z_hi = torch.where(neg, z_mid, z_hi)
        s_hi = torch.where(neg, s_mid, s_hi)
        z_mid = (z_lo + z_hi) / 2.
        work = work & ((z_hi - z_lo) > 1e-6)
        i += 1
    return z_mid, n_evals


def ray_sampler(sdf, o, d, object_mask, sampler_mask, t_lo, t_hi, cfg, training):
    """100 uniform samples per unconverged ray, first sign change, bisection.  Returns full-size
    (points, net_object_mask, dists) that are meaningful on sampler rays only."""
    R = o.shape[0]
    n = cfg.n_steps
    idx = torch.nonzero(sampler_mask).flatten()
    oo, dd = o[idx], d[idx]
    a, b = t_lo[idx], t_hi[idx]
    frac = torch.linspace(0, 1, steps=n, device=o.device, dtype=o.dtype)
    ts = a.unsqueeze(-1) + frac.view(1, -1) * (b - a).unsqueeze(-1)                       # [m, n]
    pts = oo.unsqueeze(1) + ts.unsqueeze(-1) * dd.unsqueeze(1)                            # [m, n, 3]
    vals = sdf(pts.reshape(-1, 3)).to(o.dtype).reshape(-1, n)
    n_evals = vals.numel()
    rank = torch.arange(n, 0, -1, device=o.device, dtype=o.dtype).view(1, n)
    first = torch.argmin(torch.sign(vals) * rank, -1)                                      # first negative sample
    rows = torch.arange(idx.shape[0], device=o.device)
    out_pts = torch.zeros(R, 3, dtype=o.dtype, device=o.device)
    out_t = torch.zeros(R, dtype=o.dtype, device=o.device)
    sel_pts, sel_t = pts[rows, first], ts[rows, first]
    inside_gt = object_mask[idx]
    inside_net = vals[rows, first] < 0
    fallback = ~(inside_gt & inside_net)            # P_out pixels: take the sample of minimal SDF instead
    lowest = torch.argmin(vals, -1)
    sel_pts = torch.where(fallback.unsqueeze(-1), pts[rows, lowest], sel_pts)
    sel_t = torch.where(fallback, ts[rows, lowest], sel_t)
    net_mask = sampler_mask.clone()
    net_mask[idx[~inside_net]] = False
    refine = (inside_net & inside_gt) if training else inside_net
    if bool(refine.any()):
        r = torch.nonzero(refine).flatten()
        hi_i = first[r]
        lo_i = (hi_i - 1) % n                       # python's negative index: sample 0 pairs with sample n-1
        z, ne = bisection(sdf, vals[r, lo_i].clone(), vals[r, hi_i].clone(), ts[r, lo_i].clone(), ts[r, hi_i].clone(),
                          oo[r], dd[r], cfg)
        n_evals += ne
        sel_pts = sel_pts.clone()
        sel_t = sel_t.clone()
        sel_pts[r] = along(oo[r], z, dd[r])
        sel_t[r] = z
    out_pts[idx] = sel_pts
    out_t[idx] = sel_t
    return out_pts, net_mask, out_t, n_evals


def minimal_sdf_points(sdf, o, d, mask, min_dis, max_dis, uniforms, cfg):
    """Train only: n_steps random depths per masked ray (one shared draw), keep the SDF minimum."""
    idx = torch.nonzero(mask).flatten()
    lo, hi = min_dis[idx].unsqueeze(-1), max_dis[idx].unsqueeze(-1)
    steps = uniforms.to(o.dtype).to(o.device).view(1, -1).repeat(idx.shape[0], 1) * (hi - lo) + lo
    pts = o[idx].unsqueeze(1) + steps.unsqueeze(-1) * d[idx].unsqueeze(1)
    vals = sdf(pts.reshape(-1, 3)).to(o.dtype).reshape(-1, cfg.n_steps)
    best = vals.argmin(-1)
    rows = torch.arange(idx.shape[0], device=o.device)
    return pts[rows, best], steps[rows, best], vals.numel()


def ray_trace(sdf, cam_loc, object_mask, ray_directions, cfg, training=False, uniforms=None):
    """RayTracing.forward.  cam_loc [B,3], object_mask [B*P] bool, ray_directions [B,P,3]
    -> points [B*P,3], network_object_mask [B*P], dists [B*P]  (+ a dict of statistics)."""
    B, P, _ = ray_directions.shape
    o = cam_loc.unsqueeze(1).expand(B, P, 3).reshape(-1, 3)
    d = ray_directions.reshape(-1, 3)
    t_sph, hits = sphere_intersection(o, d, cfg.object_bounding_sphere)
    st = sphere_tracing(sdf, o, d, hits, t_sph, cfg)
    pts, acc_s, acc_e = st["points"].clone(), st["acc_start"].clone(), st["acc_end"]
    if not st["entered_loop"]:
        pass  # reference: points of sphere-missing rays stay 0 when the loop never ran (don't-care lanes)
    net_mask = acc_s < acc_e
    sampler_mask = st["unfinished"]
    stats = dict(n_evals=st["n_evals"], n_sampler=int(sampler_mask.sum()), sphere_hits=hits)
    if bool(sampler_mask.any()):
        s_pts, s_net, s_t, ne = ray_sampler(sdf, o, d, object_mask, sampler_mask, acc_s, acc_e, cfg, training)
        stats["n_evals"] += ne
        pts[sampler_mask] = s_pts[sampler_mask]
        acc_s[sampler_mask] = s_t[sampler_mask]
        net_mask[sampler_mask] = s_net[sampler_mask]
    if not training:
        return pts, net_mask, acc_s, stats
    in_mask = ~net_mask & object_mask & ~sampler_mask
    out_mask = ~object_mask & ~sampler_mask
    left_out = (in_mask | out_mask) & ~hits
    if bool(left_out.any()):
        t_close = -dot3(d[left_out], o[left_out])
        acc_s[left_out] = t_close
        pts[left_out] = along(o[left_out], t_close, d[left_out])
    mask = (in_mask | out_mask) & hits
    stats["n_min_sdf"] = int(mask.sum())
    if bool(mask.any()):
        min_dis, max_dis = st["min_dis"].clone(), st["max_dis"]
        sel = net_mask & out_mask
        min_dis[sel] = acc_s[sel]
        m_pts, m_t, ne = minimal_sdf_points(sdf, o, d, mask, min_dis, max_dis, uniforms, cfg)
        stats["n_evals"] += ne
        pts[mask] = m_pts
        acc_s[mask] = m_t
    return pts, net_mask, acc_s, stats


# ------------------------------------------------------------------------------------------------
# Analytic test scene: union of spheres and axis-aligned boxes inside the unit sphere ("robot-scale").
# Only +,-,*,sqrt,min,max,abs in a fixed order, so torch and the CUDA test evaluator agree bit for bit.
# ------------------------------------------------------------------------------------------------
def robot_scene():
    """[n,8] rows: kind (0 sphere, 1 box), cx, cy, cz, then radius,0,0 or half extents hx,hy,hz, pad."""
    prims = [
        [1, 0.00, 0.05, 0.00, 0.22, 0.28, 0.14, 0],    # torso
        [0, 0.00, 0.48, 0.00, 0.16, 0.00, 0.00, 0],    # head
        [1, -0.34, 0.10, 0.00, 0.07, 0.24, 0.07, 0],   # arms
        [1, 0.34, 0.10, 0.00, 0.07, 0.24, 0.07, 0],
        [1, -0.12, -0.48, 0.00, 0.08, 0.24, 0.08, 0],  # legs
        [1, 0.12, -0.48, 0.00, 0.08, 0.24, 0.08, 0],
        [0, -0.34, -0.20, 0.00, 0.09, 0.00, 0.00, 0],  # hands
        [0, 0.34, -0.20, 0.00, 0.09, 0.00, 0.00, 0],
        [0, 0.00, 0.05, 0.17, 0.08, 0.00, 0.00, 0],    # chest light
    ]
    return torch.tensor(prims, dtype=torch.float32)


def analytic_sdf(prims):
    """Returns an `sdf(points[n,3]) -> [n]` callable for a primitive table (see robot_scene)."""

    def f(p):
        pr = prims.to(p.device, p.dtype)
        best = None
        for i in range(pr.shape[0]):
            kind = int(pr[i, 0].item())
            q0, q1, q2 = p[:, 0] - pr[i, 1], p[:, 1] - pr[i, 2], p[:, 2] - pr[i, 3]
            if kind == 0:
                val = torch.sqrt((q0 * q0 + q1 * q1) + q2 * q2) - pr[i, 4]
            else:
                a0, a1, a2 = q0.abs() - pr[i, 4], q1.abs() - pr[i, 5], q2.abs() - pr[i, 6]
                m0, m1, m2 = a0.clamp_min(0), a1.clamp_min(0), a2.clamp_min(0)
                outside = torch.sqrt((m0 * m0 + m1 * m1) + m2 * m2)
                inside = torch.maximum(a0, torch.maximum(a1, a2)).clamp_max(0)
                val = outside + inside
            best = val if best is None else torch.minimum(best, val)
        return best

    return f
