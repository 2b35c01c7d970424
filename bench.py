#!/usr/bin/env python
"""bench.py -- shaded rays/s of the NeFII per-ray-batch rendering hot path on B200.

Workload (BASELINE.json configs[2], the config the metric "shaded rays/sec (primary+indirect, fwd+bwd)" is quoted
on): one step-2 training iteration of conf.conf -- num_pixels=2048 (512 2x2 patches of a synthetic 800x800 view)
x num_rays=64 = 131072 primary rays per GPU, 128 light SGs, 8x512 SDF MLP (PE 6, frozen), near-field indirect
illumination with 3 importance-sampled secondary rays per hit, forward + loss + backward + both Adam steps.
Random-init weights of that architecture (the reference's own initialisers), synthetic data.

  python bench.py --gpus N --steps K --warmup W            # N=1 default; N>1: launched by torchrun, one rank per GPU
  python bench.py --impl reference ...                     # the reference's algorithm on the host CPU (oracle port)

One JSON line on stdout (rank 0).  value = (primary + secondary rays of all ranks) / max-over-ranks device time.
Beside the contract's keys the line carries `roofline` (the tcgen05 layer GEMM, timed live with CUDA events around every launch),
`cpu_baseline` (oracle port on the host cores, bounded sample) and three extras: `ms_per_frame_800x800` (chunked, rank-sharded
novel-view render with the gather onto rank 0), `ms_per_secondary_training_pass` (the trainer's every-10th-step pass through
forward(with_point=True)) and `gpu_launches` (kernels of this library inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "shaded rays/sec (primary+indirect, fwd+bwd)"
NUM_PIXELS, NUM_RAYS, NUM_SGS, IMG = 2048, 64, 128, 800
SDF_FLOPS_PER_POINT = 3.671e6      # 1,835,520 MAC forward


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------------------------------------------
# synthetic scene / inputs
# --------------------------------------------------------------------------------------------------------------
def make_camera():
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 2.4 * IMG      # the init-sphere covers ~50 % of the frame
    K[0, 2] = K[1, 2] = IMG / 2
    pose = torch.eye(4)
    pose[:3, 3] = torch.tensor([0.0, 0.0, -3.0])
    return pose[None], K[None]


def make_batch(seed, num_pixels=NUM_PIXELS, num_rays=NUM_RAYS, img=IMG):
    """What SceneDataset hands the trainer for one iteration (scene_dataset.py:149-177,212-251): num_pixels/4 random
    2x2 patches, one shared set of num_rays sub-pixel jitters, all-true object mask, GT rgb."""
    g = torch.Generator().manual_seed(seed)
    n_patch = num_pixels // 4
    x0 = torch.randint(0, img - 1, (n_patch,), generator=g)
    y0 = torch.randint(0, img - 1, (n_patch,), generator=g)
    px = torch.stack([torch.stack([x0, y0], -1), torch.stack([x0 + 1, y0], -1),
                      torch.stack([x0, y0 + 1], -1), torch.stack([x0 + 1, y0 + 1], -1)], 1).reshape(-1, 2).float()
    jitter = torch.rand(num_rays, 2, generator=g) - 0.5
    uv = (px.unsqueeze(1) + 0.5 + jitter.unsqueeze(0)).unsqueeze(0)          # [1, S, R, 2]
    object_mask = torch.ones(1, num_pixels, dtype=torch.bool)
    rgb = torch.rand(num_pixels, 3, generator=g)
    return uv.contiguous(), object_mask, rgb


LOSS_CONF = dict(idr_rgb_weight=1.0, sg_rgb_weight=1.0, eikonal_weight=0.1, mask_weight=100.0, alpha=50.0, normalsmooth_weight=1.0,
                 r_patch=1.0, loss_type='L1', env_loss_type='L2', background_rgb_weight=1.0)     # reference confs_sg/conf.conf:24-35
_crit = None


def idr_loss(out, rgb_gt):
    """The reference's IDRLoss with conf.conf's weights (model/loss.py:278-320) through the fused CUDA loss
    (nefii_b200/model/loss.py: one launch forward, one backward, no host round trips)."""
    global _crit
    if _crit is None:
        from nefii_b200.model.loss import IDRLoss
        _crit = IDRLoss(**LOSS_CONF)
    return _crit(out, {'rgb': rgb_gt})['loss']


def idr_loss_cpu(out, rgb_gt):
    """The same loss for the CPU arm: the oracle's restatement of IDRLoss (plain torch ops, like the reference)."""
    from oracle import loss as oloss
    terms = oloss.idr_loss_terms(out['idr_rgb_values'], out['sg_rgb_values'], rgb_gt, out['normal_values'], out['sdf_output'],
                                 out['network_object_mask'], out['object_mask'], alpha=LOSS_CONF['alpha'], r_patch=1,
                                 loss_type='L1', env_loss_type='L2')
    return oloss.idr_loss(terms)


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index, enabled=True):
        self.samples, self.proc, self.gpu, self.enabled = [], None, gpu_index, enabled

    def __enter__(self):
        if not self.enabled:        # one poller per job (rank 0's GPU): eight NVML pollers contend with the ranks' launch threads
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("NEFII_BENCH_POLL_MS", "200")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def build_model(dev):
    from nefii_b200.model.implicit_differentiable_renderer import IDRNetwork
    from nefii_b200.utils.conf import default_model_conf
    torch.manual_seed(0)
    model = IDRNetwork(default_model_conf(num_lgt_sgs=NUM_SGS)).to(dev)
    model.freeze_geometry()          # run_s2.sh --freeze_geometry
    model.train()
    return model


class FlatGrads:
    """All trainable gradients live in one flat fp32 buffer, so the data-parallel exchange is ONE NCCL all-reduce
    (12.95 MB) instead of DDP's buckets."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, world):
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat)
            self.flat.div_(world)


def run_ours(args):
    from nefii_b200 import _lib
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(dev)
    pose, K = [t.to(dev) for t in make_camera()]
    flat = FlatGrads(model.parameters())
    idr_params = [p for p in model.rendering_network.parameters() if p.requires_grad]
    sg_params = [p for p in model.envmap_material_network.parameters() if p.requires_grad]
    opt_idr = torch.optim.Adam(idr_params, lr=5e-4)
    opt_sg = torch.optim.Adam(sg_params, lr=5e-4)
    n_steps = args.steps + args.warmup
    # every rank renders its own pixel batch (weak scaling, as DDP with a fixed per-GPU num_pixels)
    host_batches = [make_batch(1000 * rank + i) for i in range(n_steps)]
    pinned = [[t.pin_memory() for t in b] for b in host_batches]
    dev_batches = [[t.to(dev) for t in b] for b in host_batches]
    ray_count = torch.zeros(1, device=dev, dtype=torch.float64)

    def step(uv, obj, rgb, count_rays=True):
        flat.zero()
        out = model({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
        loss = idr_loss(out, rgb)
        loss.backward()
        flat.all_reduce(world)
        opt_idr.step()
        opt_sg.step()
        if count_rays:
            n_hit = out['secondary_mask'].shape[1] if out['secondary_mask'] is not None else 0
            ray_count.add_(uv.shape[1] * uv.shape[2] + 3 * n_hit)
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def sum_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t)
        return t.item()

    # ---- device-resident timing ------------------------------------------------------------------------------
    for i in range(args.warmup):
        step(*dev_batches[i], count_rays=False)
    barrier()
    launches0 = int(_lib.raw().nefii_launch_count())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, enabled=(rank == 0)) as clocks:
        ev0.record()
        for i in range(args.warmup, n_steps):
            step(*dev_batches[i])
        ev1.record()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = int(_lib.raw().nefii_launch_count()) - launches0
    rays = sum_over_ranks(ray_count.item())
    value = rays / (ms * 1e-3)

    # ---- end to end through the public API: pinned host inputs -> device, loss read back, every step -------------
    ray_count.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, n_steps):
        uv, obj, rgb = [t.to(dev, non_blocking=True) for t in pinned[i]]
        loss = step(uv, obj, rgb)
        _ = loss.item()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    rays_e2e = sum_over_ranks(ray_count.item())
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])

    # ---- roofline of the dominant kernel (tcgen05 layer GEMM): per-launch CUDA events over one more step ---------
    lib = _lib.raw()
    lib.nefii_gemm_profile_enable(1)
    step(*dev_batches[-1], count_rays=False)
    import ctypes
    out3 = (ctypes.c_double * 3)()
    lib.nefii_gemm_profile_fetch(out3)
    lib.nefii_gemm_profile_enable(0)
    gemm_ms, gemm_flops, gemm_launches = out3[0], out3[1], int(out3[2])
    e_prof0, e_prof1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e_prof0.record()
    step(*dev_batches[-1], count_rays=False)
    e_prof1.record()
    torch.cuda.synchronize()
    step_ms = e_prof0.elapsed_time(e_prof1)

    # ---- extra: the secondary-training pass the reference's trainer runs every `secondary_train_interval` = 10 steps
    # (idr_train.py:804-852: <= secondary_batch_size 1024 secondary hit points x num_rays 64 directions through
    # forward(with_point=True), L1 between the SG and the radiance-field colour, backward, both optimizers) ----
    secondary_ms = None
    try:
        with torch.no_grad():
            out = model({'uv': dev_batches[-1][0], 'object_mask': dev_batches[-1][1], 'pose': pose, 'intrinsics': K})
        sp, sm, sd = out['secondary_points'].reshape(-1, 3), out['secondary_mask'].reshape(-1), out['secondary_dir'].reshape(-1, 3)
        pts, dirs = sp[sm][:1024], sd[sm][:1024]
        n_sec = pts.shape[0]
        sec_in = {'points': pts.unsqueeze(1).expand(n_sec, NUM_RAYS, 3).contiguous(),
                  'ray_dirs': dirs.unsqueeze(1).expand(n_sec, NUM_RAYS, 3).contiguous()}

        def secondary_step():
            flat.zero()
            ret = model(sec_in, with_point=True)
            loss = torch.nn.functional.l1_loss(ret['sg_rgb_values'], ret['idr_rgb_values'])
            loss.backward()
            flat.all_reduce(world)
            opt_idr.step()
            opt_sg.step()
        if n_sec > 0:
            secondary_step()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            secondary_step()
            s1.record()
            barrier()
            secondary_ms = max_over_ranks(s0.elapsed_time(s1))
    except Exception as exc:            # an extra: never let it take the headline line down
        print("secondary-training pass failed: %r" % (exc,), file=sys.stderr)

    # ---- extra (BASELINE metric tail "ms/frame 800x800"): novel-view render, eval mode, 1 ray per pixel, this rank's share ----
    frame_ms = None
    try:
        model.eval()
        from nefii_b200.utils import general
        ii, jj = torch.meshgrid(torch.arange(IMG).float(), torch.arange(IMG).float(), indexing="xy")
        uv_host = (torch.stack([ii, jj], -1).reshape(1, -1, 2) + 0.5).pin_memory()
        obj_all = torch.ones(1, IMG * IMG, dtype=torch.bool, device=dev)

        def render_frame():
            # scripts/render.py:283-360: uv upload, 2**18-ray chunks dealt round robin to the ranks, the 11 output planes
            # gathered onto rank 0 (one fixed-shape NCCL gather)
            uv_all = uv_host.to(dev, non_blocking=True)
            return general.render_frame(model, {'uv': uv_all, 'object_mask': obj_all, 'pose': pose, 'intrinsics': K}, IMG * IMG,
                                        num_rays=1, memory_capacity_level=18)
        render_frame()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        frame = render_frame()
        f1.record()
        barrier()
        del frame
        frame_ms = max_over_ranks(f0.elapsed_time(f1))
    finally:
        model.train()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))     # kernel timed inside a long step -> sustained figure
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {
        "kernel": "gemm_split_bf16_kernel (tcgen05 kind::f16, 3 MMAs per fp32 product)",
        "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf else None,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
        "tensor_issue_frac": 3 * achieved_tf / peak_tf if peak_tf else None,
        "flops_per_launch_avg": gemm_flops / max(gemm_launches, 1), "launches_per_step": gemm_launches,
        "kernel_share_of_step": gemm_ms / step_ms if step_ms > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of a 131072-row hidden-layer launch
        # (profiles/r1_gemm_ncu_final.md): 500.5 MB = 3818 B/row against 4096 B/row algorithmic (hi+lo bf16 planes in and out);
        # scaled to this run's average launch so that it is per launch like `achieved`
        "traffic": NCU_DRAM_BYTES_PER_ROW * (gemm_flops / max(gemm_launches, 1)) / (2.0 * 512 * 512),
        "traffic_source": "ncu dram bytes per row of the hidden-layer GEMM (3818 B/row, algorithmic 4096 B/row) x rows of the average launch",
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_throughput(px=1024, rays=8, steps=1, warmup=0)     # ~16k rays: 10-30 s of host time
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (MLPs: bf16x3 split on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "step-2 training iteration (BASELINE configs[2]): num_pixels=2048 x num_rays=64 = 131072 "
                                   "primary rays/GPU + 3 secondary rays per hit, 128 SGs, 8x512 SDF MLP frozen, fwd+loss+bwd+Adam",
                       "rays_per_step": rays / args.steps, "primary_rays_per_step_per_gpu": NUM_PIXELS * NUM_RAYS,
                       "l2": "per-step working set (MLP activations, several GB) >> 126 MB L2; a different pixel batch every step",
                       "parallelism": "dp%d (rays sharded by rank, one flat NCCL all-reduce of 3.2M grads)" % world},
            "e2e": {"value": rays_e2e / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "ms_per_frame_800x800": frame_ms,
            "ms_per_secondary_training_pass": secondary_ms,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------------------
NCU_DRAM_BYTES_PER_ROW = 3818.0


def cpu_port_throughput(px, rays, steps, warmup):
    """Times oracle/pipeline.py (the restated reference, plain torch CPU ops) on a bounded sample of the same
    workload: `px` pixels x `rays` rays, forward + loss + backward, all host threads."""
    from oracle import pipeline, ref_harness as rh
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    om = rh.small_model(seed=0, n_sg=NUM_SGS, bumps=0.0)
    om.lgtSGs.requires_grad_(True)
    om.material.requires_grad_(True)
    om.radiance.requires_grad_(True)
    pose, K = make_camera()
    total_rays, total_t = 0, 0.0
    for i in range(warmup + steps):
        uv, obj, rgb = make_batch(i, num_pixels=px, num_rays=rays)
        g = torch.Generator().manual_seed(i)
        U = torch.rand(px * rays, 7, generator=g)
        vecs = [torch.rand(100, generator=g) for _ in range(2)]
        t0 = time.perf_counter()
        out = pipeline.forward_with_uv(om, uv, pose, K, obj, lambda n: U[:n], True, vecs[0], vecs[1])
        loss = idr_loss_cpu(out, rgb)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            n_hit = out['secondary_mask'].shape[1] if out['secondary_mask'] is not None else 0
            total_rays += px * rays + 3 * n_hit
            total_t += dt
    return {"value": total_rays / total_t, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d px x %d rays per step, %d step(s), fwd+loss+bwd, torch CPU %d threads "
                      "(oracle/pipeline.py restating the reference)" % (px, rays, steps, cores)}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    n = args.steps + args.warmup
    px = 256 if n > 6 else 512           # ~4k / ~8k rays per step: the whole run stays within a few minutes
    cpu = cpu_port_throughput(px=px, rays=8, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "rays/s", "n_gpus": env_int("WORLD_SIZE", 1),
        "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "step-2 training iteration (BASELINE configs[2]) on the host CPU, bounded sample: " + cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
