#!/usr/bin/env python
"""bench.py -- shaded rays/s of the NeFII per-ray-batch rendering hot path on B200.

Workload (BASELINE.json configs[2], the config the metric "shaded rays/sec (primary+indirect, fwd+bwd)" is quoted on): one
step-2 training iteration of conf.conf -- ONE global batch of num_pixels=2048 (512 2x2 patches of a synthetic 800x800 view)
x num_rays=64 = 131072 primary rays, 128 light SGs, 8x512 SDF MLP (PE 6, frozen), near-field indirect illumination with 3
importance-sampled secondary rays per hit, forward + loss + backward + both Adam steps.  With N GPUs the 512 patches of that
one batch are dealt to the ranks (patch p -> rank p % N), as the reference's trainer does with its global num_pixels sample
(datasets/scene_dataset.py:268-279, training/idr_train.py:653-662): STRONG scaling.  Random-init weights of that architecture
(the reference's initialisers; the SDF network's geometric-init sphere is perturbed into a bumpy blob so that the sampler and
bisection passes have work), synthetic data.

  python bench.py --gpus N --steps K --warmup W            # N=1 default; N>1: launched by torchrun, one rank per GPU
  python bench.py --impl reference ...                     # the reference's algorithm on the host CPU (oracle port)

One JSON line on stdout (rank 0).  value = (primary + secondary rays of all ranks) / max-over-ranks device time.  Beside the
contract's keys the line carries `roofline` (the tcgen05 layer GEMM, timed live with CUDA events around every launch),
`cpu_baseline` / `cpu_baseline_1thread` (oracle port on the host cores, bounded samples), `gpu_eager_baseline` (the same oracle
port in eager PyTorch on this GPU, fp32 matmuls -- what BASELINE.md calls the comparison baseline), `configs` (stand-alone
timings of BASELINE configs[0] and [1] with their roofline fractions), `weak_scaling` (N > 1: every rank its own 2048-pixel
batch), `fitted_scene` (the step on an SDF network L1-fitted to the analytic robot scene with the step-1 trainer path),
`ms_per_frame_800x800`, `cfg4_1080p_256spp` (N = 8), `ms_per_secondary_training_pass`, `gpu_launches` and
`trace_graph_captures_in_timed_region` (trace graphs are captured per secondary-ray bucket; a new bucket costs ~5 ms once).
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

# The hit count -- and with it the size of every tensor after the primary trace -- changes from one pixel batch to the next (72 k .. 107 k
# secondary rays in the bench scene).  torch's default caching allocator answers new sizes with cudaMalloc / cudaFree of whole segments
# (a device synchronisation each: measured 250 ms instead of 180 ms on the steps that hit one); expandable segments grow in place.
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "shaded rays/sec (primary+indirect, fwd+bwd)"
NUM_PIXELS, NUM_RAYS, NUM_SGS, IMG = int(os.environ.get("NEFII_BENCH_PIXELS", 2048)), 64, 128, 800      # env: diagnostics only
SDF_FLOPS_PER_POINT = 3.671e6      # 1,835,520 MAC forward
# 1 / 2: the primary trace of batch i + 1 is enqueued on a side stream next to the whole step of batch i (1) or next to its loss /
# backward / optimizer part (2) (IDRNetwork.prefetch_trace, bit-identical
# results).  Measured (profiles/r2_small_step.md): the trace is not as idle as its launch train suggests -- 20.1 ms alone + 11.1 ms for
# the rest of a 16 384-ray step, 27.3 - 30.2 ms pipelined against 29.6 - 29.7 ms back to back -- so the default stays the reference's order.
PREFETCH = int(os.environ.get("NEFII_BENCH_PREFETCH", "0"))
SCENE_BUMPS = 0.08                 # perturbation of the geometric-init sphere (the parity tests' rough scene)
SG_FLOPS_PER_RAY, SG_BYTES_PER_RAY = 35.0e3, 72.0      # render_with_sg at M = 128, K = 1 (algorithmic)


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------------------------------------------
# synthetic scene / inputs
# --------------------------------------------------------------------------------------------------------------
def make_camera():
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 2.4 * IMG      # the object covers ~50 % of the frame
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


def shard_batch(batch, rank, world):
    """scatter_sampling_idx_patch (scene_dataset.py:268-279) with a strided patch -> rank map: rank r keeps the 2x2 patches
    r, r + world, ... of the global batch (patches stay whole: the normal-smoothness loss needs them, loss.py:255-264)."""
    if world == 1:
        return batch
    uv, obj, rgb = batch
    n_patch = uv.shape[1] // 4
    keep = torch.arange(rank, n_patch, world)
    idx = (keep.unsqueeze(1) * 4 + torch.arange(4).unsqueeze(0)).reshape(-1)
    return uv[:, idx].contiguous(), obj[:, idx].contiguous(), rgb[idx].contiguous()


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


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index, enabled=True):
        self.samples, self.proc, self.gpu, self.enabled = [], None, gpu_index, enabled
        self.first = 0

    def mark(self):
        """samples from here on count (the poller is started EARLY: nvidia-smi's own start-up takes driver locks for a few
        hundred milliseconds and stalls kernel launches -- measured +18 ms per step when it fell inside a 10-step timed loop)"""
        self.first = len(self.samples)

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
        for s in self.samples[self.first:]:
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
# model
# --------------------------------------------------------------------------------------------------------------
def build_model(dev, bumps=SCENE_BUMPS):
    """conf.conf's model with the reference's initialisers (seed 0), built on the host and moved to `dev`.  bumps > 0 gives
    the positional-encoding columns of the SDF network's first layer small random weights: the geometric-init sphere becomes a
    bumpy blob with concavities (rays that graze a bump go through the 100-sample sampler and the bisection)."""
    from nefii_b200.model.implicit_differentiable_renderer import IDRNetwork
    from nefii_b200.utils.conf import default_model_conf
    torch.manual_seed(0)
    model = IDRNetwork(default_model_conf(num_lgt_sgs=NUM_SGS))
    if bumps > 0:
        g = torch.Generator().manual_seed(12345)
        lin0 = model.implicit_network.lin0
        with torch.no_grad():
            out_dim = lin0.weight_v.shape[0]
            lin0.weight_v[:, 3:] = torch.randn(out_dim, lin0.weight_v.shape[1] - 3, generator=g) * (bumps * math.sqrt(2) / math.sqrt(out_dim))
            lin0.weight_g.copy_(lin0.weight_v.norm(dim=1, keepdim=True))      # effective weight == weight_v, as at initialisation
    model = model.to(dev)
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


def ev_ms(fn, iters, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    from nefii_b200 import _lib
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.raw()
    model = build_model(dev)
    pose, K = [t.to(dev) for t in make_camera()]
    flat = FlatGrads(model.parameters())
    idr_params = [p for p in model.rendering_network.parameters() if p.requires_grad]
    sg_params = [p for p in model.envmap_material_network.parameters() if p.requires_grad]
    opt_idr = torch.optim.Adam(idr_params, lr=5e-4)
    opt_sg = torch.optim.Adam(sg_params, lr=5e-4)
    n_steps = args.steps + args.warmup
    # ONE global batch per step, its 512 patches dealt to the ranks (strong scaling)
    host_batches = [shard_batch(make_batch(1000 + i), rank, world) for i in range(n_steps)]
    pinned = [[t.pin_memory() for t in b] for b in host_batches]
    dev_batches = [[t.to(dev) for t in b] for b in host_batches]
    ray_count = torch.zeros(1, device=dev, dtype=torch.float64)

    def step(uv, obj, rgb, count_rays=True, net=None, nxt=None):
        net = net or model
        flat.zero()
        out = net({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
        if nxt is not None and PREFETCH == 2:      # next batch's primary trace next to this batch's loss / backward / optimizer
            model.prefetch_trace({'uv': nxt[0], 'object_mask': nxt[1], 'pose': pose, 'intrinsics': K})
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

    def reduce_ranks(x, op=None):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return t.item()

    def max_over_ranks(x):
        import torch.distributed as dist
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def prefetch(bt, first=False):
        """Software pipelining across steps (geometry frozen: the trace of a batch does not depend on the parameter update before
        it): the primary trace of the NEXT batch is enqueued on a side stream before this batch's step.  Every trace still runs
        inside the timed region -- the loop below starts one per step (the last step's wraps around to the first batch) and joins
        the side stream before the closing event; the first step consumes the one the warm-up started."""
        if PREFETCH == 1 or (PREFETCH == 2 and first):
            model.prefetch_trace({'uv': bt[0], 'object_mask': bt[1], 'pose': pose, 'intrinsics': K})

    def timed_steps(batches, count=True):
        """W warm-up steps were done by the caller; EXACTLY len(batches) steps between barrier + synchronize, device-timed"""
        ray_count.zero_()
        prefetch(batches[0], first=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i, bt in enumerate(batches):
            prefetch(batches[(i + 1) % len(batches)])
            step(*bt, count_rays=count, nxt=batches[(i + 1) % len(batches)])
        model.prefetch_join()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)), reduce_ranks(ray_count.item())

    # ---- device-resident timing ------------------------------------------------------------------------------
    with ClockSampler(local, enabled=(rank == 0)) as clocks:      # started before the warm-up, read from the timed region on
        for i in range(args.warmup):
            if i > 0:
                prefetch(dev_batches[i + 1] if i + 1 < args.warmup else dev_batches[0])
            step(*dev_batches[i], count_rays=False)
        barrier()
        launches0 = int(lib.nefii_launch_count())
        captures0 = int(lib.nefii_trace_graph_captures())
        clocks.mark()
        ms, rays = timed_steps(dev_batches[args.warmup:])
    launches = int(lib.nefii_launch_count()) - launches0
    captures = int(lib.nefii_trace_graph_captures()) - captures0
    value = rays / (ms * 1e-3)

    # ---- end to end through the public API: pinned host inputs -> device, loss read back, every step -------------
    ray_count.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nxt = [t.to(dev, non_blocking=True) for t in pinned[args.warmup]]
    prefetch(nxt, first=True)
    barrier()
    e0.record()
    for i in range(args.warmup, n_steps):
        uv, obj, rgb = nxt
        nxt = [t.to(dev, non_blocking=True) for t in pinned[i + 1 if i + 1 < n_steps else args.warmup]]     # H2D of the next batch
        prefetch(nxt)
        loss = step(uv, obj, rgb, nxt=nxt)
        _ = loss.item()
    model.prefetch_join()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    rays_e2e = reduce_ranks(ray_count.item())
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])

    # ---- roofline of the dominant kernel (tcgen05 layer GEMM): per-launch CUDA events over one more step ---------
    # The per-launch events cost host time (two event creations and a pinned allocation per launch: the profiled step runs ~15 %
    # longer than a plain one) and keep consecutive layers from overlapping under programmatic dependent launch, so the share is
    # taken against the mean of three PLAIN steps of the same batch; with the SM clock moving by several per cent under the power
    # cap it can come out a little above 1 (seen: 1.01, 1.03) -- the ncu launch list puts it at 0.95 (profiles/r2_launch_summary.md).
    lib.nefii_gemm_profile_enable(1)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step(*dev_batches[-1], count_rays=False)
    p1.record()
    torch.cuda.synchronize()
    out3 = (ctypes.c_double * 3)()
    lib.nefii_gemm_profile_fetch(out3)
    lib.nefii_gemm_profile_enable(0)
    gemm_ms, gemm_flops, gemm_launches = out3[0], out3[1], int(out3[2])
    step_ms_profiled = p0.elapsed_time(p1)
    step_ms = ev_ms(lambda: step(*dev_batches[-1], count_rays=False), 3, warm=1)

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, model, step, flat, opt_idr, opt_sg, dev_batches, pose, K, dev, rank, world, barrier, max_over_ranks,
                            reduce_ranks, timed_steps)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))     # kernel timed inside a long step -> sustained figure
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    rows_avg = (gemm_flops / max(gemm_launches, 1)) / (2.0 * 512 * 512)
    roofline = {
        "kernel": "gemm_split_bf16_kernel (tcgen05 kind::f16, 3 MMAs per fp32 product)",
        "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf else None,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
        "tensor_issue_frac": 3 * achieved_tf / peak_tf if peak_tf else None,
        "flops_per_launch_avg": gemm_flops / max(gemm_launches, 1), "launches_per_step": gemm_launches,
        "kernel_share_of_step": gemm_ms / step_ms if step_ms > 0 else None,
        "share_note": "sum of the per-launch event times %.2f ms / mean of 3 plain steps of the same batch %.2f ms; the profiled step "
                      "itself took %.2f ms (host cost of the per-launch events)" % (gemm_ms, step_ms, step_ms_profiled),
        # NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of a
        # hidden-layer launch (profiles/), scaled from its rows to this run's average launch so that it is per launch like `achieved`
        "traffic": NCU_DRAM_BYTES_PER_ROW * rows_avg,
        "traffic_source": "profiles/r2_gemm_ncu.md (ncu --set full capture of one 131072-row hidden-layer launch: %.0f B/row, "
                          "algorithmic 4096 B/row), scaled to this run's average launch -- not measured in this run" % NCU_DRAM_BYTES_PER_ROW,
    }

    cpu, cpu1 = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_throughput(px=1024, rays=8, steps=1, warmup=0)     # ~16k rays: 10-30 s of host time
        cpu1 = cpu_port_throughput(px=128, rays=8, steps=1, warmup=0, threads=1)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (MLPs: 3 MMAs per fp32 product on 16-bit hi/lo planes, fp32 accumulate: fp16 split for the SDF network, bf16 split for the trainable stacks)", "data": "synthetic",
            "config": {"workload": "step-2 training iteration (BASELINE configs[2]): ONE global batch of num_pixels=2048 x num_rays=64 = "
                                   "131072 primary rays + 3 secondary rays per hit, its 512 2x2 patches dealt to the ranks (patch p -> "
                                   "rank p %% N), 128 SGs, 8x512 SDF MLP frozen, fwd+loss+bwd+Adam",
                       "rays_per_step": rays / args.steps, "primary_rays_per_step": NUM_PIXELS * NUM_RAYS,
                       "pipelining": ("the primary trace of batch i+1 runs on a side stream next to the step of batch i (frozen geometry: "
                                      "identical results); every trace is inside the timed region, the side stream is joined before the "
                                      "closing event") if PREFETCH else "none (steps strictly one after the other)",
                       "scene": "geometric-init sphere perturbed into a bumpy blob (bumps %.2f), camera at distance 3, ~50 %% hits" % SCENE_BUMPS,
                       "l2": "per-step working set (MLP activations, several GB) >> 126 MB L2; a different pixel batch every step",
                       "parallelism": "dp%d (patches dealt to the ranks, one flat NCCL all-reduce of 3.2M grads)" % world},
            "e2e": {"value": rays_e2e / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "trace_graph_captures_in_timed_region": captures,
            "gpu_launches_note": "kernels of libnefii_b200.so; a replay of a captured trace graph counts one trip per device-driven loop",
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "cpu_baseline_1thread": cpu1,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_extras(args, model, step, flat, opt_idr, opt_sg, dev_batches, pose, K, dev, rank, world, barrier, max_over_ranks, reduce_ranks,
               timed_steps):
    """Everything beside the headline: each item is wrapped so that a failure never takes the headline line down."""
    from nefii_b200 import _lib
    lib = _lib.raw()
    ex = {}

    def guarded(name, fn):
        try:
            ex[name] = fn()
        except Exception as exc:
            print("extra %s failed: %r" % (name, exc), file=sys.stderr)
            ex[name] = None

    # ---- weak scaling (N > 1): every rank its own full 2048-pixel batch, as round 1 measured -----------------------------------
    def weak():
        if world == 1:
            return None
        k = min(args.steps, 5)
        batches = [[t.to(dev) for t in make_batch(5000 + 100 * rank + i)] for i in range(k + 1)]
        step(*batches[0], count_rays=False)
        ms, rays = timed_steps(batches[1:])
        return {"value": rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / k, "steps": k,
                "note": "every rank renders its own 2048-pixel batch (131072 primary rays per GPU)"}
    guarded("weak_scaling", weak)

    # ---- the secondary-training pass the reference's trainer runs every `secondary_train_interval` = 10 steps
    # (idr_train.py:804-852: <= secondary_batch_size 1024 secondary hit points (/ world) x num_rays 64 directions through
    # forward(with_point=True), L1 between the SG and the radiance-field colour, backward, both optimizers) ----
    def secondary():
        with torch.no_grad():
            out = model({'uv': dev_batches[-1][0], 'object_mask': dev_batches[-1][1], 'pose': pose, 'intrinsics': K})
        sp, sm, sd = out['secondary_points'].reshape(-1, 3), out['secondary_mask'].reshape(-1), out['secondary_dir'].reshape(-1, 3)
        idx = torch.nonzero(sm).squeeze(1)[:max(1, 1024 // world)]
        pts, dirs = sp[idx], sd[idx]
        n_sec = pts.shape[0]
        if n_sec == 0:
            return None
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
        secondary_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        secondary_step()
        s1.record()
        barrier()
        return max_over_ranks(s0.elapsed_time(s1))
    guarded("ms_per_secondary_training_pass", secondary)

    # ---- BASELINE metric tail "ms/frame 800x800" (configs[3]): novel-view render, eval mode, 1 ray per pixel, pixels dealt to the
    # ranks, the 11 output planes gathered onto rank 0 (scripts/render.py:283-360) ----
    from nefii_b200.utils import general

    def frame(img_w, img_h, n_rays, subsample=1):
        ii, jj = torch.meshgrid(torch.arange(img_w).float(), torch.arange(img_h).float(), indexing="xy")
        uv = torch.stack([ii, jj], -1).reshape(1, -1, 2) + 0.5
        if subsample > 1:
            uv = uv[:, ::subsample]
        n_px = uv.shape[1]
        if n_rays > 1:
            g = torch.Generator().manual_seed(7)
            uv = uv.unsqueeze(2) + (torch.rand(n_rays, 2, generator=g) - 0.5).reshape(1, 1, n_rays, 2)
        uv_host = uv.contiguous().pin_memory()
        obj_all = torch.ones(1, n_px, dtype=torch.bool, device=dev)
        Kf = K.clone()
        Kf[0, 0, 0] = Kf[0, 1, 1] = 2.4 * img_h
        Kf[0, 0, 2], Kf[0, 1, 2] = img_w / 2, img_h / 2

        def render():
            uv_all = uv_host.to(dev, non_blocking=True)
            return general.render_frame(model, {'uv': uv_all, 'object_mask': obj_all, 'pose': pose, 'intrinsics': Kf}, n_px,
                                        num_rays=max(n_rays, 1), memory_capacity_level=18)
        render()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        out = render()
        f1.record()
        barrier()
        del out
        return max_over_ranks(f0.elapsed_time(f1)), n_px

    def frame_800():
        model.eval()
        try:
            return frame(IMG, IMG, 1)[0]
        finally:
            model.train()
    guarded("ms_per_frame_800x800", frame_800)

    # ---- BASELINE configs[4]: 1920x1080 with 256 rays per pixel on 8 GPUs: a bounded sample (every 64th pixel), stated as such,
    # extrapolated linearly in pixels; training-mode variant = + the measured 12.95 MB gradient all-reduce ----
    def cfg4():
        if world < 8 and not os.environ.get("NEFII_BENCH_CFG4"):
            return None
        sub = int(os.environ.get("NEFII_BENCH_CFG4_SUBSAMPLE", "64"))
        model.eval()
        try:
            ms, n_px = frame(1920, 1080, 256, subsample=sub)
        finally:
            model.train()
        ar_ms = ev_ms(lambda: flat.all_reduce(world), 5, warm=2)
        return {"sample": "every %dth pixel of 1920x1080 (%d pixels x 256 rays = %d primary rays over %d GPUs)" % (sub, n_px, n_px * 256, world),
                "ms_sample": ms, "ms_per_frame_extrapolated": ms * (1920 * 1080) / n_px, "primary_rays_per_s": n_px * 256 / (ms * 1e-3),
                "grad_allreduce_ms": ar_ms, "note": "eval-mode render; a training-mode pass adds one flat all-reduce of 12.95 MB per step"}
    guarded("cfg4_1080p_256spp", cfg4)

    if rank == 0 and world == 1:
        guarded("configs", lambda: standalone_configs(model, dev, pose, K, args))
        guarded("gpu_eager_baseline", lambda: gpu_eager_baseline(dev))
    if world == 1:
        guarded("fitted_scene", lambda: fitted_scene(dev, step, dev_batches))
    return ex


def fp32_peak_tflops(dev):
    """micro-benchmark of the FP32 FMA pipe (csrc/probe.cu), best of 5"""
    from nefii_b200 import _lib
    lib = _lib.raw()
    sink = torch.zeros(4, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, iters = sms * 16, 1 << 15
    best = 0.0
    for _ in range(5):
        ms = ev_ms(lambda: _lib.check(lib.nefii_probe_fp32(_lib.stream_ptr(dev), blocks, iters, sink.data_ptr())), 1, warm=1)
        best = max(best, blocks * 256.0 * iters * 16 / (ms * 1e-3) / 1e12)
    return best


def standalone_configs(model, dev, pose, K, args):
    """BASELINE configs[0] (render_with_sg forward, 1024 rays x 128 SGs) and configs[1] (sphere tracing + SDF MLP on 4096 rays)
    on their own: GPU time through the public module API, CPU port time, and for the SG kernel the fraction of both roofs."""
    from nefii_b200.model.sg_render import render_with_sg
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6541.1))
    fp32 = fp32_peak_tflops(dev)
    res = {"fp32_peak_tflops_measured": fp32, "hbm_peak_gbs": hbm}

    def sg_inputs(n, device):
        g = torch.Generator().manual_seed(0)
        normal = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        view = torch.nn.functional.normalize(normal + 0.5 * torch.randn(n, 3, generator=g), dim=-1)
        albedo = torch.rand(n, 3, generator=g)
        lgt = model.envmap_material_network.lgtSGs.detach().cpu()
        return [t.to(device) for t in (lgt, torch.full((1, 3), 0.04), torch.full((1, 1), 0.3), albedo, normal, view)]

    cfg0 = {}
    for n in (1024, 1 << 20):
        a = sg_inputs(n, dev)
        with torch.no_grad():
            ms = ev_ms(lambda: render_with_sg(*a), 50 if n == 1024 else 10, warm=3)
        rps = n / (ms * 1e-3)
        cfg0["gpu_%d_rays" % n] = {"ms": ms, "rays_per_s": rps, "fp32_fraction": rps * SG_FLOPS_PER_RAY / (fp32 * 1e12),
                                   "hbm_fraction": rps * SG_BYTES_PER_RAY / (hbm * 1e9)}
    # forward + backward (gradients to the light SGs, roughness, specular reflectance, albedo and the normals)
    a = sg_inputs(1 << 20, dev)
    leaves = [t.clone().requires_grad_(True) for t in a[:5]]
    gy = torch.rand(1 << 20, 3, device=dev)

    def fwd_bwd():
        for t in leaves:
            t.grad = None
        (render_with_sg(*leaves, a[5])["sg_rgb"] * gy).sum().backward()
    ms_fb = ev_ms(fwd_bwd, 5, warm=2)
    cfg0["gpu_1048576_rays_fwd_bwd"] = {"ms": ms_fb, "ms_backward_only": ms_fb - cfg0["gpu_1048576_rays"]["ms"],
                                        "note": "hand-derived adjoint (csrc/sg_adjoint_math.cuh); round 1: forward-mode duals, 34-44 ms"}
    cfg0["note"] = "render_with_sg forward, 128 SGs, K = 1: algorithmic 35 kFLOP + 7.5 k exp/div/sqrt and 72 B per ray -> bound by the " \
                   "FP32/MUFU pipes, not HBM (SURVEY 8d); 1024 rays is one launch of ~10 us (latency), 2^20 rays shows the throughput"
    if not args.no_cpu_baseline:
        from oracle import sg as osg
        a = sg_inputs(1024, "cpu")
        for threads in (os.cpu_count() or 1, 1):
            torch.set_num_threads(threads)
            osg.render_with_sg(*a)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                osg.render_with_sg(*a)
            dt = (time.perf_counter() - t0) / reps
            cfg0["cpu_port_%s" % ("1thread" if threads == 1 else "all_threads")] = {"ms": dt * 1e3, "rays_per_s": 1024 / dt, "cores": threads}
        torch.set_num_threads(os.cpu_count() or 1)
    res["cfg0_render_with_sg_1024x128"] = cfg0

    # configs[1]: 64 x 64 jittered pixels, eval-mode trace through RayTracing.forward (sphere tracing, sampler, bisection)
    from nefii_b200.utils import rend_util
    g = torch.Generator().manual_seed(1)
    n_side = 64
    ii, jj = torch.meshgrid(torch.arange(n_side).float(), torch.arange(n_side).float(), indexing="xy")
    uv = (torch.stack([ii, jj], -1).reshape(1, -1, 2) + torch.rand(1, n_side * n_side, 2, generator=g)) * (IMG / n_side)
    dirs, cam = rend_util.get_camera_params(uv.to(dev), pose, K)
    obj = torch.ones(n_side * n_side, dtype=torch.bool, device=dev)
    tracer = model.ray_tracer
    was_training = tracer.training
    tracer.eval()
    tracer.collect_stats = True
    with torch.no_grad():
        tracer(sdf=model.implicit_network, cam_loc=cam, object_mask=obj, ray_directions=dirs)
        stats = dict(tracer.last_stats)
        tracer.collect_stats = False
        ms = ev_ms(lambda: tracer(sdf=model.implicit_network, cam_loc=cam, object_mask=obj, ray_directions=dirs), 20, warm=2)
    tracer.train(was_training)
    cfg1 = {"gpu": {"ms": ms, "rays_per_s": 4096 / (ms * 1e-3), "sdf_evals_per_ray": stats["n_evals"] / 4096.0, "march_rounds": stats["n_rounds"],
                    "sampler_rays": stats["n_sampler"], "tflops_algorithmic": stats["n_evals"] * SDF_FLOPS_PER_POINT / (ms * 1e-3) / 1e12},
            "note": "latency-bound: 4096 rays are 64 row tiles; the march is max-over-rays(evaluations) dependent rounds of 8 layer GEMMs"}
    if not args.no_cpu_baseline:
        from oracle import tracer as otr
        om = oracle_from_net(model)
        torch.set_num_threads(os.cpu_count() or 1)
        d_c, c_c, o_c = dirs.cpu(), cam.cpu(), obj.cpu()
        t0 = time.perf_counter()
        otr.ray_trace(om.sdf_fn, c_c, o_c, d_c, om.trace_cfg, training=False)
        dt = time.perf_counter() - t0
        cfg1["cpu_port_all_threads"] = {"ms": dt * 1e3, "rays_per_s": 4096 / dt, "cores": os.cpu_count() or 1}
    res["cfg1_trace_4096_rays"] = cfg1
    return res


def gpu_eager_baseline(dev):
    """"the comparison baseline on the GPU box is the reference's own eager-PyTorch path run on the same
    B200".  The reference cannot travel to the box; its restatement (oracle/pipeline.py, plain torch ops, cuBLAS fp32 matmuls with
    TF32 off) runs here on a bounded sample of the same workload: 512 pixels x 64 rays, forward + loss + backward."""
    from oracle import pipeline
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    net = build_model("cpu")
    om = oracle_from_net(net).to(dev)
    om.lgtSGs.requires_grad_(True)
    om.material.requires_grad_(True)
    om.radiance.requires_grad_(True)
    pose, K = [t.to(dev) for t in make_camera()]
    px, rays = 512, NUM_RAYS
    total_rays, total_t = 0, 0.0
    for i in range(2):
        uv, obj, rgb = [t.to(dev) for t in make_batch(2000 + i, num_pixels=px, num_rays=rays)]
        g = torch.Generator().manual_seed(i)
        U = torch.rand(px * rays, 7, generator=g).to(dev)
        vecs = [torch.rand(100, generator=g) for _ in range(2)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pipeline.forward_with_uv(om, uv, pose, K, obj, lambda n: U[:n], True, vecs[0], vecs[1])
        loss = idr_loss_cpu(out, rgb)
        loss.backward()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= 1:
            n_hit = out['secondary_mask'].shape[1] if out['secondary_mask'] is not None else 0
            total_rays += px * rays + 3 * n_hit
            total_t += dt
    return {"value": total_rays / total_t, "unit": "rays/s", "ms_per_step_sample": total_t * 1e3,
            "sample": "%d px x %d rays, 1 step after 1 warm-up, fwd+loss+bwd, torch eager on this GPU, fp32 matmuls (TF32 off), "
                      "oracle/pipeline.py restating the reference" % (px, rays)}


ROBOT_PRIMS = [[1, 0.00, 0.05, 0.00, 0.22, 0.28, 0.14, 0], [0, 0.00, 0.48, 0.00, 0.16, 0, 0, 0],
               [1, -0.34, 0.10, 0.00, 0.07, 0.24, 0.07, 0], [1, 0.34, 0.10, 0.00, 0.07, 0.24, 0.07, 0],
               [1, -0.12, -0.48, 0.00, 0.08, 0.24, 0.08, 0], [1, 0.12, -0.48, 0.00, 0.08, 0.24, 0.08, 0],
               [0, -0.34, -0.20, 0.00, 0.09, 0, 0, 0], [0, 0.34, -0.20, 0.00, 0.09, 0, 0, 0],
               [0, 0.00, 0.05, 0.17, 0.08, 0, 0, 0]]


def fit_geometry(dev, n_fit=None):
    """conf.conf's model with its SDF network L1-fitted to the analytic "robot-scale" primitive union (SURVEY 8d cfg 2(iii)) with the
    step-1 trainer path (geometry_train.py:354-378: Adam, L1(sdf(points), gt), batch 16384) on the trainable tcgen05 stack.
    -> (model with an un-frozen geometry, seconds, last L1)"""
    from nefii_b200.model.ray_tracing import AnalyticSDF
    fit = build_model(dev, bumps=0.0)
    fit.unfreeze_geometry()
    target = AnalyticSDF(torch.tensor(ROBOT_PRIMS, dtype=torch.float32), dev)
    opt = torch.optim.Adam(fit.implicit_network.parameters(), lr=5e-4)
    g = torch.Generator(device=dev).manual_seed(3)
    n_fit = env_int("NEFII_BENCH_FIT_STEPS", 400) if n_fit is None else n_fit
    t0 = time.perf_counter()
    last = None
    for i in range(n_fit):
        x = torch.rand(16384, 3, device=dev, generator=g) * 2 - 1
        with torch.no_grad():
            gt = target(x).unsqueeze(-1)
        opt.zero_grad(set_to_none=True)
        pred = fit.implicit_network(x)[:, 0:1]
        last = (pred - gt).abs().mean()
        last.backward()
        opt.step()
    torch.cuda.synchronize()
    return fit, time.perf_counter() - t0, float(last.item())


def fitted_scene(dev, step, dev_batches):
    """The same step on the fitted geometry of fit_geometry()."""
    pose, K = [t.to(dev) for t in make_camera()]
    n_fit = env_int("NEFII_BENCH_FIT_STEPS", 400)
    fit, fit_s, last_l1 = fit_geometry(dev, n_fit)
    fit.freeze_geometry()
    fit.train()
    # the step function's optimizers and flat gradient buffer belong to the headline model: time forward + loss + backward here
    batches = dev_batches[-5:]

    def fwd_bwd(uv, obj, rgb):
        for p in fit.parameters():
            p.grad = None
        out = fit({'uv': uv, 'object_mask': obj, 'pose': pose, 'intrinsics': K})
        idr_loss(out, rgb).backward()
        return out
    warm = max(1, len(batches) - 3)        # the new network's trace graphs are captured during these calls
    for bt in batches[:warm]:
        out = fwd_bwd(*bt)
    hit_frac = out['network_object_mask'].float().mean().item()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rays = 0
    batches = batches[warm - 1:]           # batches[1:] below are the timed ones
    for bt in batches[1:]:
        o = fwd_bwd(*bt)
        rays += bt[0].shape[1] * bt[0].shape[2] + 3 * (o['secondary_mask'].shape[1] if o['secondary_mask'] is not None else 0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / len(batches[1:])
    return {"fit_steps": n_fit, "fit_seconds": fit_s, "fit_l1": last_l1, "pixel_hit_fraction": hit_frac,
            "ms_per_step_fwd_loss_bwd": ms, "value": rays / (a.elapsed_time(b) * 1e-3), "unit": "rays/s",
            "note": "SDF MLP fitted to the analytic robot scene with the step-1 path (batch 16384, Adam 5e-4), then frozen; "
                    "same batches as the headline, forward + loss + backward (no optimizer step)"}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------------------
NCU_DRAM_BYTES_PER_ROW = 3831.0      # profiles/r2_gemm_ncu.md (502.1 MB per 131 072-row launch)


def idr_loss_cpu(out, rgb_gt):
    """The same loss for the baseline arms: the oracle's restatement of IDRLoss (plain torch ops, like the reference)."""
    from oracle import loss as oloss
    terms = oloss.idr_loss_terms(out['idr_rgb_values'], out['sg_rgb_values'], rgb_gt, out['normal_values'], out['sdf_output'],
                                 out['network_object_mask'], out['object_mask'], alpha=LOSS_CONF['alpha'], r_patch=1,
                                 loss_type='L1', env_loss_type='L2')
    return oloss.idr_loss(terms)


def oracle_from_net(net):
    """The oracle's plain-tensor model holding exactly the weights of an IDRNetwork (the bench scene), on the CPU."""
    from oracle import mlp as omlp, pipeline
    from nefii_b200.model.implicit_differentiable_renderer import _effective_weight
    with torch.no_grad():
        imp = net.implicit_network
        sdf = omlp.SdfParams([_effective_weight(l).detach().cpu().clone() for l in imp._layers()],
                             [l.bias.detach().cpu().clone() for l in imp._layers()], n_freqs=imp.multires,
                             skip_layer=imp.skip_in[0] if len(imp.skip_in) else 0)
        rn = net.rendering_network
        rl = [getattr(rn, "lin%d" % l) for l in range(rn.num_layers - 1)]
        radiance = omlp.DenseParams([_effective_weight(l).detach().cpu().clone() for l in rl], [l.bias.detach().cpu().clone() for l in rl])
        ml = [l for l in net.envmap_material_network.diffuse_albedo_layers if hasattr(l, "weight")]
        material = omlp.DenseParams([l.weight.detach().cpu().clone() for l in ml], [l.bias.detach().cpu().clone() for l in ml])
        lgt = net.envmap_material_network.lgtSGs.detach().cpu().clone()
    return pipeline.OracleModel(sdf, radiance, material, lgt)


def cpu_port_throughput(px, rays, steps, warmup, threads=None):
    """Times oracle/pipeline.py (the restated reference, plain torch CPU ops) on a bounded sample of the same
    workload and scene: `px` pixels x `rays` rays, forward + loss + backward."""
    from oracle import pipeline
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    om = oracle_from_net(build_model("cpu"))
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
    torch.set_num_threads(os.cpu_count() or 1)
    return {"value": total_rays / total_t, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d px x %d rays per step, %d step(s), fwd+loss+bwd, torch CPU %d thread(s) "
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
        "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "step-2 training iteration (BASELINE configs[2]) on the host CPU, same scene and camera, bounded sample: " + cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
