#!/usr/bin/env python
"""bench.py -- local_thickness voxels/s on synthetic blobs volumes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl ours|reference]

One "step" = one full pass of the hot path over one volume: exact EDT -> 25 log-spaced radii ->
sphere-insertion loop -> float64 radius map (ps.filters.local_thickness(im, sizes=25)).

  value : whole-job voxels/s with the input volume already resident in HBM and the float64
          result left in HBM (CUDA events on the launching stream, max over ranks).
  e2e   : the same metric through the public API with HOST buffers
          (porespy_b200.filters.local_thickness(numpy) -> numpy), H2D and D2H inside the
          timed region.
  roofline     : dominant kernel family, algorithmic bytes (SURVEY 8(d) model) / its CUDA-event
          duration inside the timed steps, against MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle port of the reference path (oracle/cpu.py, OpenMP EDT) on a
          bounded crop-equivalent sample, timed on this box's host cores (rank 0, N=1).

`--impl reference` times only that CPU port (the reference is pure Python over an absent
native wheel, so the oracle port is the reference arm here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

POROSITY, BLOBINESS, SIZES = 0.6, 2, 25
METRIC = "local_thickness_voxels_per_s"
UNIT = "voxels/s"

# algorithmic bytes per voxel per launch of each kernel family (SURVEY.md 8(d) pass model)
ALG_BYTES = {
    "edt_x": 5, "edt_y": 8, "edt_z": 8, "edt_fh_x": 5, "edt_fh_y": 8, "edt_fh_z": 8,              # 1+4, 4+4, 4+4  (B_edt = 21)
    "lt_xy": 16, "lt_x": 8, "lt_y": 8, "lt_z": 6, "lt_point": 22,         # x(4+4) + y(4+4); z(4+1+1)  (B_rad = 22)
    "lt_expand": 9, "lt_classify": 5,
    "lt_bitball": 22,                                                     # one launch = the whole radius step
    "generic_x": 8, "generic_y": 8, "generic_z": 6,
}


def ncu_traffic(kernel, edge):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, scripts/ncu_traffic.py); None when the
    capture was taken at another volume size."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if int(t["edge"]) != int(edge):
            return None
        return float(t["kernels"][kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ inputs
def device_blobs(shape, porosity, blobiness, seed, device, sigma_shape=None):
    """Device-side look-alike of ps.generators.blobs (generators/_imgen.py:1023-1051): uniform
    noise -> separable gaussian blur (sigma = mean(shape)/(40*blobiness), reflect borders) ->
    erfc uniformisation -> `< porosity`.  float32 and a different RNG, so not bit-equal to the
    host generator -- it is only an input."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f = torch.rand(tuple(shape), generator=g, device=device, dtype=torch.float32)
    sigma = float(np.mean(sigma_shape if sigma_shape is not None else shape)) / (40.0 * blobiness)
    rad = int(4.0 * sigma + 0.5)
    x = torch.arange(-rad, rad + 1, device=device, dtype=torch.float32)
    w = torch.exp(-0.5 * (x / sigma) ** 2)
    w = w / w.sum()
    f = f[None, None]
    for ax in range(3):
        n = f.shape[2 + ax]
        r = min(rad, n - 1)
        wk = w[rad - r:rad + r + 1] / w[rad - r:rad + r + 1].sum()
        pad = [0, 0, 0, 0, 0, 0]
        pad[2 * (2 - ax)] = pad[2 * (2 - ax) + 1] = r
        kshape = [1, 1, 1, 1, 1]
        kshape[2 + ax] = 2 * r + 1
        chunks = []
        for part in torch.split(f, 64, dim=2 if ax != 0 else 3):      # bound cuDNN workspace
            chunks.append(F.conv3d(F.pad(part, pad, mode="reflect"), wk.view(kshape)))
        f = torch.cat(chunks, dim=2 if ax != 0 else 3)
        del chunks
    f = f[0, 0]
    z = (f - f.mean()) / f.std()
    del f
    c = 0.5 * torch.erfc(-z / np.sqrt(2.0))
    del z
    u = (c - c.min()) / (c.max() - c.min())
    return (u < porosity).to(torch.uint8)


def sample_blobs(edge, full_edge):
    """Host blobs with the feature size of the full workload (sigma = full_edge/(40*blobiness))."""
    from oracle import cpu as oc
    b = BLOBINESS * edge / float(full_edge)
    return oc.blobs([edge] * 3, porosity=POROSITY, blobiness=b, seed=0)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples that arrived inside the timed region [t0, t1] (the sampler itself is started long before,
        nvidia-smi needs a few hundred ms to print its first line); if the region was shorter than one
        sampling period, the samples nearest to it."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)                                   # let the last in-region samples arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        rows = self.rows
        window = "timed region"
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.06]
            if not inside and rows:
                inside = sorted(rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
                window = "nearest to the timed region"
            rows = inside
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# --------------------------------------------------------------------------- reference arm
def time_cpu_port(edge, full_edge, steps, warmup):
    """Oracle port of the reference path (float semantics, OpenMP EDT) on a crop-equivalent sample."""
    from oracle import cpu as oc
    im = sample_blobs(edge, full_edge)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oc.local_thickness(im, sizes=SIZES, mode="dt")
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return im.size / t, t, oc.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    edge = args.ref_edge
    v, t, cores = time_cpu_port(edge, args.size, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{edge}^3 blobs with the workload's feature size "
                                   f"(sigma={args.size / (40.0 * BLOBINESS):.1f} vox), sizes={SIZES}, "
                                   f"oracle/cpu.py local_thickness(mode='dt'), all host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_blobiness(size, n):
    """Weak scaling keeps the per-GPU WORK fixed, so the blobs keep the feature size of the one-GPU
    workload (sigma = S/(40*2) voxels): blobs() ties sigma to mean(shape)/(40*blobiness), hence the
    blobiness of the grown volume is scaled by mean(shape)/S (a larger field of view of the same
    material, not a magnified copy of it)."""
    return BLOBINESS * float(np.mean(global_shape(size, n))) / float(size)


def workload_config(args, n):
    shape = global_shape(args.size, n)
    return {"workload": f"ps.filters.local_thickness(blobs({list(shape)}, porosity={POROSITY}, "
                        f"blobiness={workload_blobiness(args.size, n):.4g}), sizes={SIZES})",
            "feature_size": f"sigma = {args.size / (40.0 * BLOBINESS):.1f} voxels at every N (blobiness scaled with "
                            f"mean(shape) so that per-GPU work is fixed)",
            "shape": list(shape), "sizes": SIZES, "sharding": "none" if n == 1 else f"z-slab x{n}",
            "l2": "inputs (>=1 B/voxel x 1e9 voxels) exceed the 126 MB L2; no flush needed"}


def global_shape(size, n):
    """Weak scaling at equal per-GPU volume: 1 -> S^3, 2 -> [2S,S,S], 4 -> [2S,2S,S], 8 -> [2S]^3."""
    s = size
    return {1: (s, s, s), 2: (2 * s, s, s), 4: (2 * s, 2 * s, s), 8: (2 * s, 2 * s, 2 * s)}[n]


# ------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # streams from now on; only the samples inside the timed region are used
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import porespy_b200 as psb
    from porespy_b200 import _lib
    ctx = _lib.context(local_rank)
    shape = global_shape(args.size, world)

    if world == 1:
        im = device_blobs(shape, POROSITY, BLOBINESS, seed=0, device=device)
        torch.cuda.synchronize()

        def step():
            return psb.filters.local_thickness(im, sizes=SIZES)
    else:
        from porespy_b200 import sharded
        job = sharded.ShardedVolume(shape, ctx)
        # every rank generates its own slab (independent noise per slab: it is only an input)
        im = device_blobs(job.local_shape, POROSITY, BLOBINESS, seed=rank, device=device,
                          sigma_shape=(args.size,) * 3)       # feature size of the one-GPU workload

        def step():
            return job.local_thickness(im, sizes=SIZES)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
        del out
    barrier()
    ctx.set_profile(True)
    ctx.profile_read()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step()
        del out
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    prof = ctx.profile_read()
    ctx.set_profile(False)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    nvox = float(np.prod(shape))
    value = nvox / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers (every rank: its own volume / slab)
    # the user's volume: a numpy bool array in page-locked memory (psb.pinned_empty); the result
    # comes back as a fresh numpy float64 array (page-locked too) every call
    lshape = shape if world == 1 else job.local_shape
    im_host = psb.pinned_empty(lshape, np.bool_)
    np.copyto(im_host, im.cpu().numpy().astype(bool).reshape(lshape))
    # result bytes that cross PCIe: index bytes for the share widened by host threads, float64 for the rest
    from porespy_b200 import _device as pdev
    share = pdev.HOST_WIDEN_PERMILLE / 1000.0
    # input bytes that cross PCIe: volumes from UPLOAD_PACK_MIN_BYTES on are packed to one bit per voxel by host threads
    packed = im_host.nbytes >= pdev.UPLOAD_PACK_MIN_BYTES
    h2d = (im_host.nbytes // 8 if packed else im_host.nbytes) * world
    d2h = int(im_host.size * (share * 1 + (1.0 - share) * 8)) * world

    def step_e2e():
        if world == 1:
            return psb.filters.local_thickness(im_host, sizes=SIZES)
        return job.local_thickness(im_host, sizes=SIZES, to_host=True)

    n_e2e = max(1, min(args.steps, args.e2e_steps))
    res = step_e2e()      # warm-up
    del res
    barrier()
    t0 = time.perf_counter()
    per_call = []
    for _ in range(n_e2e):
        t1 = time.perf_counter()
        res = step_e2e()
        del res
        per_call.append(time.perf_counter() - t1)
    barrier()
    te = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([te], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    api = ("porespy_b200.filters.local_thickness(numpy bool, page-locked) -> numpy float64" if world == 1 else
           "ShardedVolume.local_thickness(numpy bool slab, page-locked, to_host=True) -> numpy float64 slab, every rank")
    e2e = {"value": nvox / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": te * 1e3, "steps": n_e2e,
           "ms_per_call": [round(t * 1e3, 1) for t in per_call], "api": api,
           "input": (f"numpy bool volume of {im_host.nbytes * world} bytes in host memory, packed to bits by the library's host "
                     f"threads before the upload (psb200_upload_mask_u8)" if packed else "numpy bool volume, uploaded as bytes"),
           "result": f"float64 map of {im_host.size * 8 * world} bytes in host memory; {pdev.HOST_WIDEN_PERMILLE / 10:.0f} % of "
                     f"the volume leaves the device as 1-byte radius indices and is widened by the library's host "
                     f"threads (psb200_expand_idx_f64_to_host)"}
    del im_host

    if rank != 0:
        return
    peak, peak_src = peaks()
    per_voxels = nvox / world
    fam = {k: v for k, v in prof.items() if k in ALG_BYTES}
    dom = max(fam, key=lambda k: fam[k][0]) if fam else None
    roof = None
    if dom:
        tot_ms, cnt = fam[dom]
        avg_s = tot_ms / cnt * 1e-3
        achieved = ALG_BYTES[dom] * per_voxels / avg_s / 1e9
        # one per effective radius: byte pipeline (lt_xy | lt_y), bit pipeline (lt_bitball), T == 1 (lt_point)
        n_eff = sum(c for k, (m, c) in prof.items()
                    if k in ("lt_xy", "lt_y", "lt_point", "lt_bitball", "generic_x")) / args.steps
        path_bytes = (21 + 22 * n_eff + 9) * nvox
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": ncu_traffic(dom, args.size) if world == 1 else None,
                "traffic_unit": "bytes per launch (ncu dram read+write, profiles/ncu_traffic.json)",
                "alg_bytes_per_launch": ALG_BYTES[dom] * per_voxels, "peak_source": peak_src,
                "alg_bytes_per_voxel_per_launch": ALG_BYTES[dom],
                "avg_launch_ms": tot_ms / cnt, "launches_per_step": cnt / args.steps,
                "share_of_step": tot_ms / ms if world == 1 else None,
                "path": {"n_eff": n_eff, "alg_bytes_per_voxel": 21 + 22 * n_eff + 9,
                         "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9 / world,
                         "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / world / peak},
                "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items())}}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, t, cores = time_cpu_port(args.cpu_edge, args.size, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_edge}^3 blobs with the workload's feature size, sizes={SIZES}, "
                         f"oracle/cpu.py local_thickness(mode='dt'); {t:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer (f64 radius map out)",
        "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024, help="edge S of the per-GPU S^3 volume")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-edge", type=int, default=320)
    ap.add_argument("--ref-edge", type=int, default=192)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
