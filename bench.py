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
  roofline     : dominant kernel family: algorithmic bytes (SURVEY 8(d) model) / its CUDA-event
          duration inside the timed steps against MEASURED_PEAKS.json (`frac`), the same with the
          kernel's REAL dram bytes from the committed ncu capture (`dram_frac`), the whole path
          (`path`), and the per-kernel table (`kernels`).
  parity_check : after the timed loop one result is compared with the CPU oracle on a crop whose
          interior (margin >= max dt + r_max from every cut face) is provably the full-volume
          result; N > 1 additionally runs a small sharded volume against the oracle before warm-up
          and takes the crop across the slab face between ranks 0 and 1.
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
    # lt_bitball / lt_pack / lt_wmask work on 1 bit per voxel: the u32 pass model does not apply to them
    # (they are covered by the path figure and by their real DRAM traffic)
    "generic_x": 8, "generic_y": 8, "generic_z": 6,
}


def ncu_traffic_table(edge):
    """{kernel family: dram bytes per launch} from profiles/ncu_traffic.json ({} at another volume size)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if int(t["edge"]) != int(edge):
            return {}
        return {k: float(v["dram_bytes_per_launch"]) for k, v in t["kernels"].items()}
    except Exception:
        return {}


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
def make_input(psb, shape, blobiness, job=None):
    """The workload's volume, generated on the device by the product generator
    (porespy_b200.generators.blobs, hand-written kernels).  One GPU: numpy's seeded noise stream, i.e. the image
    IS ps.generators.blobs(shape, porosity=0.6, blobiness=2, seed=0) (generators/_imgen.py:1023-1051).
    Sharded: every rank generates its slab of ONE global image from Philox noise keyed by the global voxel
    index (the host generator cannot produce 2048^3: >= 4 float64 temporaries of 69 GB)."""
    import torch
    if job is None:
        im = psb.generators.blobs(list(shape), porosity=POROSITY, blobiness=blobiness, seed=0, as_numpy=False)
    else:
        im = job.blobs(porosity=POROSITY, blobiness=blobiness, seed=0).view(*job.local_shape)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return im


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
def host_threads():
    """Every hardware thread of the box, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    return max(1, os.cpu_count() or 1)


def time_cpu_port(edge, full_edge, steps, warmup):
    """Oracle port of the reference path (float semantics, F:1124-1212 as one C loop over the OpenMP EDT) on a
    crop-equivalent sample: `edge`^3 blobs with the feature size of the full workload, all host threads."""
    from oracle import cpu as oc
    im = sample_blobs(edge, full_edge)
    nthreads = host_threads()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oc.local_thickness_c(im, sizes=SIZES, nthreads=nthreads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return im.size / t, t, nthreads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    edge = args.ref_edge
    v, t, cores = time_cpu_port(edge, args.size, args.steps, min(args.warmup, 1))
    cfg = workload_config(args, args.gpus)
    cfg["reference_sample"] = (f"{edge}^3 crop-equivalent of the workload (BASELINE.md section 3: CPU on a 256^3 crop, "
                               f"linear-in-voxels rate); voxels/s is the rate on that sample")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{edge}^3 blobs with the workload's feature size "
                                   f"(sigma={args.size / (40.0 * BLOBINESS):.1f} vox), sizes={SIZES}, "
                                   f"oracle/cpu.py local_thickness_c (reference loop F:1124-1212 in C, OpenMP EDT), "
                                   f"{cores} host threads set explicitly"},
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
    if getattr(args, "stated_blobiness", False):
        b = float(BLOBINESS)
        feat = (f"sigma = {float(np.mean(shape)) / (40.0 * b):.1f} voxels: BASELINE config 4 as stated "
                f"(blobiness=2 at every N: a magnified copy, per-GPU work grows with N)")
    else:
        b = workload_blobiness(args.size, n)
        feat = (f"sigma = {args.size / (40.0 * BLOBINESS):.1f} voxels at every N (blobiness scaled with "
                f"mean(shape) so that per-GPU work is fixed)")
    return {"workload": f"ps.filters.local_thickness(blobs({list(shape)}, porosity={POROSITY}, "
                        f"blobiness={b:.4g}, seed=0), sizes={SIZES})",
            "feature_size": feat, "blobiness": b,
            "input": ("porespy_b200.generators.blobs on the device, numpy-seeded noise (= the reference's image for seed 0)"
                      if n == 1 else "ShardedVolume.blobs: every rank its slab of one global image (Philox noise per global voxel)"),
            "shape": list(shape), "sizes": SIZES, "sharding": "none" if n == 1 else f"z-slab x{n}",
            "l2": "inputs (>=1 B/voxel x 1e9 voxels) exceed the 126 MB L2; no flush needed"}


def global_shape(size, n):
    """Weak scaling at equal per-GPU volume: 1 -> S^3, 2 -> [2S,S,S], 4 -> [2S,2S,S], 8 -> [2S]^3."""
    s = size
    return {1: (s, s, s), 2: (2 * s, s, s), 4: (2 * s, 2 * s, s), 8: (2 * s, 2 * s, 2 * s)}[n]


# ------------------------------------------------------------------------- parity checks
CROP, MARGIN_MIN = 384, 64


def _oracle_crop(crop_in, radii, nthreads):
    from oracle import cpu as oc
    return oc.local_thickness_c(crop_in, sizes=radii, nthreads=nthreads)


def parity_check_one_gpu(im, out, max_d2_device):
    """Result of one timed-loop step against the CPU oracle.  The radii come from an independent CPU EDT of the
    whole volume (max dt, F:1131-1134); the map is compared on the corner crop [0:C)^3 minus a margin of
    max dt + r_max at the three cut faces (inside it the crop's result IS the full-volume result: a voxel's value
    depends on the image within r_max of seeds within r_max... of it, and the other three faces are true borders)."""
    from oracle import cpu as oc
    nthreads = host_threads()
    t0 = time.perf_counter()
    im_h = im.cpu().numpy().astype(bool)
    max_d2 = int(oc.edt_sq(im_h, nthreads=nthreads).max())
    dtmax = np.sqrt(np.float32(max_d2))
    radii = np.logspace(start=np.log10(dtmax), stop=0, num=SIZES)              # F:1132
    C = min(CROP, min(im_h.shape))
    margin = int(np.ceil(float(dtmax) + float(radii[0]))) + 2
    keep = C - margin if C < min(im_h.shape) else C
    info = {"oracle": "oracle/cpu.py local_thickness_c on the corner crop", "crop": C, "margin": margin,
            "compared_voxels": int(keep) ** 3, "max_d2_cpu": max_d2, "max_d2_gpu": int(max_d2_device)}
    if keep < 32:
        info.update(ok=False, why="crop smaller than the margin")
        return info
    want = _oracle_crop(im_h[:C, :C, :C], radii, nthreads)[:keep, :keep, :keep]
    got = out[:keep, :keep, :keep].cpu().numpy()
    nbad = int((got != want).sum())
    info.update(ok=bool(nbad == 0 and max_d2 == int(max_d2_device)), mismatches=nbad,
                seconds=round(time.perf_counter() - t0, 1))
    return info


def parity_check_sharded(job, im, out, max_d2, rank, world, device):
    """N > 1: the crop sits across the slab face between ranks 0 and 1 (z in [nzl0 - C/2, nzl0 + C/2), x, y in
    [0, C)); every rank contributes its part, rank 0 runs the oracle.  Radii from the sharded max d2 (a CPU EDT of
    8.6e9 voxels is out of reach; the one-GPU check covers the radii)."""
    import torch
    import torch.distributed as dist
    nz, ny, nx = job.shape
    C = min(CROP, ny, nx)
    zc = job.zstarts[1]
    z0, z1 = max(0, zc - C // 2), min(nz, zc + C // 2)
    cin = torch.zeros((z1 - z0, C, C), dtype=torch.uint8, device=device)
    cout = torch.zeros((z1 - z0, C, C), dtype=torch.float64, device=device)
    s0 = job.zstarts[rank]
    lo, hi = max(z0, s0), min(z1, s0 + job.nzl)
    if lo < hi:
        cin[lo - z0:hi - z0] = im.view(job.local_shape)[lo - s0:hi - s0, :C, :C]
        cout[lo - z0:hi - z0] = out.view(job.local_shape)[lo - s0:hi - s0, :C, :C]
    dist.all_reduce(cin)
    dist.all_reduce(cout)
    if rank != 0:
        return None
    t0 = time.perf_counter()
    dtmax = np.sqrt(np.float32(max_d2))
    radii = np.logspace(start=np.log10(dtmax), stop=0, num=SIZES)
    margin = int(np.ceil(float(dtmax) + float(radii[0]))) + 2
    zl = 0 if z0 == 0 else margin
    zh = (z1 - z0) if z1 == nz else (z1 - z0) - margin
    keep = C - margin if C < min(ny, nx) else C
    info = {"oracle": "oracle/cpu.py local_thickness_c on a crop across the slab face of ranks 0|1", "crop": [z1 - z0, C, C],
            "margin": margin, "compared_voxels": int(max(0, zh - zl)) * int(keep) ** 2, "max_d2_gpu": int(max_d2)}
    if zh - zl < 16 or keep < 32:
        info.update(ok=False, why="crop smaller than the margin")
        return info
    want = _oracle_crop(cin.cpu().numpy().astype(bool), radii, host_threads())[zl:zh, :keep, :keep]
    got = cout[zl:zh, :keep, :keep].cpu().numpy()
    nbad = int((got != want).sum())
    info.update(ok=bool(nbad == 0), mismatches=nbad, seconds=round(time.perf_counter() - t0, 1))
    return info


def sharded_small_check(ctx, rank, world):
    """A small sharded volume (EDT, local_thickness, access-limited porosimetry) against the CPU oracle on the
    whole volume, on the ranks of this very run (tests/sharded_gpu_check.py)."""
    from tests import sharded_gpu_check
    try:
        sharded_gpu_check.run_cases(ctx, rank, world, sharded_gpu_check.CASES[:1], verbose=False)
        return True
    except AssertionError as e:
        print(f"[bench] sharded small check failed on rank {rank}: {e}", file=sys.stderr, flush=True)
        return False


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
    from porespy_b200 import _device as pdev
    ctx = _lib.context(local_rank)
    shape = global_shape(args.size, world)
    cfg = workload_config(args, world)

    small_ok = None
    if world == 1:
        job = None
        im = make_input(psb, shape, cfg["blobiness"])

        def step():
            return psb.filters.local_thickness(im, sizes=SIZES)
    else:
        from porespy_b200 import sharded
        if not args.no_check:
            ok = sharded_small_check(ctx, rank, world)
            t = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            small_ok = bool(t.item())
        job = sharded.ShardedVolume(shape, ctx)
        im = make_input(psb, shape, cfg["blobiness"], job)

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
        if _ + 1 < args.steps:
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

    # ---- what was timed is checked: the last result of the timed loop against the CPU oracle
    parity = None
    if not args.no_check:
        if world == 1:
            d2, max_d2 = pdev.edt_run(ctx, im.reshape(-1), shape, want_max=True)
            del d2
            parity = parity_check_one_gpu(im, out, max_d2)
        else:
            parity = parity_check_sharded(job, im, out, job.last_max_d2, rank, world, device)
            if rank == 0 and parity is not None:
                parity["small_sharded_volume_vs_oracle"] = small_ok
                parity["ok"] = bool(parity["ok"] and small_ok)
    del out
    torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers (every rank: its own volume / slab)
    # headline: the user's volume is a numpy bool array in page-locked memory (psb.pinned_empty); also reported:
    # an ordinary pageable numpy array (what an unmodified PoreSpy script holds).  The result comes back as a
    # fresh numpy float64 array every call.
    lshape = shape if world == 1 else job.local_shape
    im_host = psb.pinned_empty(lshape, np.bool_)
    np.copyto(im_host, im.cpu().numpy().astype(bool).reshape(lshape))
    share = pdev.HOST_WIDEN_PERMILLE / 1000.0
    packed = im_host.nbytes >= pdev.UPLOAD_PACK_MIN_BYTES
    h2d = (im_host.nbytes // 8 if packed else im_host.nbytes) * world
    d2h = int(im_host.size * (share * 1 + (1.0 - share) * 8)) * world

    def time_e2e(src):
        def step_e2e():
            if world == 1:
                return psb.filters.local_thickness(src, sizes=SIZES)
            return job.local_thickness(src, sizes=SIZES, to_host=True)
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
        return te, per_call, n_e2e

    te, per_call, n_e2e = time_e2e(im_host)
    im_page = np.array(im_host, copy=True)                 # ordinary (pageable) numpy memory
    tp, per_call_p, _ = time_e2e(im_page)
    del im_page
    # the same call for consumers of the INDEX form (pore_size_distribution, size_to_satn, ... on the device:
    # porespy_b200.sizemap): numpy in, pore-size distribution out, no float64 map anywhere
    idx_form = None
    if world == 1:
        def step_psd():
            m = psb.local_thickness_index(im_host, sizes=SIZES)
            return psb.metrics.pore_size_distribution(m, bins=10)
        step_psd()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_psd()
        torch.cuda.synchronize()
        ti = (time.perf_counter() - t0) / n_e2e
        idx_form = {"value": nvox / ti, "ms_per_step": ti * 1e3,
                    "api": "porespy_b200.local_thickness_index(numpy bool) -> IndexMap -> metrics.pore_size_distribution",
                    "note": "numpy volume in, pore-size distribution out; the radius map stays on the device as 1 B/voxel"}
    api = ("porespy_b200.filters.local_thickness(numpy bool, page-locked) -> numpy float64" if world == 1 else
           "ShardedVolume.local_thickness(numpy bool slab, page-locked, to_host=True) -> numpy float64 slab, every rank")
    e2e = {"value": nvox / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": te * 1e3, "steps": n_e2e,
           "ms_per_call": [round(t * 1e3, 1) for t in per_call], "api": api,
           "pageable_input": {"value": nvox / tp, "ms_per_step": tp * 1e3,
                              "ms_per_call": [round(t * 1e3, 1) for t in per_call_p],
                              "note": "the same call with the volume in ordinary (pageable) numpy memory"},
           "index_form": idx_form,
           "input": (f"numpy bool volume of {im_host.nbytes * world} bytes in host memory, packed to bits by the library's host "
                     f"threads before the upload (psb200_upload_mask_u8)" if packed else "numpy bool volume, uploaded as bytes"),
           "result": f"float64 map of {im_host.size * 8 * world} bytes in host memory; {pdev.HOST_WIDEN_PERMILLE / 10:.0f} % of "
                     f"the volume leaves the device as 1-byte radius indices and is widened by the library's host "
                     f"threads (psb200_expand_idx_f64_to_host)"
                     + ("; the radius loop runs in z-slabs and the epilogue of a finished slab overlaps the next slab's "
                        "kernels (porespy_b200.filters.SLAB_PIPELINE)" if world == 1 else "")}
    del im_host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    per_voxels = nvox / world
    traffic = ncu_traffic_table(args.size)
    table = {}
    dram_total = 0.0
    dram_known = True
    for k, (tot_ms, cnt) in sorted(prof.items()):
        avg_s = tot_ms / cnt * 1e-3
        row = {"ms_per_step": round(tot_ms / args.steps, 4), "launches_per_step": cnt / args.steps,
               "share_of_step": round(tot_ms / ms, 4)}
        if k in ALG_BYTES:
            row["model_bytes_per_voxel"] = ALG_BYTES[k]
            row["model_frac"] = round(ALG_BYTES[k] * per_voxels / avg_s / 1e9 / peak, 4)
        if k in traffic:
            row["dram_bytes_per_launch"] = traffic[k]
            row["dram_frac"] = round(traffic[k] / avg_s / 1e9 / peak, 4)
            dram_total += traffic[k] * cnt / args.steps
        elif tot_ms / ms > 0.01:
            dram_known = False
        table[k] = row
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
        dom_traffic = traffic.get(dom) if world == 1 else None
        path = {"n_eff": n_eff, "model_bytes_per_voxel": 21 + 22 * n_eff + 9,
                "model_achieved": path_bytes / (ms_per_step * 1e-3) / 1e9 / world,
                "model_frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / world / peak,
                "note": "model = SURVEY 8(d) u32 pass model; the kernels move narrower data, so > 1 is possible"}
        if world == 1 and traffic and dram_known:
            path["dram_bytes_per_step"] = dram_total
            path["dram_achieved"] = dram_total / (ms_per_step * 1e-3) / 1e9
            path["dram_frac"] = dram_total / (ms_per_step * 1e-3) / 1e9 / peak
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": dom_traffic,
                "dram_achieved": (dom_traffic / avg_s / 1e9) if dom_traffic else None,
                "dram_frac": (dom_traffic / avg_s / 1e9 / peak) if dom_traffic else None,
                "path": path,
                "traffic_unit": "bytes per launch (ncu dram read+write, profiles/ncu_traffic.json)",
                "alg_bytes_per_launch": ALG_BYTES[dom] * per_voxels, "peak_source": peak_src,
                "alg_bytes_per_voxel_per_launch": ALG_BYTES[dom],
                "avg_launch_ms": tot_ms / cnt, "launches_per_step": cnt / args.steps,
                "share_of_step": tot_ms / ms,
                "kernels": table}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, t, cores = time_cpu_port(args.cpu_edge, args.size, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_edge}^3 blobs with the workload's feature size, sizes={SIZES}, "
                         f"oracle/cpu.py local_thickness_c (reference loop in C, OpenMP EDT); {t:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer (f64 radius map out)",
        "data": "synthetic", "config": cfg, "clocks": clocks,
        "parity_check": (parity["ok"] if parity else None), "parity": parity,
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
    ap.add_argument("--ref-edge", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle parity checks")
    ap.add_argument("--stated-blobiness", action="store_true",
                    help="blobiness=2 at every N (BASELINE config 4 as stated) instead of the fixed feature size")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
