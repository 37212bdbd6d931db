"""Full-size runs of BASELINE.json configs 1-3 on one B200 (config 0 is tests/golden/config0.npz,
config 3/4 multi-GPU go through bench.py --gpus N).  Device-resident timing with CUDA events, the
per-kernel-family breakdown from the C ABI's event profiler, and size-independent parity checks:
the default (`fast`) device path must equal the `generic` device path (full lower-envelope EDT per
radius, different kernels) voxel for voxel, and the 512^3 EDT must equal the CPU oracle.

    python scripts/configs_bench.py [--size 1024] [--edt-size 512] [--tag r1d] [--skip-generic]
Writes gpurun_out/configs_<tag>.json (one JSON object) and prints it."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import porespy_b200 as psb
from porespy_b200 import _device as dev
from porespy_b200 import _host, _lib


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        r = fn()
        del r
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
        del r
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def profile_once(ctx, fn):
    ctx.set_profile(True)
    ctx.profile_read()
    r = fn()
    prof = ctx.profile_read()
    ctx.set_profile(False)
    return r, {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in sorted(prof.items())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--edt-size", type=int, default=512)
    ap.add_argument("--tag", default="r1d")
    ap.add_argument("--skip-generic", action="store_true")
    ap.add_argument("--only-edt", action="store_true", help="stop after the EDT timings (configs 1 and edt_full)")
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    ctx = _lib.context(0)
    peak, peak_src = bench.peaks()
    out = {"peak_gbs": peak, "peak_source": peak_src}

    # ---- config 1: standalone exact EDT, 512^3 uint8 blobs, bit-exact squared distances
    S = args.edt_size
    from oracle import cpu as oc                        # checker only
    im_h = oc.blobs([S] * 3, porosity=0.6, blobiness=2, seed=0)
    im = torch.from_numpy(im_h.view(np.uint8)).to(device)
    n = im.numel()
    t_sq = timed(lambda: dev.edt_run(ctx, im, im.shape)[0], reps=10, warm=3)
    t_f32 = timed(lambda: dev.edt_run(ctx, im, im.shape, as_f32=True)[0], reps=10, warm=3)
    _, prof = profile_once(ctx, lambda: dev.edt_run(ctx, im, im.shape, as_f32=True)[0])
    d2 = dev.edt_run(ctx, im, im.shape)[0].cpu().numpy().view(np.uint32).reshape(im_h.shape)
    t0 = time.perf_counter()
    want = oc.edt_sq(im_h)
    t_cpu = time.perf_counter() - t0
    f32 = dev.edt_run(ctx, im, im.shape, as_f32=True)[0].cpu().numpy().reshape(im_h.shape)
    out["config1_edt"] = {
        "shape": [S] * 3, "ms_d2": t_sq, "ms_f32": t_f32, "voxels_per_s": n / (t_f32 * 1e-3),
        "model_gbs": 21 * n / (t_f32 * 1e-3) / 1e9, "model_frac": 21 * n / (t_f32 * 1e-3) / 1e9 / peak,
        "compulsory_gbs": 5 * n / (t_f32 * 1e-3) / 1e9, "kernels": prof,
        "d2_bit_exact_vs_oracle": bool(np.array_equal(d2, want)),
        "f32_bit_exact_vs_numpy_sqrt": bool(np.array_equal(f32, np.sqrt(want.astype(np.float32)))),
        "cpu_oracle_s": t_cpu, "cpu_threads": oc.num_threads(),
    }
    del im, d2, want, f32
    print(json.dumps(out["config1_edt"]), flush=True)

    # ---- the 1024^3 volume of configs 2 and 3 (and the same EDT at full size)
    S = args.size
    im = psb.generators.blobs([S] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox", as_numpy=False)
    torch.cuda.empty_cache()
    n = im.numel()
    t_edt = timed(lambda: dev.edt_run(ctx, im, im.shape, as_f32=True)[0], reps=5, warm=2)
    out["edt_full"] = {"shape": [S] * 3, "ms_f32": t_edt, "voxels_per_s": n / (t_edt * 1e-3),
                       "model_gbs": 21 * n / (t_edt * 1e-3) / 1e9, "model_frac": 21 * n / (t_edt * 1e-3) / 1e9 / peak}
    print(json.dumps(out["edt_full"]), flush=True)
    _, prof_full = profile_once(ctx, lambda: dev.edt_run(ctx, im, im.shape, as_f32=True)[0])
    out["edt_full"]["kernels"] = prof_full
    if args.only_edt:
        print(json.dumps(prof_full), flush=True)
        return

    def check_vs_generic(fn, what):
        if args.skip_generic:
            return None
        fast = fn()
        ctx.set_algo(_lib.ALGO_GENERIC)
        try:
            t0 = time.perf_counter()
            gen = fn()
            torch.cuda.synchronize()
            tg = time.perf_counter() - t0
        finally:
            ctx.set_algo(_lib.ALGO_FAST)
        same = bool(torch.equal(fast, gen))
        info = {"fast_equals_generic": same, "generic_s": tg, "nonzero_voxels": int((fast != 0).sum().item()),
                "distinct_values": int(torch.unique(fast).numel())}
        if not same:
            info["voxels_differ"] = int((fast != gen).sum().item())
        del fast, gen
        return info

    # ---- config 2: porosimetry, sizes=50, inlets = z-face, access-limited, mode='dt'
    inl = torch.zeros_like(im)
    inl[0] = 1
    f2 = lambda: psb.filters.porosimetry(im, sizes=50, inlets=inl, access_limited=True, mode="dt")
    t2 = timed(f2, reps=3, warm=1)
    _, prof2 = profile_once(ctx, f2)
    d2 = dev.edt_run(ctx, im, im.shape, want_max=True)
    mx = d2[1]
    T, R = _host.effective_thresholds(_host.reference_sizes(50, mx), mx)
    del d2
    n_eff = len(T)
    b2 = 21 + (22 + 6) * n_eff + 9
    out["config2_porosimetry"] = {
        "shape": [S] * 3, "sizes": 50, "inlets": "z=0 face", "ms": t2, "voxels_per_s": n / (t2 * 1e-3),
        "n_eff": n_eff, "max_d2": int(mx), "model_bytes_per_voxel": b2, "model_gbs": b2 * n / (t2 * 1e-3) / 1e9,
        "model_frac": b2 * n / (t2 * 1e-3) / 1e9 / peak, "kernels": prof2,
        "parity": check_vs_generic(f2, "config2"),
    }
    # porosimetry <= local_thickness voxel-wise (trimmed seeds are a subset of the seeds)
    mip, lt = f2(), psb.filters.local_thickness(im, sizes=50)
    out["config2_porosimetry"]["mip_le_lt"] = bool((mip <= lt).all().item())
    out["config2_porosimetry"]["invaded_fraction_of_pore"] = float((mip != 0).sum().item() / max(1, (lt != 0).sum().item()))
    del mip, lt, inl
    print(json.dumps(out["config2_porosimetry"]), flush=True)

    # ---- config 3: local_thickness with 100 linear float64 radii (fp64 compare path)
    sizes = np.linspace(1, float(np.sqrt(np.float32(mx))), 100)
    f3 = lambda: psb.filters.local_thickness(im, sizes=sizes)
    t3 = timed(f3, reps=3, warm=1)
    _, prof3 = profile_once(ctx, f3)
    T3, _ = _host.effective_thresholds(_host.reference_sizes(sizes, mx), mx)
    b3 = 21 + 22 * len(T3) + 9
    out["config3_lt100"] = {
        "shape": [S] * 3, "sizes": "np.linspace(1, max dt, 100) float64", "ms": t3, "voxels_per_s": n / (t3 * 1e-3),
        "n_eff": len(T3), "model_bytes_per_voxel": b3, "model_gbs": b3 * n / (t3 * 1e-3) / 1e9,
        "model_frac": b3 * n / (t3 * 1e-3) / 1e9 / peak, "kernels": prof3,
        "parity": check_vs_generic(f3, "config3"),
    }
    print(json.dumps(out["config3_lt100"]), flush=True)

    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/configs_{args.tag}.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
