"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list,
next to the CUDA-event shares bench.py measured live (kernels_ms_per_step of a bench JSON line)."""
import csv
import json
import re
import sys
from collections import OrderedDict

FAMILY = [("xdist_kernel<0>", "edt_x"), ("xdist_kernel<1>", "lt_x"), ("xdist_bits", "lt_x"), ("edt_col_kernel", "edt_fallback"), ("edt_minplus_kernel<MpSrcU16", "edt_y"),
          ("edt_minplus_kernel<MpSrcU32", "edt_z"), ("edt_minplus16_kernel<MpSrcU16", "edt_y"),
          ("edt_minplus16_kernel<MpSrcU32", "edt_z"), ("edt_fix_inf", "edt_fix_inf"), ("lt_classify", "lt_classify"),
          ("lt_y2", "lt_y"), ("lt_y3", "lt_y"), ("lt_zsweep", "lt_z"), ("lt_z_kernel", "lt_z"), ("lt_xy", "lt_xy"), ("lt_pack", "lt_pack"),
          ("lt_wmask", "lt_wmask"), ("lt_bitball", "lt_bitball"), ("lt_ballz", "lt_bitball"), ("lt_expand", "lt_expand"),
          ("lt_point", "lt_point"), ("uf_", "flood"), ("noise_philox", "blobs"), ("gauss_", "blobs"),
          ("stats_kernel", "blobs"), ("blobs_finish", "blobs"), ("mask_unpack", "upload"), ("mask_pack", "upload")]


def family(name):
    for key, fam in FAMILY:
        if key in name:
            return fam
    return name[:30]


def main(launch_csv, bench_json=None, steps_in_list=None):
    rows = [r for r in csv.reader(l for l in open(launch_csv) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = OrderedDict(), {}
    for r in rows[1:]:
        f = family(r[ki])
        tot[f] = tot.get(f, 0.0) + float(r[vi].replace(",", "")) / 1e6
        cnt[f] = cnt.get(f, 0) + 1
    total = sum(tot.values())
    ev = {}
    if bench_json:
        line = [l for l in open(bench_json) if l.startswith("{")][-1]
        roof = json.loads(line)["roofline"]
        ev = roof.get("kernels_ms_per_step") or {k: v["ms_per_step"] for k, v in roof["kernels"].items()}
    evtot = sum(ev.values()) or 1.0
    print(f"# {launch_csv}: {len(rows) - 1} launches, {total:.2f} ms under ncu (cold-cache, serialised)")
    print(f"{'kernel':14s} {'launches':>8s} {'ncu ms':>9s} {'ncu share':>9s} {'event ms/step':>13s} {'event share':>11s}")
    for f, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        e = ev.get(f)
        print(f"{f:14s} {cnt[f]:8d} {ms:9.3f} {ms / total:9.3f} "
              f"{(f'{e:13.3f}' if e is not None else ' ' * 13)} {(f'{e / evtot:11.3f}' if e is not None else '')}")


if __name__ == "__main__":
    main(*sys.argv[1:])
