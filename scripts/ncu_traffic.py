"""Build profiles/ncu_traffic.json from `ncu --set full --page raw --csv` exports: per kernel family the
mean dram__bytes_read.sum + dram__bytes_write.sum per launch (bench.py's roofline.traffic).
    python scripts/ncu_traffic.py <edge> out.json raw1.csv raw2.csv ..."""
import csv
import json
import sys

from launch_shares import family

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(edge, out, *paths):
    acc = {}
    for p in paths:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        rd, wr = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
        for r in rows[2:]:
            fam = family(r[col["Kernel Name"]])
            b = float(r[rd].replace(",", "")) * SCALE.get(units[rd], 1) + float(r[wr].replace(",", "")) * SCALE.get(units[wr], 1)
            acc.setdefault(fam, {"bytes": [], "source": []})
            acc[fam]["bytes"].append(b)
            if p not in acc[fam]["source"]:
                acc[fam]["source"].append(p)
    res = {"edge": int(edge), "kernels": {
        f: {"dram_bytes_per_launch": sum(v["bytes"]) / len(v["bytes"]), "launches_captured": len(v["bytes"]),
            "source": v["source"]} for f, v in sorted(acc.items())}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
