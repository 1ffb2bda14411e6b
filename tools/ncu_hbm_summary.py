"""Join an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of
tools/gpu_hbm_kernels.py with that tool's manifest (gpurun_out/hbm_kernels_events.json): per kernel and launch shape the
median device duration, the DRAM bytes actually moved and the algorithmic bytes, both as GB/s against the measured peak.

    python tools/ncu_hbm_summary.py <ncu csv> <manifest of the ncu run> [<manifest of a plain run>] > profiles/r02_hbm_kernels_ncu.summary.txt
"""
import collections
import csv
import json
import re
import statistics
import sys


def load_csv(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        lid = r[col["ID"]]
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        name = re.sub(r"void |hg::|at::native::|<unnamed>::", "", name)
        d = launches.setdefault(lid, {"name": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        else:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["rd" if "read" in m else "wr"] = v * scale
    return list(launches.values())


def main(csv_path, manifest_path, plain_manifest_path=None):
    launches = load_csv(csv_path)
    man = json.load(open(manifest_path))
    # event timings are only meaningful from a run WITHOUT ncu: take them from the plain run's manifest when given
    plain = {k["label"]: k for k in json.load(open(plain_manifest_path))["kernels"]} if plain_manifest_path else {}
    peak = float(man["hbm_gbs_peak"])
    print(f"# HBM-bound kernels of the hot path, B={man['B']} (tools/gpu_hbm_kernels.py under ncu --metrics "
          f"gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none); peak = {peak:.1f} GB/s "
          "(MEASURED_PEAKS.json hbm_gbs, a device-to-device copy)")
    print("# ncu durations are cold-cache and serialised; `events` = CUDA events around back-to-back launches of the same "
          "call on rotating buffers > L2 (what the step sees)")
    print(f"# {'case':58s} {'alg MB':>8s} {'evt us':>8s} {'evt GB/s':>9s} {'frac':>6s} | {'ncu us':>8s} {'dram rd MB':>10s} "
          f"{'dram wr MB':>10s} {'dram GB/s':>9s} {'alg GB/s':>9s} {'frac':>6s}")
    pos = 0
    for k in man["kernels"]:
        rx = re.compile(k["kernel"])
        got = []
        while pos < len(launches) and len(got) < k["launches"]:
            if rx.search(launches[pos]["name"]):
                got.append(launches[pos])
            pos += 1
        if not got:
            print(f"  {k['label']:58s} (no ncu rows matched {k['kernel']})")
            continue
        # several kernels may belong to one case (e.g. softmax_stats + pckh_sweep): sum per call
        per_call = max(1, len(got) // k["launches"])
        us = statistics.median(g["us"] for g in got) * per_call
        rd = statistics.median(g.get("rd", 0.0) for g in got) * per_call
        wr = statistics.median(g.get("wr", 0.0) for g in got) * per_call
        alg = k["algorithmic_bytes"]
        ev = plain.get(k["label"], k).get("event_us", k.get("event_us_median"))
        print(f"  {k['label']:58s} {alg / 1e6:8.1f} {ev:8.2f} {alg / ev / 1e3:9.0f} {alg / ev / 1e3 / peak:6.3f} | {us:8.2f} "
              f"{rd / 1e6:10.1f} {wr / 1e6:10.1f} {(rd + wr) / us / 1e3:9.0f} {alg / us / 1e3:9.0f} {alg / us / 1e3 / peak:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
