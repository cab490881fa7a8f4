"""profiles/r02_hbm_kernels.csv + profiles/r02_hbm_traffic.json from the ncu report of tools/hbm_probe.py --once:

    ncu --profile-from-start off --set full --clock-control none -k regex:'sweep|resolve|decode_kernel|head_cand|nms_kernel|bank_predict' \\
        -o gpurun_out/hbm_kernels python tools/hbm_probe.py --once
    python tools/hbm_traffic.py gpurun_out/hbm_kernels.ncu-rep

The probe launches the kernels in a fixed order (NAMES below); bench.py reads the JSON for the `traffic` of its `hbm_kernels`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["decode", "head_candidates", "nms", "sweep_coast", "resolve_coast", "sweep_frame", "resolve_frame", "bank_predict"]
METRICS = ("dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,"
           "sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active")


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", METRICS], capture_output=True, text=True, check=True).stdout
    txt = txt[txt.index('"ID"'):]
    open(os.path.join(ROOT, "profiles", "r02_hbm_kernels.csv"), "w").write(txt)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    assert len(body) == len(NAMES), f"{len(body)} launches in the report, {len(NAMES)} expected"
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, unit_scale):
        u = units[col[name]].lower()
        return float(r[col[name]].replace(",", "")) * unit_scale[u]
    out = {"source": "profiles/r02_hbm_kernels.csv (ncu --set full --clock-control none on tools/hbm_probe.py --once, L2 flushed before each launch)", "kernels": {}}
    for name, r in zip(NAMES, body):
        rd = val(r, "dram__bytes_read.sum", {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3})
        wr = val(r, "dram__bytes_write.sum", {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3})
        us = val(r, "gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3})
        k = r[col["Kernel Name"]]
        k = k.replace("void ", "").replace("<unnamed>::", "")
        k = k[:k.index("(")] if "(" in k else k
        out["kernels"][name] = {"kernel": k, "dram_read_mb": rd, "dram_write_mb": wr, "duration_us": us, "dram_gbs": (rd + wr) / us * 1e3}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_hbm_traffic.json"), "w"), indent=1)
    for n, v in out["kernels"].items():
        print(f"{n:16s} {v['duration_us']:7.1f} us  {v['dram_read_mb'] + v['dram_write_mb']:8.1f} MB  {v['dram_gbs']:7.0f} GB/s")


if __name__ == "__main__":
    main()
