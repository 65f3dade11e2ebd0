#!/usr/bin/env python
"""profiles/r02_traffic.json from the `ncu --set full` captures of tools/gpu_profile.sh: DRAM bytes
of one encode and one decode launch at the bench configuration, keyed by a hash of the kernel
sources (bench.py reports roofline.traffic only while the hash still matches the tree).

    python tools/make_traffic_json.py gpurun_out/r02_encode_65536.ncu-rep gpurun_out/r02_decode_65536.ncu-rep 65536
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, row = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(key)
        res[key] = int(float(row[i].replace(",", "")) * scale[units[i]])
    res["gpu__time_duration_ms"] = float(row[hdr.index("gpu__time_duration.sum")].replace(",", "")) * \
        {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6}.get(units[hdr.index("gpu__time_duration.sum")], 1)
    return res


def main():
    enc, dec, streams = sys.argv[1], sys.argv[2], int(sys.argv[3])
    import bench
    e, d = dram_bytes(enc), dram_bytes(dec)
    t = {"kernel_source_hash": bench.kernel_source_hash(),
         "source": "ncu --set full --clock-control none, tools/gpu_profile.sh (encode: kernel replay; decode: application replay)",
         "config3": {
             "slzw_encode_kernel": {"streams": streams, "dram_bytes_read": e["dram__bytes_read.sum"],
                                    "dram_bytes_write": e["dram__bytes_write.sum"], "launch_ms_under_ncu": e["gpu__time_duration_ms"]},
             "slzw_decode_kernel": {"streams": streams, "dram_bytes_read": d["dram__bytes_read.sum"],
                                    "dram_bytes_write": d["dram__bytes_write.sum"], "launch_ms_under_ncu": d["gpu__time_duration_ms"]}}}
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(t, f, indent=1)
    print(json.dumps(t))


if __name__ == "__main__":
    main()
