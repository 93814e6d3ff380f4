#!/usr/bin/env python
"""After tools/final_gpu.sh <tag>: copy what is to be judged from gpurun_out/ into profiles/ and write profiles/render_traffic.json
(DRAM bytes per launch of the render kernels from the `ncu --set full` captures + the sha256 of the binary they were taken on).
    python tools/collect_profiles.py r02z"""
import csv
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    out = []
    for x in r[2:]:
        d, u = dict(zip(hdr, x)), dict(zip(hdr, units))
        b = lambda k: float(d[k]) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u[k].lower()]
        us = float(d["gpu__time_duration.sum"]) * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(
            u["gpu__time_duration.sum"].lower(), 1)
        out.append(dict(name=d["Kernel Name"], us=us, rd=b("dram__bytes_read.sum"), wr=b("dram__bytes_write.sum"), inst=float(d["smsp__inst_executed.sum"])))
    return out


def main(tag):
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    sha = hashlib.sha256(open(os.path.join(ROOT, "xworld_b200", "libxworld_b200.so"), "rb").read()).hexdigest()
    W = {}
    for key, cap, n in [("c3", "c3_render", 65536), ("c2", "c2_render", 65536), ("c4", "c4_render", 32768), ("fpv", "fpv_frame", 65536)]:
        big = max(rows(os.path.join(go, "%s_%s_raw.csv" % (tag, cap))), key=lambda r: r["us"])  # (two launches captured: the long one is the frame kernel)
        ent = dict(kernel=big["name"][:60], envs=n, dram_bytes_read=int(big["rd"]), dram_bytes_write=int(big["wr"]),
                   dram_bytes_per_launch=int(big["rd"] + big["wr"]), ncu_duration_us=big["us"], warp_instructions=int(big["inst"]),
                   raw="profiles/%s_%s_ncu_raw.csv" % (tag, cap))
        if key == "fpv":
            g = max(rows(os.path.join(go, "%s_fpv_goal_raw.csv" % tag)), key=lambda r: r["us"])
            ent["goal_kernel"] = dict(kernel=g["name"][:40], dram_bytes_read=int(g["rd"]), dram_bytes_write=int(g["wr"]), ncu_duration_us=g["us"],
                                      warp_instructions=int(g["inst"]), raw="profiles/%s_fpv_goal_ncu_raw.csv" % tag)
            ent["dram_bytes_per_launch"] = int(big["rd"] + big["wr"] + g["rd"] + g["wr"])
            ent["note"] = "frame kernel + goal kernel (bench.py's render time covers both)"
        W[key] = ent
        print(key, big["name"][:48], "%.1f us" % big["us"], ent["dram_bytes_per_launch"])
    sys.path.insert(0, ROOT)
    import bench
    rec = dict(lib_sha256=sha, src_sha256=bench.src_sha256(), capture="%s: ncu --set full --clock-control none, one launch per kernel of `bench.py --workload <w> --steps 4 --warmup 4` "
                                       "(tools/final_ncu.sh), on the binary with this sha256" % tag, workloads=W)
    json.dump(rec, open(os.path.join(pr, "render_traffic.json"), "w"), indent=1)
    for cap in ["c3_render", "c2_render", "c4_render", "fpv_frame", "fpv_goal", "c2_reset", "c3_step"]:
        shutil.copy(os.path.join(go, "%s_%s_raw.csv" % (tag, cap)), os.path.join(pr, "%s_%s_ncu_raw.csv" % (tag, cap)))
    for f in ["launches.csv", "bench.json", "bench_c2.json", "bench_c4.json", "bench_fpv.json", "bench_c5.json", "bench_reference_arm.json", "pytest_gpu.log"]:
        shutil.copy(os.path.join(go, "%s_%s" % (tag, f)), os.path.join(pr, "%s_%s" % (tag, f)))
    print("lib sha256", sha)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02z")
