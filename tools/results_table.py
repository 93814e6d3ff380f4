#!/usr/bin/env python
"""Markdown table of an end-of-round pass (tools/final_gpu.sh): python tools/results_table.py profiles/r02z"""
import json
import sys


def load(f):
    return json.loads(open(f).read().strip().splitlines()[-1])


def main(prefix):
    d = load(prefix + "_bench.json")
    rows = []
    r, e = d["roofline"], d["e2e"]
    rows.append(("C3 = BASELINE `configs[2]`: walls.json rules, 11×11 → 84×84×3, 65,536 envs, ~271 resets per step (default line, %d steps)" % d["steps"],
                 d["value"], d["ms_per_step"], r["kernel"], r["kernel_ms"], r["achieved"], r["frac"], e["value"], e.get("pipelined", {}).get("value")))
    names = {"c2": "C2: navigation2d.json rules, 7×7 → 84×84×3, 65,536 envs", "c4": "C4 = `configs[3]` per GPU: 15×15 → 128×128×3, 32,768 envs",
             "fpv": "first-person view: walls.json rules, 11×11, `visible_radius` 7 → 84×84×3, 65,536 envs",
             "c5": "C5 = `configs[4]`: SimpleRace, 1,048,576 envs, render off"}
    for k in ("c2", "c4", "fpv", "c5"):
        try:
            w = load("%s_bench_%s.json" % (prefix, k))
        except IOError:
            continue
        r, e = w["roofline"], w["e2e"]
        rows.append((names[k] + " (`--workload %s`, %d steps)" % (k, w["steps"]), w["value"], w["ms_per_step"], r["kernel"], r["kernel_ms"], r["achieved"],
                     r["frac"], e["value"], e.get("pipelined", {}).get("value")))
    out = ["| workload | `value` env-steps/s (inputs in HBM) | ms / step | render kernel | kernel ms | GB/s | of 6551.7 GB/s | `e2e` (synchronous `xw_step_hd`) | pipelined |",
           "|---|---|---|---|---|---|---|---|---|"]
    fmt = lambda v: "—" if v is None else ("%.2f G" % (v / 1e9) if v >= 1e10 else "%.1f M" % (v / 1e6))
    for n, v, ms, kn, kms, gbs, fr, ev, pv in rows:
        out.append("| %s | **%s** | %.4f | `%s` | %.4f | %.0f | **%.1f %%** | %s | %s |" % (n, fmt(v), ms, kn, kms, gbs, 100 * fr, fmt(ev), fmt(pv)))
    cfgs = d.get("configs", {})
    if cfgs:
        out.append("")
        out.append("`configs` block of the default line (40 steps each): " + "; ".join(
            "%s %s (render %.1f %%, e2e %s)" % (k, fmt(v["value"]), 100 * v["roofline"]["frac"], fmt(v["e2e"]["value"])) for k, v in cfgs.items()) + ".")
        if cfgs.get("c5", {}).get("multi_step"):
            out.append("C5 with 32 steps per launch (`xw_step_seq`): %s." % fmt(cfgs["c5"]["multi_step"]["value"]))
    cb = d.get("cpu_baseline")
    if cb:
        out.append("")
        out.append("CPU legs on the same box (%d host threads): oracle C port, all threads %.0f env-steps/s; one thread %.0f; step + teacher only, one thread %.1f M; "
                   "real OpenCV following the reference's call sequence (one `warpAffine` per item per frame), one thread %s." % (
                       cb["cores"], cb["value"], cb["one_core"]["value"], cb["step_teacher_only"]["value"] / 1e6,
                       "%.0f" % cb["cv2_faithful"]["value"] if cb.get("cv2_faithful") else "n/a"))
    r = d["roofline"]
    out.append("")
    out.append("Default line details: %d launches in the timed region, render kernel share of a step %.3f, `roofline.traffic` %s, `k_step` + reset launch %.4f ms, "
               "clocks %s MHz (max %s), reasons %s; with the frames copied to the host as well: %s." % (
                   d["gpu_launches"], r["kernel_share_of_step"], ("%.3f GB" % (r["traffic"] / 1e9)) if r.get("traffic") else "null (capture of another build)",
                   r["step_reset_ms"], d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], d["clocks"]["reasons"], fmt(d["e2e"]["with_frames_to_host"]["value"])))
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
