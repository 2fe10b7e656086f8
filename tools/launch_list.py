#!/usr/bin/env python
"""ncu launch list (CSV of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) ->
profiles/launches_rNN.md (per-kernel share of the step, DRAM bytes per launch) + profiles/traffic_rNN.json (per family).

    python tools/launch_list.py gpurun_out/launches_r02.csv r02 "<command line that was profiled>"
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [("gemm", r"gemm2?::gemm2?_kernel"), ("attention", r"attention"), ("ln_modulate", r"ln_modulate"), ("sample", r"k3_"),
          ("verify", r"k4_|top1_match"), ("vq", r"vq_"), ("embed", r"embed_next_map|first_map"),
          ("decoder_glue", r"gn_stats|gn_apply|bias_residual|upsample2x|image_to_u8"), ("conv", r"conv::conv_"),
          ("library", r"cutlass|cudnn|implicit_gemm|fmha|conv")]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("sdvar::", "")[:110]


def main():
    path, tag, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    per = collections.defaultdict(lambda: dict(ms=0.0, n=0, rd=0.0, wr=0.0, nd=0))
    seen = set()
    for r in rows:
        if r is hdr or r[ik] == "Kernel Name":
            continue
        k = short(r[ik])
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        if r[im] == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3, "s": 1e3}.get(unit, 1e-6)
            per[k]["ms"] += v * scale
            per[k]["n"] += 1
        elif r[im].startswith("dram__bytes_"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            per[k]["rd" if "read" in r[im] else "wr"] += v * scale
            if "read" in r[im]:
                per[k]["nd"] += 1
    tot = sum(p["ms"] for p in per.values())
    nl = sum(p["n"] for p in per.values())
    lines = [f"# ncu launch list of ONE timed bench step, {tag}", f"command: `{cmd}`",
             "(the same command exited 0 without ncu directly before; per-launch times are cold-cache and serialised under ncu: compare SHARES,",
             "not absolutes).  DRAM columns: dram__bytes_read.sum / dram__bytes_write.sum of the same pass, mean per launch.", "",
             f"total {tot:.1f} ms over {nl} launches", "",
             "| share | total ms | launches | DRAM read MB/launch | DRAM write MB/launch | kernel |", "|---|---|---|---|---|---|"]
    for k, p in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
        if p["ms"] / max(tot, 1e-9) < 0.0005:
            continue
        rd = f"{p['rd'] / p['nd'] / 1e6:.2f}" if p["nd"] else "-"
        wr = f"{p['wr'] / p['nd'] / 1e6:.2f}" if p["nd"] else "-"
        lines.append(f"| {100 * p['ms'] / tot:5.1f}% | {p['ms']:8.2f} | {p['n']} | {rd} | {wr} | `{k}` |")
    fam = collections.OrderedDict()
    for k, p in per.items():
        f = next((name for name, pat in FAMILY if re.search(pat, k)), "other")
        d = fam.setdefault(f, dict(launches=0, ncu_ms=0.0, rd=0.0, wr=0.0, nd=0))
        d["launches"] += p["n"]; d["ncu_ms"] += p["ms"]; d["rd"] += p["rd"]; d["wr"] += p["wr"]; d["nd"] += p["nd"]
    out = {"source": f"profiles/launches_{tag}.md (ncu, one timed step of the default bench workload, time and DRAM bytes in the same pass)", "families": {}}
    for f, d in fam.items():
        e = dict(launches=d["launches"], ncu_ms=d["ncu_ms"], ncu_share=d["ncu_ms"] / max(tot, 1e-9))
        if d["nd"]:
            e.update(dram_bytes_per_launch=(d["rd"] + d["wr"]) / d["nd"], dram_read_bytes_per_launch=d["rd"] / d["nd"],
                     dram_write_bytes_per_launch=d["wr"] / d["nd"], dram_launches_sampled=d["nd"])
        out["families"][f] = e
    lines += ["", "## families", "", "| family | launches | ncu ms | share |", "|---|---|---|---|"]
    for f, e in out["families"].items():
        lines.append(f"| {f} | {e['launches']} | {e['ncu_ms']:.2f} | {100 * e['ncu_share']:.1f}% |")
    open(os.path.join(ROOT, "profiles", f"launches_{tag}.md"), "w").write("\n".join(lines) + "\n")
    json.dump(out, open(os.path.join(ROOT, "profiles", f"traffic_{tag}.json"), "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
