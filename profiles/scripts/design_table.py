"""Regenerates the table of measured numbers in DESIGN.md (between the MEASURED_TABLE markers) from a stored bench
line. usage: python profiles/scripts/design_table.py profiles/bench_r02/n1.json [profiles/bench_r02/ref_n1.json]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load(p):
    return json.loads(open(p).read().strip().splitlines()[-1])


def row(name, w, ref=None):
    r = w["roofline"]
    e2e = w["e2e"]["ms_per_step"]
    cpu = w.get("cpu_baseline", {})
    refv = ref.get("value") if ref else None
    shim = w.get("shim_latency", {}).get("median_us")
    cells = [name, f"{w['ms_per_step']:.3f}", f"{w['value']:.3g}", f"{e2e:.3f}",
             f"{r['frac']:.3f} ({r['kernel'].split(' (')[0]})",
             f"{r['traffic'] / 1e6:.0f} MB" if r.get("traffic") else "-",
             f"{cpu.get('value', 0):.3g} (1 thread)" + (f" / {refv:.3g} (all threads)" if refv else ""),
             f"{shim:.0f} us" if shim else "-"]
    return "| " + " | ".join(cells) + " |"


def main():
    d = load(sys.argv[1])
    ref = load(sys.argv[2]) if len(sys.argv) > 2 else None
    refw = ref.get("workloads", {}) if ref else {}
    lines = ["| workload | device ms / step | edges linearised / s | e2e ms / step | contract-byte roofline fraction (class) | "
             "DRAM traffic of the dominant kernel per launch (ncu) | CPU oracle edges / s | through the shim |",
             "|---|---|---|---|---|---|---|---|",
             row("C2 (headline)", d, ref)]
    for k, v in d.get("workloads", {}).items():
        lines.append(row(k, v, refw.get(k)))
    text = "\n".join(lines)
    p = os.path.join(ROOT, "DESIGN.md")
    s = open(p).read()
    if "@@MEASURED_TABLE@@" in s:
        s = s.replace("@@MEASURED_TABLE@@", "<!-- MEASURED_TABLE -->\n" + text + "\n<!-- /MEASURED_TABLE -->")
    else:
        s = re.sub(r"<!-- MEASURED_TABLE -->.*?<!-- /MEASURED_TABLE -->",
                   "<!-- MEASURED_TABLE -->\n" + text.replace("\\", "\\\\") + "\n<!-- /MEASURED_TABLE -->", s, flags=re.S)
    open(p, "w").write(s)
    print(text)


if __name__ == "__main__":
    main()
