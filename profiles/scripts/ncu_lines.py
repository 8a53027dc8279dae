"""Per-source-line summary of an ncu report's `--page source --csv --print-source cuda,sass` dump.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > dump.csv; python ncu_lines.py dump.csv [kernel substring] [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file, cur_fn, hdr = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or want not in (cur_fn or ""):
        continue
    if r[2] != "-":  # SASS rows: skip, the CUDA row above carries the per-line totals
        continue
    d = dict(zip(hdr, r))
    try:
        samples = int(d["# Samples"])
    except Exception:
        continue
    key = (cur_fn.split("(")[0][-40:], cur_file, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, {"samples": 0, "inst": 0, "long_sb": 0, "short_sb": 0, "barrier": 0, "wait": 0, "math": 0, "shared_wave": 0, "shared_ideal": 0})
    a["samples"] += samples
    a["inst"] += int(d.get("Instructions Executed", 0) or 0)
    for k, col in (("long_sb", "stall_long_sb"), ("short_sb", "stall_short_sb"), ("barrier", "stall_barrier"), ("wait", "stall_wait"), ("math", "stall_math"),
                   ("shared_wave", "L1 Wavefronts Shared"), ("shared_ideal", "L1 Wavefronts Shared Ideal")):
        try:
            a[k] += int(d.get(col, 0) or 0)
        except Exception:
            pass
tot = sum(a["samples"] for a in agg.values()) or 1
print("total samples", tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print(f"{100.0 * a['samples'] / tot:5.1f}% {a['samples']:6d} inst={a['inst']:9d} lsb={a['long_sb']:5d} ssb={a['short_sb']:5d} bar={a['barrier']:5d} wait={a['wait']:5d} "
          f"smem={a['shared_wave']}/{a['shared_ideal']} {key[1]}:{key[2]} {key[3]}")
