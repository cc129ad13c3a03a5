import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception:
        print(l.strip()); continue
    print(d["B"], d["S"], d["H"], d["dtype"], "ragged" if d["ragged"] else "full", "ours %.3f ms" % d["ours_ms"], "aten %.3f ms" % d["aten_reference_ms"], "x%.1f" % d["speedup_vs_aten"], "%.0f GB/s" % d["achieved_gbs"], "frac %.2f" % d["frac_of_measured_hbm"], "err %.1e" % d["max_abs_err_vs_reference"])
