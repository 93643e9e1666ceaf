"""Per-step intervals of the two recurrences from gpurun_out/timeline.csv (tools/timeline.py): start-to-start of the attention kernels."""
import csv, re, sys
rows = list(csv.DictReader(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv")))
for tag in ("attn_step_fwd", "attn_step_bwd"):
    st = [float(r["start_us"]) for r in rows if tag in r["name"]]
    iv = [round(b - a, 1) for a, b in zip(st, st[1:])]
    print(tag, "mean %.1f" % (sum(iv) / max(1, len(iv))), iv)
end = max(float(r["end_us"]) for r in rows)
main_end = max(float(r["end_us"]) for r in rows if "mha_bwd" in r["name"] or "ln_bwd" in r["name"])
print("span %.1f us; last refinement-backward kernel ends %.1f" % (end, main_end))
