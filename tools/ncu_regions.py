"""Per-region dynamic profile of one kernel from an .ncu-rep (source page): executed warp-instructions,
stall samples and opcode mix for consecutive SASS address ranges.

    python tools/ncu_regions.py rep.ncu-rep [frames] [block=100]
`frames` (units the launch processed) turns totals into per-unit figures.
"""
import csv, io, re, subprocess, sys
from collections import Counter

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
block = int(sys.argv[3]) if len(sys.argv) > 3 else 100
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
hdr = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, h):
    try:
        return float(r[ix[h]].replace(",", ""))
    except ValueError:
        return 0.0


def opcode(s):
    s = re.sub(r"^\s*@!?U?P\w+\s+", "", s.strip())
    return s.split()[0].split(".")[0] if s else "?"


tot_inst = sum(num(r, "Instructions Executed") for r in body)
tot_samp = sum(num(r, "# Samples") for r in body) or 1
print(f"SASS instructions {len(body)}; executed warp-instructions {tot_inst:.0f}"
      + (f" = {tot_inst / units:.0f} per unit" if units else "") + f"; samples {tot_samp:.0f}")
mix = Counter()
for r in body:
    mix[opcode(r[ix["Source"]])] += num(r, "Instructions Executed")
print("dynamic opcode mix" + (" (per unit)" if units else "") + ":")
print("  " + "  ".join(f"{k}:{v / (units or 1):.0f}" for k, v in mix.most_common(28)))
print(f"{'idx':>6} {'inst':>12} {'per-unit':>9} {'samp%':>6} {'samp/inst':>9}  top stalls")
for i in range(0, len(body), block):
    blk = body[i:i + block]
    inst = sum(num(r, "Instructions Executed") for r in blk)
    samp = sum(num(r, "# Samples") for r in blk)
    st = Counter()
    for r in blk:
        for k in ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_mio", "stall_math", "stall_lg",
                  "stall_not_selected", "stall_no_inst", "stall_dispatch", "stall_branch_resolving", "stall_barrier"):
            st[k.replace("stall_", "")] += num(r, k)
    top = " ".join(f"{k}={v / (samp or 1) * 100:.0f}%" for k, v in st.most_common(3))
    print(f"{i:6d} {inst:12.0f} {inst / units if units else 0:9.1f} {samp / tot_samp * 100:6.1f} "
          f"{samp / (inst or 1) * 1e3:9.2f}  {top}")
