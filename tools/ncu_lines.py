"""Per-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on): hottest lines of every file."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[0].strip():
        try:
            s = float(r[hdr["# Samples"]]); ex = float(r[hdr["Instructions Executed"]])
        except ValueError:
            continue
        st = {k[6:]: float(r[i] or 0) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k}
        lines.append((s, ex, cur_file, r[0], r[1].strip()[:110], st))
tot = sum(l[0] for l in lines) or 1
totex = sum(l[1] for l in lines) or 1
print(f"total samples {tot:.0f}, warp-instructions {totex:.0f}")
for s, ex, f, ln, src, st in sorted(lines, key=lambda l: -l[0])[:top]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{s / tot * 100:5.1f}% samp {ex / totex * 100:5.1f}% inst  {f}:{ln:>4s}  {tops[0][0]}/{tops[1][0]}  {src}")
