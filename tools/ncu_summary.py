"""Summarise an .ncu-rep (raw metrics + per-instruction stall hot spots) as text, for profiles/."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__icc_request_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for w in want:
        if w in d:
            print(f"{w} = {d[w]} {units[hdr.index(w)]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try: st.append((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print("stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
    print("---")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
if hi:
    hdr = rows[hi[0]]
    end = hi[1] - 1 if len(hi) > 1 else len(rows)
    body = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    def f(r, h):
        try: return float(r[ix[h]].replace(",", ""))
        except ValueError: return 0.0
    tot = sum(f(r, "# Samples") for r in body) or 1
    print(f"SASS instructions: {len(body)}, samples {tot:.0f}")
    print("top instructions by stall samples:")
    for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        reasons = {k: f(r, k) for k in ("stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_mio", "stall_math", "stall_lg", "stall_not_selected", "stall_no_inst")}
        top = max(reasons, key=reasons.get)
        print(f"  {f(r,'# Samples')/tot*100:5.2f}%  {top:18s} {r[ix['Source']][:80]}")
    print("samples by 250-instruction block:")
    for i in range(0, len(body), 250):
        s = sum(f(r, "# Samples") for r in body[i:i + 250])
        print(f"  [{i:5d}] {s/tot*100:5.1f}%")
