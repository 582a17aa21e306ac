"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, os
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = []
fname = "?"
hdr = None
seen_kernel = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if seen_kernel is None:
            seen_kernel = r[1]
        cur_kernel = r[1]; continue
    if r[0] == "Line No":
        hdr = r; col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or cur_kernel != seen_kernel:
        continue
    if r[col["Address"]] != "-":      # per-SASS rows; keep only the per-line aggregate rows
        continue
    try:
        samp = float(r[col["# Samples"]] or 0); inst = float(r[col["Instructions Executed"]] or 0)
    except ValueError:
        continue
    stalls = {h: float(r[i] or 0) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h}
    rows.append((samp, inst, r[1].strip()[:105], fname + ":" + r[0], stalls))
print("kernel:", seen_kernel)
tot_s = sum(r[0] for r in rows) or 1; tot_i = sum(r[1] for r in rows) or 1
print("total samples %d, total warp-instr %d" % (tot_s, tot_i))
for samp, inst, src, ln, stalls in sorted(rows, key=lambda r: -r[0])[:top]:
    top3 = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% smp %5.1f%% ins  %-16s %-105s %s" % (100 * samp / tot_s, 100 * inst / tot_i, ln, src,
          " ".join("%s=%d" % (k[6:], v) for k, v in top3 if v)))
