"""manual: per-source-line stall samples / executed instructions of one kernel in an .ncu-rep (ncu --import-source on)
usage: python tools/ncu_lines.py report.ncu-rep [kernel-regex] [top N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre:
    cmd += ["-k", "regex:" + kre]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = [r for r in rows if r and r[0] == "Line No"][0]
iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
stall = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
lines = [r for r in rows if len(r) >= len(hdr) - 2 and r[0].isdigit() and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in lines) or 1; totE = sum(int(r[iE]) for r in lines) or 1
print("samples", tot, "warp instructions", totE)
lines.sort(key=lambda r: -int(r[iS]))
for r in lines[:top]:
    st = sorted(((int(r[i]), h) for h, i in stall.items()), reverse=True)[:3]
    print(f"{r[0]:>4} {int(r[iS]) * 100 / tot:5.1f}% instr {int(r[iE]) * 100 / totE:5.1f}%  {' '.join(f'{h[6:]}={v * 100 // max(1, int(r[iS]))}' for v, h in st):44s} | {r[1].strip()[:110]}")
