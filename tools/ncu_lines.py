"""Per-source-line execution profile from an ncu report taken with --import-source on (kernels carry -lineinfo):
    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_HEADER_DIR [top]
prints the executed warp instructions and stall samples per source line of rmt_kernels.cu / rmt_model.cuh, aggregated
into the line ranges given in RANGES (edit to taste), plus the `top` hottest lines.  NVRTC names the translation unit
/root/repo/rmt_kernels.cu, so the two files are linked there for the duration of the call."""
import csv, io, os, subprocess, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, hdr = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
links = [(os.path.join(ROOT, "rmt_app_b200/csrc/rmt_kernels.cu"), os.path.join(ROOT, "rmt_kernels.cu")),
         (os.path.join(hdr, "rmt_model.cuh"), os.path.join(ROOT, "rmt_model.cuh"))]
for s, d in links:
    if os.path.lexists(d): os.remove(d)
    shutil.copyfile(s, d)
try:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True, cwd="/tmp").stdout
finally:
    for _, d in links: os.remove(d)
cur = None; head = None; per = {}
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] in ("File Path", "File Name"): cur = os.path.basename(row[1]); head = None; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": head = row; continue
    if head is None or cur is None: continue
    if row[0] == "-" or not row[0].isdigit(): continue          # SASS rows under a source line
    d = dict(zip(head, row))
    try:
        n = int(d["Instructions Executed"]); smp = int(d["# Samples"])
    except (ValueError, KeyError):
        continue
    k = (cur, int(row[0]))
    a = per.setdefault(k, [0, 0, row[1]])
    a[0] += n; a[1] += smp
tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values())
print("total warp instructions %.4g, samples %d" % (tot, tots))
for f in sorted(set(k[0] for k in per)):
    s = sum(v[0] for k, v in per.items() if k[0] == f); ss = sum(v[1] for k, v in per.items() if k[0] == f)
    print("  %-18s %6.2f %% inst  %6.2f %% samples" % (f, 100.0*s/tot, 100.0*ss/max(tots, 1)))
rng = os.environ.get("RANGES")
if rng:
    print("ranges of rmt_kernels.cu:")
    for r in rng.split(","):
        name, lo, hi = r.split(":")
        lo, hi = int(lo), int(hi)
        s = sum(v[0] for k, v in per.items() if k[0] == "rmt_kernels.cu" and lo <= k[1] <= hi)
        ss = sum(v[1] for k, v in per.items() if k[0] == "rmt_kernels.cu" and lo <= k[1] <= hi)
        print("  %-28s %5d-%5d  %6.2f %% inst  %6.2f %% samples" % (name, lo, hi, 100.0*s/tot, 100.0*ss/max(tots, 1)))
print("hottest lines:")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %-14s %5d  %5.2f %% inst %5.2f %% smp  %s" % (k[0], k[1], 100.0*v[0]/tot, 100.0*v[1]/max(tots, 1), v[2][:110]))
