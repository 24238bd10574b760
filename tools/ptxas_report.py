"""Offline resource report (registers, spills, shared memory) of the kernels of one model / launch shape:
    python tools/ptxas_report.py [n1|n2] [block] [lanes] [-DNAME=VALUE ...]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from rmt_app_b200 import engine, build

which = sys.argv[1] if len(sys.argv) > 1 else "n1"
block = int(sys.argv[2]) if len(sys.argv) > 2 else None
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
defs = [a for a in sys.argv[4:] if a.startswith("-D")]
if which == "n1":
    mi = cases.methanol_readme_input("N1")
    cm = engine.compile_model(mi, method="ros4", block=block)
else:
    mi = cases.methanol_readme_input("N2")
    cm = engine.compile_model(mi, block=block or 128, lanes=lanes)
out = os.path.join(ROOT, "build", "gen", "_report")
os.makedirs(out, exist_ok=True)
open(os.path.join(out, "rmt_model.cuh"), "w").write(cm.header)
cmd = ["nvcc", "-Xptxas", "-v"] + build.NVCC_ARCH + ["-lineinfo", "-O3", "-std=c++17", "-DRMT_BLOCK=%d" % cm.block, "-I", out,
       "-cubin", "-o", os.path.join(out, "r.cubin"), os.path.join(build.CSRC, "rmt_kernels.cu")] + defs
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    print(r.stderr[-3000:]); sys.exit(1)
txt = r.stderr
for m in re.finditer(r"Compiling entry function '(\w+)'.*?\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt):
    print("%-22s regs %3s  stack %5s  spill st %5s ld %5s  smem %s" % (m.group(1), m.group(5), m.group(2), m.group(3), m.group(4), m.group(7)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(out, "r.cubin")], capture_output=True, text=True).stdout
cur = None; counts = {}
for line in sass.split("\n"):
    m = re.search(r"Function : (\w+)", line)
    if m: cur = m.group(1); counts[cur] = 0
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/", line): counts[cur] += 1
print({k: v for k, v in counts.items() if "solve" in k})
