"""Launch the stand-alone RHS / Jacobian kernels on the bench's 2^20 config-3 reactors (same buffers and states as
bench.py's roofline_rhs_kernel / roofline_jac_kernel legs).  Run plain for the CUDA-event times, under
`ncu --set full -k regex:rmt_n1_(rhs|jac)` for dram__bytes (profiles/r02_ncu_n1_rhs_jac.csv)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import torch  # noqa: E402
from rmt_app_b200 import engine  # noqa: E402

B = 1 << 20
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
base = cases.methanol_readme_input("N1")
sweep = cases.config3_sweep(B, 20240611)
cm = engine.compile_model(base, method="ros4")
mod = cm.load(0)
info, spec = mod.info, cm.spec
n = info.n
ws = engine.Workspace()
h_rows, n_rows, row_map = engine.sweep_rows_into(spec, sweep, B, ws)
uniform = engine.uniform_inputs(spec, base)
d_rows = h_rows.to(dev)
d_consts = torch.empty((info.nconst, B), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
mod.setup(B, d_rows, n_rows, row_map, uniform, d_consts, stream=stream)
g = torch.Generator(device=dev); g.manual_seed(1)
d_y = torch.rand((n, B), dtype=torch.float64, device=dev, generator=g)*0.5 + 0.25
d_y[n - 2] = 1.0
d_y[n - 1] = 0.1
d_f = torch.empty((n, B), dtype=torch.float64, device=dev)
d_J = torch.empty((n*n, B), dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("rmt_n1_rhs", lambda: mod.n1_rhs(B, d_consts, d_y, d_f, stream=stream)),
                 ("rmt_n1_jac", lambda: mod.n1_jac(B, d_consts, d_y, d_f, d_J, stream=stream))):
    ts = []
    for _ in range(REPS):
        flush.zero_()                               # 256 MB > 126 MB L2: the next launch reads from HBM
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "ms per launch (L2 flushed):", ["%.4f" % t for t in ts], "min %.4f" % min(ts))
assert torch.isfinite(d_f).all() and torch.isfinite(d_J).all()
print("ok")
