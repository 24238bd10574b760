"""ncu reports -> profiles/r02_calibration.json (+ one summary CSV per report under profiles/).

bench.py quotes a few numbers that only a profiler can measure — DRAM bytes per launch, executed FP64 thread
instructions, FP64 pipe utilisation.  They are read from `ncu --set full` captures of the SAME kernels on the SAME
workload and stored here together with `module_key`, the identity of the cubin (capi.cubin_key: hash of the generated
header + rmt_kernels.cu + options) the capture was taken from.  bench.py recomputes the key of the module it runs and
prints `calibration_stale: true` when they differ, so a number can never silently outlive the kernel it describes.
Run right after pulling the reports, before touching the kernels again.

usage: python tools/make_calibration.py [--n1 REP --n1-attempts-per-solve A] [--rhsjac REP] [--n2 REP --n2-attempts-per-reactor A]
                                         [--instances 1048576] [--n2-instances 12500 --n2-nodes 200]
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from rmt_app_b200 import engine  # noqa: E402

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
KEEP += ["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op for op in ("dfma", "dmul", "dadd")] + ["smsp__cycles_elapsed.avg", "smsp__inst_executed.sum"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {}
        for k, u, v in zip(head, units, r):
            if k == "Kernel Name":
                d[k] = v
            elif k in KEEP:
                try:
                    d[k] = float(v.replace(",", ""))*UNIT_SCALE.get(u, 1.0)
                except ValueError:
                    d[k] = v
        recs.append(d)
    return recs


def thread_inst(r, op):
    """Executed thread instructions of one FP64 opcode: the .sum counter, or (what `--set full` stores) its per-cycle
    rate times the elapsed SM-sub-partition cycles."""
    k = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op
    if k in r:
        return r[k]
    return r[k + ".per_cycle_elapsed"]*r["smsp__cycles_elapsed.avg"]


def write_summary(recs, path, title):
    with open(path, "w") as f:
        f.write("# %s\n# values in SI units (bytes, seconds); one column per profiled launch\n" % title)
        f.write("metric," + ",".join(r["Kernel Name"] for r in recs) + "\n")
        for k in KEEP:
            if any(k in r for r in recs):
                f.write(k + "," + ",".join(repr(r.get(k, "")) for r in recs) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n1"); ap.add_argument("--n1-attempts-per-solve", type=float)
    ap.add_argument("--rhsjac")
    ap.add_argument("--n2"); ap.add_argument("--n2-attempts-per-reactor", type=float)
    ap.add_argument("--instances", type=int, default=1 << 20)
    ap.add_argument("--n2-instances", type=int, default=12500); ap.add_argument("--n2-nodes", type=int, default=200)
    ap.add_argument("--n2wf"); ap.add_argument("--n2wf-attempts-per-reactor", type=float); ap.add_argument("--n2wf-instances", type=int, default=9472)
    ap.add_argument("--tag", default="r02")
    a = ap.parse_args()
    path = os.path.join(ROOT, "profiles", "%s_calibration.json" % a.tag)
    cal = {"kernels": {}}
    if os.path.exists(path):
        cal = json.load(open(path))
    base = cases.methanol_readme_input("N1")
    cm1 = engine.compile_model(base, method="ros4")
    if a.n1:
        recs = [r for r in raw_rows(a.n1) if r["Kernel Name"] == "rmt_n1_solve"]
        out = os.path.join("profiles", "%s_ncu_n1_solve.csv" % a.tag)
        write_summary(recs, os.path.join(ROOT, out), "rmt_n1_solve, %d config-3 reactors, Ros4, rtol 1e-3 (ncu --set full --clock-control none)" % a.instances)
        r = recs[-1]
        att = a.n1_attempts_per_solve*a.instances
        cal["kernels"]["rmt_n1_solve"] = {
            "module_key": cm1.key(), "source": out, "instances": a.instances, "attempts_per_solve": a.n1_attempts_per_solve,
            "dfma_per_attempt": thread_inst(r, "dfma")/att,
            "dmul_per_attempt": thread_inst(r, "dmul")/att,
            "dadd_per_attempt": thread_inst(r, "dadd")/att,
            "warp_instructions_per_attempt": r.get("sm__inst_executed.sum", r.get("smsp__inst_executed.sum"))*r.get("smsp__thread_inst_executed_per_inst_executed.ratio", 32.0)/att,
            "dram_bytes_per_reactor": (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"])/a.instances,
            "fp64_pipe_busy": r["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]/100.0,
            "registers": r["launch__registers_per_thread"], "ncu_duration_s": r["gpu__time_duration.sum"]}
    if a.rhsjac:
        recs = raw_rows(a.rhsjac)
        out = os.path.join("profiles", "%s_ncu_n1_rhs_jac.csv" % a.tag)
        write_summary(recs, os.path.join(ROOT, out), "rmt_n1_rhs / rmt_n1_jac, %d config-3 reactors, L2 flushed before each launch" % a.instances)
        for name in ("rmt_n1_rhs", "rmt_n1_jac"):
            rr = [r for r in recs if r["Kernel Name"] == name]
            if not rr:
                continue
            r = rr[-1]
            cal["kernels"][name] = {
                "module_key": cm1.key(), "source": out, "instances": a.instances,
                "dram_bytes_per_launch": r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"],
                "dram_bytes_read": r["dram__bytes_read.sum"], "dram_bytes_written": r["dram__bytes_write.sum"],
                "fp64_pipe_busy": r["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]/100.0,
                "registers": r["launch__registers_per_thread"], "ncu_duration_s": r["gpu__time_duration.sum"]}
    if a.n2:
        recs = [r for r in raw_rows(a.n2) if r["Kernel Name"] == "rmt_n2_solve"]
        out = os.path.join("profiles", "%s_ncu_n2_solve.csv" % a.tag)
        write_summary(recs, os.path.join(ROOT, out), "rmt_n2_solve, %d reactors x %d nodes (config-5 share), 8 lanes per reactor" % (a.n2_instances, a.n2_nodes))
        r = recs[-1]
        mi2 = cases.methanol_readme_input("N2")
        cm2 = engine.compile_model_n2(mi2, a.n2_instances, a.n2_nodes)
        natt = a.n2_attempts_per_reactor*a.n2_instances*a.n2_nodes
        fl = 2*thread_inst(r, "dfma") + thread_inst(r, "dmul") \
            + thread_inst(r, "dadd")
        cal["kernels"]["rmt_n2_solve"] = {
            "module_key": cm2.key(), "source": out, "instances": a.n2_instances, "nodes": a.n2_nodes,
            "attempts_per_reactor": a.n2_attempts_per_reactor,
            "executed_fp64_flop_per_node_attempt": fl/natt,
            "warp_instructions_per_node_attempt": r.get("sm__inst_executed.sum", r.get("smsp__inst_executed.sum"))*32.0/natt,
            "dram_bytes_per_launch": r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"],
            "dram_bytes_per_node_attempt": (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"])/natt,
            "fp64_pipe_busy": r["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]/100.0,
            "local_loads": r.get("sass__inst_executed_local_loads"), "local_stores": r.get("sass__inst_executed_local_stores"),
            "registers": r["launch__registers_per_thread"], "ncu_duration_s": r["gpu__time_duration.sum"]}
    if a.n2wf:
        recs = [r for r in raw_rows(a.n2wf) if r["Kernel Name"] == "rmt_n2_solve"]
        out = os.path.join("profiles", "%s_ncu_n2_solve_pipeline.csv" % a.tag)
        write_summary(recs, os.path.join(ROOT, out), "rmt_n2_solve (stage pipeline), %d reactors x %d nodes = one full round" % (a.n2wf_instances, a.n2_nodes))
        r = recs[-1]
        mi2 = cases.methanol_readme_input("N2")
        cm2 = engine.compile_model_n2(mi2, a.n2wf_instances, a.n2_nodes)
        assert cm2.lanes == 0
        natt = a.n2wf_attempts_per_reactor*a.n2wf_instances*a.n2_nodes
        fl = 2*thread_inst(r, "dfma") + thread_inst(r, "dmul") \
            + thread_inst(r, "dadd")
        cal["kernels"]["rmt_n2_solve_pipeline"] = {
            "module_key": cm2.key(), "source": out, "instances": a.n2wf_instances, "nodes": a.n2_nodes,
            "attempts_per_reactor": a.n2wf_attempts_per_reactor,
            "executed_fp64_flop_per_node_attempt": fl/natt,
            "warp_instructions_per_node_attempt": r.get("sm__inst_executed.sum", r.get("smsp__inst_executed.sum"))*32.0/natt,
            "dram_bytes_per_launch": r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"],
            "dram_bytes_per_node_attempt": (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"])/natt,
            "fp64_pipe_busy": r["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]/100.0,
            "local_loads": r.get("sass__inst_executed_local_loads"), "local_stores": r.get("sass__inst_executed_local_stores"),
            "registers": r["launch__registers_per_thread"], "ncu_duration_s": r["gpu__time_duration.sum"]}
    with open(path, "w") as f:
        json.dump(cal, f, indent=1)
    print(json.dumps(cal, indent=1)[:3000])


if __name__ == "__main__":
    main()
