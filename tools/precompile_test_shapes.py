"""Fill the in-tree cubin cache (rmt_app_b200/_cubin_cache, travels with gpurun) with every model / launch shape the
GPU tests and bench.py use, so that the GPU box does not spend its time in NVRTC.  Called by `__graft_entry__.build()`;
can be run on its own."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main(verbose=True):
    import cases
    from rmt_app_b200 import engine, capi
    n = [0]

    def put(cm):
        capi.cached_cubin(cm.header, block=cm.block)
        n[0] += 1

    for mk in (cases.methanol_testfile_input, cases.ch4_input):
        mi = mk("N2")
        for lanes, block in ((1, 64), (4, 32), (8, 64), (32, 32)):
            put(engine.compile_model(mi, block=block, lanes=lanes))
        for B, z in ((1, 20), (1, 50), (40, 16), (70, 12), (1, 12), (1, 16)):
            put(engine.compile_model_n2(mi, B, z))
    mi = cases.ch4_input("N2", "iso-thermal")
    for lanes, block in ((1, 64), (8, 64), (32, 32)):
        put(engine.compile_model(mi, block=block, lanes=lanes))
    for B, z in ((1, 20), (3, 20)):
        put(engine.compile_model_n2(mi, B, z))
    mi = cases.methanol_readme_input("N2")
    for B, z in ((1, 20), (1, 50), (12500, 200), (1, 200)):
        put(engine.compile_model_n2(mi, B, z))
    for B, z in ((9472, 200), (100000, 200), (50000, 50)):                # stage-pipelined launch shapes
        put(engine.compile_model_n2(mi, B, z))
    for mk in (cases.methanol_testfile_input, cases.ch4_input, lambda m: cases.ch4_input(m, "iso-thermal")):
        put(engine.compile_model(mk("N2"), block=256, lanes=0))
    mi = cases.methanol_m9_input()
    for blk in (64, 32):
        put(engine.compile_model(mi, block=blk))
    put(engine.compile_model(mi, block=256, lanes=0))
    base4 = cases.methanol_readme_input("N1"); base4["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
    for m in ("ros4", "rodas4"):
        put(engine.compile_model(base4, method=m))
        put(engine.compile_model(cases.methanol_readme_input("N1"), method=m, reduced=False))
        put(engine.compile_model(cases.ch4_three_reaction_input("N1"), method=m))
    for mk in (lambda: cases.methanol_testfile_input("N1"), lambda: cases.ch4_input("N1"), lambda: cases.ch4_input("N1", "iso-thermal")):
        for m in ("ros4", "rodas4"):
            for red in (None, False):
                put(engine.compile_model(mk(), method=m, reduced=red))
    if verbose:
        print("[build] cubin cache holds the %d test / bench launch shapes" % n[0])


if __name__ == "__main__":
    main()
