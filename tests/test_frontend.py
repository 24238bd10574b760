"""Host-side front-end logic (no GPU): input row mapping, block-size choice, structural cache, API errors."""
import numpy as np
import pytest

import cases
from rmt_app_b200 import engine, rmtCom
from rmt_app_b200.componentdb import COMPONENTS, componentSymbolList
from rmt_app_b200.model import ModelSpec


def test_rmtcom_matches_reference_order():
    assert rmtCom() == "CO2,H2,CH3OH,H2O,CO,DME,N2,CH4,C2H4,C3H6,C3H8,C4H10"
    assert tuple(COMPONENTS) == componentSymbolList


def test_uniform_inputs_follow_header_order():
    mi = cases.methanol_testfile_input("N1")
    spec = ModelSpec(mi)
    u = engine.uniform_inputs(spec, mi)
    assert spec.input_names() == ["temperature", "pressure"] + ["concentration[%d]" % i for i in range(6)] + [
        "volumetric-flowrate", "ReInDi", "ReLe", "PaDi", "BeVoFr", "OvHeTrCo", "MeTe", "mixture-viscosity", "EfHeTrAr",
        "CaDe", "CaSpHeCa", "VARS:CaBeDe"]
    assert u.size == spec.nin == 20
    assert u[0] == 523 and u[1] == 5e6 and u[-1] == pytest.approx(1982*0.61)
    np.testing.assert_array_equal(u[2:8], mi["feed"]["concentration"])
    assert u[13] == 100 and u[14] == 522
    # the scalar VARS value comes from THIS modelInput even though the compiled model is shared
    mi2 = cases.methanol_readme_input("N1")
    cm1, cm2 = engine.compile_model(mi), engine.compile_model(mi2)
    assert cm1 is cm2
    assert engine.uniform_inputs(cm2.spec, mi2)[-1] == 1171.2


def test_sweep_rows_mapping_and_errors():
    mi = cases.methanol_readme_input("N1")
    mi["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
    spec = ModelSpec(mi)
    B = 5
    sw = {"pressure": np.arange(B) + 1.0, "concentration": np.arange(B*6, dtype=float).reshape(B, 6),
          "E1": -np.ones(B), "MeTe": np.full(B, 500.0)}
    rows, row_map = engine.sweep_rows(spec, sw, B)
    assert rows.shape == (9, B)
    assert spec.nin == 2 + 6 + 11 + 13
    assert row_map[1] == 0 and list(row_map[2:8]) == [1, 2, 3, 4, 5, 6]
    names = spec.input_names()
    assert row_map[names.index("VARS:E1")] == 7 and row_map[names.index("MeTe")] == 8
    assert (row_map == -1).sum() == spec.nin - 9
    np.testing.assert_array_equal(rows[3], sw["concentration"][:, 2])
    with pytest.raises(KeyError, match="neither"):
        engine.sweep_rows(spec, {"bogus": np.ones(B)}, B)
    with pytest.raises(ValueError, match="shape"):
        engine.sweep_rows(spec, {"pressure": np.ones(B + 1)}, B)
    with pytest.raises(ValueError, match="concentration"):
        engine.sweep_rows(spec, {"concentration": np.ones((B, 5))}, B)


def test_block_size_keeps_one_block_per_sm_within_shared_memory():
    from rmt_app_b200.codegen import system_size, use_extents
    for mk, n, m in ((lambda: cases.methanol_readme_input("N1"), 8, 5), (lambda: cases.ch4_input("N1"), 5, 3),
                     (lambda: cases.ch4_input("N1", "iso-thermal"), 4, 2), (cases.methanol_m7_input, 8, 5)):
        spec = ModelSpec(mk())
        assert spec.n == n
        # fewer reactions than species: the integrator works in reaction extents (nr + P + T unknowns)
        assert use_extents(spec) and system_size(spec) == m and system_size(spec, reduced=False) == n
        for reduced, dim in ((None, m), (False, n)):
            for stages in (4, 6):
                b = engine.default_block(spec, stages, reduced)
                assert b % 32 == 0 and 32 <= b <= 384
                assert b*8*(dim*dim + stages*dim) + 1024 <= 227*1024
    assert engine.default_block(ModelSpec(cases.methanol_readme_input("N1")), 6, False) == 256
    assert engine.default_block(ModelSpec(cases.methanol_readme_input("N1")), 4) == 384
    hdr = engine.compile_model(cases.methanol_readme_input("N1")).header
    assert "#define RMT_REDUCED 1" in hdr
    assert "#define RMT_REDUCED 0" in engine.compile_model(cases.methanol_readme_input("N1"), reduced=False).header
    assert "RMT_REDUCED 0" in engine.compile_model(cases.methanol_readme_input("N2")).header


def test_model_key_depends_on_structure_only():
    a = ModelSpec(cases.methanol_readme_input("N1")).key()
    b = ModelSpec(cases.methanol_testfile_input("N1")).key()
    c = ModelSpec(cases.methanol_readme_input("N2")).key()
    mi = cases.methanol_readme_input("N1")
    mi["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
    d = ModelSpec(mi).key()
    assert a == b and a != c and a != d


def test_m7_inputs():
    mi = cases.methanol_m7_input()
    spec = ModelSpec(mi)
    assert spec.model == "M7" and spec.n == 8 and not spec.iso
    u = engine.uniform_inputs(spec, mi)
    names = spec.input_names()
    assert u[names.index("mixture-viscosity")] == 1e-5 and u[names.index("EfHeTrAr")] == pytest.approx(4/0.0381)
    assert spec.kin.param_names == ["CaDe", "CaBeDe", "CaPo"]


def test_m9_inputs_and_launch_shape():
    mi = cases.methanol_m9_input()
    spec = ModelSpec(mi)
    assert spec.model == "M9" and spec.n == 7 and not spec.iso
    u = engine.uniform_inputs(spec, mi)
    names = spec.input_names()
    assert u[names.index("CaDe")] == 1982 and u[names.index("CaSpHeCa")] == pytest.approx(0.96)
    assert u[names.index("concentration[0]")] == pytest.approx(0.5749, rel=1e-3)        # kmol/m^3, as test_rmt_DME5.py
    cm = engine.compile_model_n2(mi, 100, 100)          # a few reactors: the lanes kernel, one lane per reactor
    assert cm.lanes == 1 and "#define RMT_MODEL_M9 1" in cm.header and "#define RMT_N2_G 1" in cm.header
    cp = engine.compile_model_n2(mi, 5000, 100)         # ensembles: the stage pipeline (64 reactors x 4 roles per block)
    assert cp.lanes == 0 and cp.block == 256 and "#define RMT_N2_G 0" in cp.header
    with pytest.raises(ValueError, match="one lane per reactor"):
        engine.compile_model(mi, lanes=4)
    assert engine.solverSetting["S2"] == {"tNo": 10, "zNo": 100, "rNo": 7, "timesNo": 5}
    # the steady-state models ignore the two extra inputs; their headers are unchanged in size
    assert ModelSpec(cases.methanol_readme_input("N1")).nin == spec.nin - spec.nkp + ModelSpec(cases.methanol_readme_input("N1")).nkp


def test_launch_shapes_and_pipeline_cuts():
    # lanes per reactor of the dynamic integrator: about three resident waves of threads, never more lanes than nodes
    assert [engine.n2_lanes(B, 200) for B in (1, 100, 4096, 12500, 50000, 10**6)] == [32, 32, 8, 8, 8, 1]
    assert engine.n2_lanes(1, 12) == 16 and engine.n2_lanes(1, 3) == 4
    assert engine.n2_lanes(12500, 200, n=3) == 4 and engine.n2_lanes(12500, 200, n=4) == 8
    assert engine.n2_block(12500, lanes=8) == 64 and engine.n2_block(1, lanes=32) == 32
    assert engine.n2_block(10**6) == 64 and engine.n2_block(5000) == 32 and engine.n2_block(8000) == 64
    # N2: the stage pipeline takes ensembles that fill its rounds of 148 x 64 reactors, the lanes kernel the rest
    assert [engine.n2_use_pipeline(B, 200) for B in (500, 2048, 9472, 12500, 18944, 50000, 100000)] == \
        [False, False, True, False, True, True, True]
    assert not engine.n2_use_pipeline(9472, 4)
    mi2 = cases.methanol_readme_input("N2")
    assert engine.compile_model_n2(mi2, 12500, 200).lanes == 8 and engine.compile_model_n2(mi2, 9472, 200).lanes == 0
    assert engine.compile_model_n2(mi2, 9472, 200).block == 256
    with pytest.raises(ValueError):
        engine.compile_model(mi2, block=64, lanes=3)
    # copy/compute pipeline: chunks cover the ensemble exactly, are non-empty and 1024-aligned inside
    for B in (3*1024, 5000, 1 << 18, (1 << 20) + 7, 10**7):
        cuts = engine.pipeline_cuts(B)
        assert cuts[0] == 0 and cuts[-1] == B and all(b > a for a, b in zip(cuts, cuts[1:]))
        assert all(c % 1024 == 0 for c in cuts[1:-1]) and 2 <= len(cuts) <= 4
    assert engine.pipeline_cuts(1 << 20) == [0, 131072, 917504, 1048576]
    assert engine.pipeline_cuts(2048, split=(0.5, 0.5)) == [0, 1024, 2048]
    assert engine.pipeline_cuts(1000, split=(0.5, 0.5)) == [0, 1000]


def test_automatic_method_choice():
    mi = cases.methanol_readme_input("N1")
    assert engine.choose_method(mi, 1e-3, 1) == "ros4"              # outlet only, loose tolerance
    assert engine.choose_method(mi, 1e-3, 101) == "rodas4"          # profile: needs dense output
    assert engine.choose_method(mi, 1e-3, 101, dense=False) == "ros4"
    assert engine.choose_method(mi, 1e-6, 1) == "rodas4"            # tight tolerance
    mi["solver-config"]["method"] = "rodas3"
    assert engine.choose_method(mi, 1e-3, 1) == "rodas3"
    mi["solver-config"]["method"] = "euler"
    with pytest.raises(ValueError, match="solver-config.method"):
        engine.choose_method(mi, 1e-3, 1)


def test_n_unknowns():
    assert ModelSpec(cases.methanol_readme_input("N2")).n == 7
    assert ModelSpec(cases.ch4_input("N2", "iso-thermal")).n == 3
    assert ModelSpec(cases.ch4_input("N1", "iso-thermal")).n == 4


K0_GLOBAL = 7.2e-4
_TABLE = np.array([1.0, 2.0])


def _ch4_with_global_kinetics(extra_vars=None):
    mi = cases.ch4_input("N1")
    varis = {"C": lambda x: x['SpCoi'][0]}
    varis.update(extra_vars or {})
    mi["reaction-rates"] = {"VARS": varis, "RATES": {"r1": lambda x: K0_GLOBAL*_TABLE[1]*(x['C']**2)}}
    return mi


def test_compile_cache_sees_globals_closures_and_baked_vars():
    """the fast cache key must change when anything the tracer bakes into the graph changes — a module-level
    global a lambda reads, the contents of an array it indexes, a closure cell, a non-scalar VARS entry — and must NOT
    change for scalar VARS values (those are run-time kinetic-parameter slots)."""
    global K0_GLOBAL
    cm_a = engine.compile_model(_ch4_with_global_kinetics())
    assert engine.compile_model(_ch4_with_global_kinetics()) is cm_a
    old = K0_GLOBAL
    try:
        K0_GLOBAL = 9.9e-4                                  # rebinding a global constant between two calls
        cm_b = engine.compile_model(_ch4_with_global_kinetics())
        assert cm_b is not cm_a and cm_b.header != cm_a.header
        assert "0.00099" in cm_b.header.replace("9.9e-04", "0.00099").replace("9.9000000000000002e-04", "0.00099") \
            or cm_b.spec.key() != cm_a.spec.key()
    finally:
        K0_GLOBAL = old
    assert engine.compile_model(_ch4_with_global_kinetics()) is cm_a
    _TABLE[1] = 3.0                                         # mutating an array the lambda indexes
    try:
        cm_c = engine.compile_model(_ch4_with_global_kinetics())
        assert cm_c is not cm_a and cm_c.spec.key() != cm_a.spec.key()
    finally:
        _TABLE[1] = 2.0

    def with_closure(k):
        mi = cases.ch4_input("N1")
        scale = [k]                                         # a mutable object in a closure cell
        mi["reaction-rates"] = {"VARS": {"C": lambda x: x['SpCoi'][0]}, "RATES": {"r1": lambda x: scale[0]*(x['C']**2)}}
        return mi
    assert engine.compile_model(with_closure(1.0)).spec.key() != engine.compile_model(with_closure(2.0)).spec.key()
    # non-scalar VARS entries are constants of the graph; scalar ones are parameter slots
    def with_vars(extra, rate):
        mi = _ch4_with_global_kinetics(extra)
        mi["reaction-rates"]["RATES"] = {"r1": rate}
        return mi
    a = engine.compile_model(with_vars({"w": [1.0, 2.0]}, lambda x: x['w'][1]*(x['C']**2)))
    b = engine.compile_model(with_vars({"w": [1.0, 5.0]}, lambda x: x['w'][1]*(x['C']**2)))
    assert a is not b and a.spec.key() != b.spec.key()
    p1 = engine.compile_model(with_vars({"k": 1.0}, lambda x: x['k']*(x['C']**2)))
    p2 = engine.compile_model(with_vars({"k": 2.0}, lambda x: x['k']*(x['C']**2)))
    assert p1 is p2 and p1.spec.kin.param_names == ["k"]
