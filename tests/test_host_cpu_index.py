"""index_state of the host mirror (external.jl:275-365) for every model family, without a GPU"""
import types

import pytest

from petlion_b200.api import Model


@pytest.mark.parametrize("temperature,aging,grid,ntot", [(False, False, 10, 301), (True, False, 10, 351),
                                                          (False, "SEI", 10, 322), (False, True, 20, 642),
                                                          (False, False, 20, 601)])
def test_index_state_covers_the_state_vector(temperature, aging, grid, ntot):
    m = Model.__new__(Model)
    m.N = types.SimpleNamespace(p=grid, s=grid, n=grid, a=10, z=10, r_p=10, r_n=10, tot=ntot)
    m.numerics = types.SimpleNamespace(temperature=temperature, aging=aging)
    ind = m._index_state()
    assert ind["I"] == slice(ntot - 1, ntot)
    assert ind["c_e"] == slice(0, 3 * grid)
    assert ("T" in ind) == bool(temperature) and ("j_s" in ind) == bool(aging)
    # App. A of SURVEY.md: differential block c_e, c_s_avg, T, film, SOH; algebraic block j, Phi_e, Phi_s, j_s, I
    order = [k for k in ("c_e", "c_s_avg", "T", "film", "SOH", "j", "Φ_e", "Φ_s", "j_s", "I") if k in ind]
    stops = [ind[k].stop for k in order]
    assert stops == sorted(stops) and all(ind[a].stop == ind[b].start for a, b in zip(order[:-1], order[1:]))


def test_table_host_semantics_match_the_oracle_table():
    """petlion_b200.Table (host side of plb_input_table) evaluates like the oracle's table: right-continuous at a
    repeated knot, linear between knots, constant outside"""
    import numpy as np
    import oracle as O
    from petlion_b200 import Table
    t = Table([0.0, 100.0, 100.0, 200.0], [1.0, 1.0, 0.5, 0.25])
    assert t.jumps() == [100.0]
    for x in (-1.0, 0.0, 50.0, 99.999999, 100.0, 100.000001, 150.0, 200.0, 1e6):
        assert t(x) == O.table_eval((t.t, t.v), x)
    assert t(100.0) == 0.5 and t(99.0) == 1.0
    s = Table.sample(lambda x: 2.0 * x, np.linspace(0.0, 1.0, 5), scale=3.0)
    assert s(0.3) == pytest.approx(0.6) and s.scale == 3.0
    with pytest.raises(ValueError):
        Table([0.0, 1.0], [1.0])
    with pytest.raises(ValueError):
        Table([1.0, 0.0], [1.0, 1.0])


@pytest.mark.parametrize("kw,msg", [
    (dict(N_r_p=12), "N_r_p = N_r_n = 10"),
    (dict(N_r_p=11, N_r_n=11), "N_r_p = N_r_n = 10"),
    (dict(N_r_p=16, N_r_n=16), "N_r_p = N_r_n = 10"),
    (dict(N_r_p=12, N_r_n=12, N_p=20, N_s=20, N_n=20), "N_r = 12 / 14 is built for"),
    (dict(N_r_p=14, N_r_n=14, temperature=True, aging="SEI"), "N_r = 12 / 14 is built for"),
    (dict(N_r_p=12, N_r_n=12, rxn_p="rxn_MHC"), "N_r = 12 / 14 is built for"),
    (dict(Fickian_method="spectral", N_p=20, N_s=20, N_n=20), "Fickian_method = :spectral is built for"),
    (dict(Fickian_method="spectral", temperature=True, aging="SEI"), "Fickian_method = :spectral is built for"),
    (dict(Fickian_method="spectral", N_r_p=12, N_r_n=12), "N_r = 12 / 14 is built for"),
    (dict(N_p=30, N_s=10, N_n=30), "<= 64"),
    (dict(N_p=1), "2 <= N_p"),
    (dict(temperature=True, N_p=4), "N_p, N_n >= 5"),
    (dict(temperature=True, N_a=20, N_z=20), "N_a\\+N_z <="),
])
def test_unsupported_model_options_are_refused_with_a_message(kw, msg):
    """plb_create validates the model description before it touches the device: every unsupported option set of
    SURVEY section 8 is an error with a message, never a silent fallback"""
    import petlion_b200
    with pytest.raises(RuntimeError, match=msg):
        petlion_b200.petlion("LCO", **kw)


def test_nmc_has_no_thermal_or_aging_parameters():
    import petlion_b200
    with pytest.raises(RuntimeError, match="LCO or NMC_LGM50 parameter set"):
        petlion_b200.petlion("NMC", temperature=True)
    with pytest.raises(RuntimeError, match="LCO parameter set"):
        petlion_b200.petlion("NMC", aging="SEI")


_G = dict(N_p=20, N_s=20, N_n=20)
_MHC = dict(rxn_p="rxn_MHC", rxn_n="rxn_MHC")


@pytest.mark.parametrize("cathode,kw", [
    ("LCO", {}), ("LCO", dict(temperature=True)), ("LCO", dict(aging="SEI")), ("LCO", dict(temperature=True, aging="SEI")),
    ("LCO", _G), ("LCO", dict(temperature=True, **_G)), ("LCO", dict(aging="SEI", **_G)), ("LCO", dict(temperature=True, aging="SEI", **_G)),
    ("LCO", _MHC), ("LCO", dict(temperature=True, **_MHC)), ("LCO", dict(aging="SEI", **_MHC)), ("LCO", dict(temperature=True, aging="SEI", **_MHC)),
    ("LCO", dict(**_G, **_MHC)), ("LCO", dict(temperature=True, **_G, **_MHC)), ("LCO", dict(aging="SEI", **_G, **_MHC)),
    ("LCO", dict(temperature=True, aging="SEI", **_G, **_MHC)), ("LCO", dict(rxn_n="rxn_MHC", N_p=20, N_s=10, N_n=20)),
    ("NMC", {}), ("NMC", _G), ("NMC", dict(N_r_p=14, N_r_n=14)), ("NMC", dict(Fickian_method="spectral")),
    ("NMC_LGM50", {}), ("NMC_LGM50", dict(temperature=False)), ("NMC_LGM50", _G), ("NMC_LGM50", dict(temperature=False, **_G)),
    ("LCO", dict(N_r_p=12, N_r_n=12)), ("LCO", dict(N_r_p=12, N_r_n=12, temperature=True)), ("LCO", dict(N_r_p=12, N_r_n=12, aging="SEI")),
    ("LCO", dict(N_r_p=14, N_r_n=14)), ("LCO", dict(N_r_p=14, N_r_n=14, temperature=True)), ("LCO", dict(N_r_p=14, N_r_n=14, aging="SEI")),
    ("LCO", dict(Fickian_method="spectral")), ("LCO", dict(Fickian_method="spectral", temperature=True)),
    ("LCO", dict(Fickian_method="spectral", aging="SEI")),
    ("LCO", dict(N_p=7, N_s=5, N_n=9)), ("LCO", dict(N_p=20, N_s=24, N_n=20)), ("LCO", dict(N_p=15, N_s=2, N_n=15)),
])
def test_every_built_option_set_finds_its_family(cathode, kw):
    """plb_create validates the options and looks the compiled family up BEFORE it touches a device: on a machine without a GPU
    every built option set gets as far as "no CUDA device available" (and with one it simply succeeds)"""
    import petlion_b200
    try:
        petlion_b200.petlion(cathode, **kw)
    except RuntimeError as e:
        assert "no CUDA device available" in str(e), str(e)
