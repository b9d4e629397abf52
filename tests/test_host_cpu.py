"""CPU tests (no GPU needed): C-ABI library loads and exports every declared symbol, generated
constitutive laws match their definitions, host-side sharding logic incl. a world_size-2 gloo run."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import petlion_b200
    from petlion_b200 import _lib
    path = petlion_b200.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "petlion_b200.h")).read()
    declared = set(re.findall(r"\b(plb_[a-z_0-9A-Z]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import petlion_b200
    with pytest.raises(RuntimeError, match="no CUDA device"):
        petlion_b200.petlion("LCO")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "petlion.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def _load_generated():
    src = open(os.path.join(ROOT, "petlion.jl_b200", "csrc", "laws_generated.cuh")).read()

    def run(name, **kw):
        i = src.index("void " + name + "(")
        body = src[src.index("{", i) + 1: src.index("\n}", i)]
        code = "\n".join(ln.strip().replace("const double ", "") for ln in body.strip().split("\n"))
        env = dict(kw)
        exec(code, {"exp": math.exp, "atan": math.atan, "sqrt": math.sqrt, "__drcp_rn": lambda x: 1.0 / x}, env)
        return env
    return run


def test_generated_laws_match_reference_formulas():
    """custom_functions.jl:83,96,123-174 restated directly; derivatives against central differences"""
    run = _load_generated()
    t = 0.7
    U = (-4.656 + 88.669 * t**2 - 401.119 * t**4 + 342.909 * t**6 - 462.471 * t**8 + 433.434 * t**10) / \
        (-1 + 18.933 * t**2 - 79.532 * t**4 + 37.311 * t**6 - 73.083 * t**8 + 95.96 * t**10)
    assert run("OCV_LCO", th=t)["U"] == pytest.approx(U, rel=1e-14)
    c, T = 1100.0, 300.0
    K = 1e-4 * c * ((-10.5 + 0.668e-3 * c + 0.494e-6 * c**2) + (0.074 - 1.78e-5 * c - 8.86e-10 * c**2) * T
                    + (-6.96e-5 + 2.8e-8 * c) * T**2) ** 2
    assert run("K_eff", c=c, T=T)["K"] == pytest.approx(K, rel=1e-13)
    D = 1e-4 * 10.0 ** (-4.43 - 54.0 / (T - 229 - 5e-3 * c) - 0.22e-3 * c)
    assert run("D_eff_nl", c=c, T=T)["D"] == pytest.approx(D, rel=1e-13)
    t = 0.3
    Un = 0.7222 + 0.1387 * t + 0.029 * math.sqrt(t) - 0.0172 / t + 0.0019 / (math.sqrt(t) * t) + \
        0.2808 * math.exp(0.9 - 15 * t) - 0.7984 * math.exp(0.4465 * t - 0.4108)
    assert run("OCV_LiC6", th=t, s=math.sqrt(t))["U"] == pytest.approx(Un, rel=1e-14)
    h = 1e-6
    for name, kw, f, df in (("OCV_LCO", dict(th=0.7), "U", "dU"), ("OCV_LCO", dict(th=0.7), "dUdT", "ddUdT"),
                            ("OCV_NMC", dict(th=0.6), "U", "dU"), ("OCV_LiC6_NMC", dict(th=0.4), "U", "dU")):
        a = run(name, **kw)
        b = run(name, th=kw["th"] + h); cm = run(name, th=kw["th"] - h)
        assert a[df] == pytest.approx((b[f] - cm[f]) / (2 * h), rel=2e-6)
    a = run("OCV_LiC6", th=0.3, s=math.sqrt(0.3))
    b = run("OCV_LiC6", th=0.3 + h, s=math.sqrt(0.3 + h)); cm = run("OCV_LiC6", th=0.3 - h, s=math.sqrt(0.3 - h))
    assert a["dU"] == pytest.approx((b["U"] - cm["U"]) / (2 * h), rel=2e-6)
    assert a["ddUdT"] == pytest.approx((b["dUdT"] - cm["dUdT"]) / (2 * h), rel=2e-6)
    a = run("K_eff", c=1100.0, T=300.0); b = run("K_eff", c=1100.0 + 1e-3, T=300.0); cm = run("K_eff", c=1100.0 - 1e-3, T=300.0)
    assert a["dKdc"] == pytest.approx((b["K"] - cm["K"]) / 2e-3, rel=1e-6)
    a = run("K_eff_T", c=1100.0, T=310.0); b = run("K_eff_T", c=1100.0, T=310.0 + 1e-3); cm = run("K_eff_T", c=1100.0, T=310.0 - 1e-3)
    assert a["dKdT"] == pytest.approx((b["K"] - cm["K"]) / 2e-3, rel=1e-6)
    assert a["K"] == run("K_eff", c=1100.0, T=310.0)["K"]
    a = run("D_eff_nl", c=1100.0, T=300.0); b = run("D_eff_nl", c=1100.0 + 1e-3, T=300.0); cm = run("D_eff_nl", c=1100.0 - 1e-3, T=300.0)
    assert a["dDdc"] == pytest.approx((b["D"] - cm["D"]) / 2e-3, rel=1e-6)


def test_generated_particle_operator_matches_oracle_residual():
    """MC/BJ of laws_generated.cuh against the oracle's Fickian FD residual (residuals.jl:128-180)"""
    import oracle as O
    src = open(os.path.join(ROOT, "petlion.jl_b200", "csrc", "laws_generated.cuh")).read()
    rows = re.findall(r"^\s+\{([^}]*)\},$", src, flags=re.M)
    MC = np.array([[float(x) for x in r.split(",")] for r in rows[:10]])      # MC, then EV and EVI
    EV = np.array([[float(x) for x in r.split(",")] for r in rows[10:20]])
    EVI = np.array([[float(x) for x in r.split(",")] for r in rows[20:30]])
    EL = np.array([float(x) for x in re.search(r"EL\[NR\] = \{([^}]*)\}", src).group(1).split(",")])
    # eigen-basis of the particle operator used by the thermal variant's solver
    np.testing.assert_allclose(EV @ np.diag(EL) @ EVI, MC, rtol=0, atol=1e-11 * np.abs(MC).max())
    np.testing.assert_allclose(EV @ EVI, np.eye(10), atol=1e-13)
    BJ = float(re.search(r"BJ = ([0-9.eE+-]+);", src).group(1))
    assert MC.shape == (10, 10)
    m = O.make_model("LCO"); th = O.theta_defaults("LCO"); L = O.layout(m); names = O.theta_names()
    rng = np.random.default_rng(0)
    Y = O.initial_guess(m, th, 0.5)
    Y[L.c_s_p:L.c_s_p + 10] *= 1 + 0.05 * rng.uniform(-1, 1, 10)
    Y[L.j] = 1.3e-5
    res = O.residual(m, th, O.make_run("I", 0.0), 0.0, Y, np.zeros_like(Y))
    Ds, Rp = th[names.index("D_sp")], th[names.index("Rp_p")]
    kap = Ds / Rp**2
    d1bc = -Y[L.j] * Rp / Ds
    rhs = kap * (MC @ Y[L.c_s_p:L.c_s_p + 10])
    rhs[9] += kap * BJ * d1bc
    np.testing.assert_allclose(res[L.c_s_p:L.c_s_p + 10], rhs, rtol=1e-9, atol=1e-12 * np.abs(rhs).max())


def test_shard_bounds_cover_batch():
    from petlion_b200.sharding import shard_bounds
    for B in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import oracle as O
from tests import util
from petlion_b200.sharding import shard_bounds, gather_summaries
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B = 11
lo, hi = shard_bounds(B, world, rank)
tho = util.oracle_theta_batch(B)[lo:hi]           # every rank draws its own systems from the counter RNG
# the CPU stand-in for the GPU integrate step on this rank's shard (host logic under test = sharding + gather)
r = O.simulate_batch(O.make_model("LCO"), tho, O.make_run("I", -1.0, tf=600.0), O.default_opts(),
                     O.default_bounds("LCO"), SOC0=1.0)
local = torch.tensor(np.stack([r["t_end"], r["V_end"], r["I_end"], r["SOC_end"], r["flag"].astype(float),
                               r["n_steps"].astype(float), r["n_res"].astype(float), r["n_jac"].astype(float)], axis=1))
full = gather_summaries(local, B)
assert full.shape == (B, 8)
if rank == 0:
    np.save({out!r}, full.numpy())
dist.destroy_process_group()
'''


def test_gloo_world2_sharded_run_equals_single(tmp_path):
    """world_size-2 gloo: sharded run + all-gather of summaries == the unsharded run"""
    import oracle as O
    from tests import util
    out = str(tmp_path / "gathered.npy")
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)], env=env, timeout=300)
    full = np.load(out)
    r = O.simulate_batch(O.make_model("LCO"), util.oracle_theta_batch(11), O.make_run("I", -1.0, tf=600.0),
                         O.default_opts(), O.default_bounds("LCO"), SOC0=1.0)
    np.testing.assert_array_equal(full[:, 0], r["t_end"])
    np.testing.assert_array_equal(full[:, 1], r["V_end"])
    np.testing.assert_array_equal(full[:, 5], r["n_steps"])


def test_julia_shim_covers_the_header():
    """julia/PETLIONB200.jl (the ccall layer of INTEGRATION.md; Julia itself is not in the image) binds every entry
    point include/petlion_b200.h declares, mirrors the struct fields in order, and is bracket-balanced"""
    hdr = open(os.path.join(ROOT, "include", "petlion_b200.h")).read()
    jl = open(os.path.join(ROOT, "julia", "PETLIONB200.jl")).read()
    declared = set(re.findall(r"\b(plb_[a-z_0-9A-Z]+)\s*\(", hdr))
    bound = set(re.findall(r"\(:(plb_[a-z_0-9A-Z]+), lib\)", jl))
    assert declared == bound, declared ^ bound
    code = re.sub(r'""".*?"""', "", jl, flags=re.S)
    code = "\n".join(ln.split("#")[0] for ln in code.split("\n"))
    code = re.sub(r'"[^"\n]*"', '""', code)
    for a, b in ("()", "[]", "{}"):
        assert code.count(a) == code.count(b), (a, code.count(a), code.count(b))
    assert len(re.findall(r"^\s*(module|struct|function|begin|if)\b", code, flags=re.M)) + len(re.findall(r"\bbegin\s*$", code, flags=re.M)) >= len(re.findall(r"^\s*end\b", code, flags=re.M)) - 1

    def c_fields(name):
        end = hdr.index("} " + name + ";")
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                out += [x.strip().lstrip("*") for x in decl.split(None, 1)[1].replace("const double", "").split(",")]
        return [x.split()[-1].lstrip("*") for x in out]

    def jl_fields(name):
        body = re.search(r"struct " + name + r"\n(.*?)\nend", jl, flags=re.S).group(1)
        return [x.split("::")[0].strip() for x in re.split(r"[;\n]", body) if "::" in x]
    for cname, jname in (("plb_model_desc", "ModelDesc"), ("plb_run", "Run"), ("plb_opts", "Opts"), ("plb_summary", "Summary")):
        assert c_fields(cname) == jl_fields(jname), (cname, c_fields(cname), jl_fields(jname))
    assert [f.replace("η", "eta") for f in jl_fields("Bounds")] == c_fields("plb_bounds")


def test_compiled_variant_cache_key():
    """the built library is reused exactly when nothing that determines it changed (sources, flags, compiler)"""
    import petlion_b200
    from petlion_b200 import _lib
    path = petlion_b200.build()
    key_file = path + ".key"
    assert os.path.exists(key_file)
    key = open(key_file).read().strip()
    t0 = os.path.getmtime(path)
    assert petlion_b200.build() == path and os.path.getmtime(path) == t0          # cache hit: not rebuilt
    srcs = [os.path.join(_lib.CSRC, f) for f in os.listdir(_lib.CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "petlion_b200.h"))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    assert _lib._build_key(srcs, _lib.NVCC_FLAGS, nvcc) == key
    assert _lib._build_key(srcs, _lib.NVCC_FLAGS + ["-DX"], nvcc) != key
    assert _lib._build_key(srcs[1:], _lib.NVCC_FLAGS, nvcc) != key


def test_model_key_follows_the_reference_directory_hash():
    """strings_directory_func (src/external.jl:417-456): <Cathode>_<Anode>/sha1(every field of options_numerical but
    cathode and anode (outputs.jl:13-30), Np, Ns, Nn[, Na_Nz][, Nr])"""
    import hashlib
    from petlion_b200.api import _NS, model_key
    N = _NS(p=10, s=10, n=10, a=10, z=10, r_p=10, r_n=10)
    nm = _NS(temperature=False, solid_diffusion="Fickian", Fickian_method="finite_difference", aging=False, jacobian="symbolic",
             rxn_p="rxn_BV", rxn_n="rxn_BV")
    k = model_key(_NS(cathode="LCO", numerics=nm, N=N))
    assert k == "LCO_LiC6/" + hashlib.sha1(b"false_Fickian_finite_difference_false_rxn_BV_rxn_BV_OCV_LCO_OCV_LiC6_D_s_eff_rxn_rate_"
                                           b"D_eff_linear_K_eff_thermodynamic_factor_linear_symbolic_Np10_Ns10_Nn10__Nr_p10_Nr_n10").hexdigest()
    nm3 = nm.copy(); nm3.rxn_n = "rxn_MHC"
    assert model_key(_NS(cathode="LCO", numerics=nm3, N=N)) != k
    nm2 = nm.copy(); nm2.temperature = True
    assert model_key(_NS(cathode="LCO", numerics=nm2, N=N)) != k
    N2 = N.copy(); N2.a = 5
    assert model_key(_NS(cathode="LCO", numerics=nm, N=N2)) == k          # N_a only matters to thermal models
    assert model_key(_NS(cathode="NMC", numerics=nm, N=N)).startswith("NMC_LiC6_NMC/")


def test_every_family_unit_is_compiled_and_declared():
    """one translation unit per compiled family: the build list, the files in csrc/ and the variant declarations of the host
    (PLB_DECLARE_VARIANT + the VTAB rows plb_create looks a model's family up in) must name the same families"""
    import glob
    import re
    from petlion_b200 import _lib
    csrc = os.path.join(ROOT, "petlion.jl_b200", "csrc")
    files = {os.path.basename(f) for f in glob.glob(os.path.join(csrc, "plb_*.cu"))}
    assert files == set(_lib.UNITS)
    ns = set()
    for f in files - {"plb_kernels.cu"}:
        m = re.search(r"#define PLB_NS (\w+)", open(os.path.join(csrc, f)).read())
        assert m, f
        ns.add(m.group(1))
    assert len(ns) == len(files) - 1                                   # no two units share a namespace
    declared = set(re.findall(r"^PLB_DECLARE_VARIANT\((\w+)\)", open(os.path.join(csrc, "plb_common.cuh")).read(), re.M))
    assert declared == ns
    host = open(os.path.join(csrc, "plb_kernels.cu")).read()
    tables = set(re.findall(r"static const Variant V_\w+ = (?:PLB_VARIANT_TABLE\((\w+)\)|\{(\w+)::info)", host))
    assert {a or b for a, b in tables} == ns
    vtab = host[host.index("static const VEntry VTAB[]"):host.index("static const Variant* find_variant")]
    rows = re.findall(r"\{(\d), (\d), (\d), (\d), (\d), (\d+), (\d), &(V_\w+)\}", vtab)
    assert len(rows) == len(set(r[:7] for r in rows)) == len(set(r[7] for r in rows)) == 29      # one family per option set
    assert len(ns) == 29 + 2                                            # + the two concentration-rate-input siblings (isodc, widedc)
