"""ctypes binding of libpetlion_b200.so (C ABI: include/petlion_b200.h).

This is the Python stand-in for the `ccall` layer a Julia host would use.
There is NO CPU fallback: if the CUDA library is missing or no GPU is present, calls fail loudly.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("PLB_LIB", os.path.join(CSRC, "libpetlion_b200.so"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _build_key(srcs, flags, nvcc):
    """Compiled-variant cache key, the role of the reference's saved-model directory hash (src/external.jl:417-456:
    sha1 of the option string + discretisation, invalidated when the generator changes).  There the cache holds the
    Julia functions generated per model; here every model family is compiled ahead of time into ONE library, so the
    cached object is the library and the key is the sha1 of everything that determines it: the sources, the flags and
    the compiler version.  (mtimes do not survive a git checkout or a copy to the GPU box; the key does.)"""
    import hashlib
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    try:
        h.update(subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.split("release")[-1].encode())
    except Exception:
        h.update(b"nvcc?")
    for f in sorted(srcs):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


# one translation unit per compiled model family (plb_variant_*.cu) + the host ABI
UNITS = ("plb_kernels.cu", "plb_variant_iso.cu", "plb_variant_th.cu", "plb_variant_sei.cu",
         "plb_variant_wide.cu", "plb_variant_wsei.cu", "plb_variant_wth.cu", "plb_variant_thsei.cu", "plb_variant_wthsei.cu", "plb_variant_isodc.cu", "plb_variant_widedc.cu", "plb_variant_isomhc.cu", "plb_variant_thmhc.cu", "plb_variant_seimhc.cu", "plb_variant_isolgm.cu", "plb_variant_thlgm.cu",
         "plb_variant_iso_r12.cu", "plb_variant_th_r12.cu", "plb_variant_sei_r12.cu", "plb_variant_iso_r14.cu", "plb_variant_th_r14.cu", "plb_variant_sei_r14.cu",
         "plb_variant_iso_sp.cu", "plb_variant_th_sp.cu", "plb_variant_sei_sp.cu",
         "plb_variant_widemhc.cu", "plb_variant_wseimhc.cu", "plb_variant_wthmhc.cu", "plb_variant_thseimhc.cu", "plb_variant_wthseimhc.cu",
         "plb_variant_widelgm.cu", "plb_variant_wthlgm.cu")


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in UNITS + ("plb_common.cuh", "plb_variant.cuh", "plb_device.cuh",
                                                    "plb_integrator.cuh", "plb_tick.cuh", "laws_generated.cuh")]
    srcs.append(os.path.join(_HERE, "..", "include", "petlion_b200.h"))
    gen = os.path.join(CSRC, "laws_generated.cuh")
    if not os.path.exists(gen):
        subprocess.check_call([sys.executable, os.path.join(_HERE, "codegen", "gen_laws.py")])
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    key_path = LIB_PATH + ".key"
    key = _build_key(srcs, NVCC_FLAGS, nvcc) if os.path.exists(nvcc) else None
    if not force and os.path.exists(LIB_PATH):
        have = open(key_path).read().strip() if os.path.exists(key_path) else None
        # no compiler on this machine (a deployment box): the shipped library is what there is
        if key is None or have == key:
            return LIB_PATH
    # one translation unit per model family (isothermal / thermal) + the host ABI, compiled in parallel
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    objs, procs = [], []
    for u in UNITS:
        o = os.path.join(CSRC, u[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([nvcc] + flags + ["-c", "-o", o, os.path.join(CSRC, u)]))
    for p_ in procs:
        if p_.wait() != 0:
            raise subprocess.CalledProcessError(p_.returncode, p_.args)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-ldl", "-lpthread"])
    with open(key_path, "w") as fh:
        fh.write(key + "\n")
    return LIB_PATH


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("cathode", "N_p", "N_s", "N_n", "N_a", "N_z", "N_r_p", "N_r_n",
                                       "temperature", "aging", "device", "rxn_p", "rxn_n", "fickian_spectral")]


class Run(C.Structure):
    _fields_ = [("method", C.c_int), ("input_kind", C.c_int), ("value", C.c_double), ("tf", C.c_double),
                ("new_run", C.c_int), ("reserved", C.c_int)]


class Opts(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("abstol_init", C.c_double),
                ("reltol_init", C.c_double), ("maxiters", C.c_int), ("check_bounds", C.c_int),
                ("interp_final", C.c_int), ("skip_alg_deriv", C.c_int)]


class Bounds(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("V_max", "V_min", "SOC_max", "SOC_min", "T_max", "c_s_n_max",
                                          "I_max", "I_min", "eta_plating_min", "c_e_min", "dfilm_max")]


class Summary(C.Structure):
    _fields_ = [("t_end", C.c_double), ("V_end", C.c_double), ("I_end", C.c_double),
                ("SOC_end", C.c_double), ("T_end", C.c_double), ("aux_end", C.c_double), ("flag", C.c_int), ("n_steps", C.c_int), ("n_res", C.c_int),
                ("n_jac", C.c_int), ("n_netf", C.c_int), ("n_ncfn", C.c_int), ("n_newton_init", C.c_int),
                ("n_reinit", C.c_int)]


class InputTable(C.Structure):
    _fields_ = [("n", C.c_int), ("t", C.POINTER(C.c_double)), ("v", C.POINTER(C.c_double)),
                ("n_tdiscon", C.c_int), ("tdiscon", C.POINTER(C.c_double))]


SUMMARY_DTYPE = [("t_end", "f8"), ("V_end", "f8"), ("I_end", "f8"), ("SOC_end", "f8"), ("T_end", "f8"),
                 ("aux_end", "f8"), ("flag", "i4"),
                 ("n_steps", "i4"), ("n_res", "i4"), ("n_jac", "i4"), ("n_netf", "i4"), ("n_ncfn", "i4"),
                 ("n_newton_init", "i4"), ("n_reinit", "i4")]

EXPORTS = ["plb_last_error", "plb_create", "plb_destroy", "plb_set_stream", "plb_nstates", "plb_ndiff",
           "plb_ntheta", "plb_jac_nnz", "plb_theta_keys", "plb_theta_index", "plb_theta_defaults",
           "plb_bounds_defaults", "plb_opts_defaults", "plb_calc_I1C", "plb_jac_pattern",
           "plb_initial_guess", "plb_resjac", "plb_newton_init", "plb_linear_solve", "plb_simulate", "plb_simulate_table", "plb_set_tstops", "plb_set_dense_output", "plb_group_create", "plb_group_destroy", "plb_group_size",
           "plb_group_handle", "plb_group_simulate", "plb_group_device_summaries", "plb_group_last_gather_ms", "plb_variant_info", "plb_launch_count",
           "plb_last_kernel_ms"]

_lib = None


def lib():
    """Load the CUDA extension; raise (never fall back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " petlion.jl_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.c_void_p
        L.plb_last_error.restype = C.c_char_p
        L.plb_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
        L.plb_destroy.argtypes = [vp]
        L.plb_set_stream.argtypes = [vp, vp]
        for f in ("plb_nstates", "plb_ndiff", "plb_ntheta"):
            getattr(L, f).argtypes = [vp]
        L.plb_jac_nnz.argtypes = [vp, C.c_int]
        L.plb_theta_keys.argtypes = [vp, C.POINTER(C.c_char_p)]
        L.plb_theta_index.argtypes = [vp, C.c_char_p]
        L.plb_theta_defaults.argtypes = [vp, dp]
        L.plb_bounds_defaults.argtypes = [vp, C.POINTER(Bounds)]
        L.plb_opts_defaults.argtypes = [vp, C.POINTER(Opts)]
        L.plb_calc_I1C.argtypes = [vp, C.c_int, dp, dp]
        L.plb_jac_pattern.argtypes = [vp, C.c_int, ip, ip, C.c_int]
        L.plb_initial_guess.argtypes = [vp, C.c_int, dp, dp, dp, C.c_int]
        L.plb_resjac.argtypes = [vp, C.c_int, dp, dp, dp, dp, C.POINTER(Run), dp, dp, dp, C.c_int]
        L.plb_newton_init.argtypes = [vp, C.c_int, dp, dp, dp, C.POINTER(Run), dp, C.POINTER(Opts), vp, C.c_int]
        L.plb_linear_solve.argtypes = [vp, C.c_int, dp, dp, dp, dp, C.POINTER(Run), dp, dp, dp, vp, C.c_int]
        L.plb_simulate.argtypes = [vp, C.c_int, dp, C.POINTER(Run), dp, C.POINTER(Opts), C.POINTER(Bounds),
                                   dp, dp, dp, dp, dp, vp, C.c_int, dp, dp, dp, dp, dp, dp, vp, C.c_int]
        L.plb_simulate_table.argtypes = [vp, C.c_int, dp, C.POINTER(Run), C.POINTER(InputTable), dp, C.POINTER(Opts),
                                         C.POINTER(Bounds), dp, dp, dp, dp, dp, vp, C.c_int, dp, dp, dp, dp, dp, dp, vp,
                                         C.c_int]
        L.plb_set_tstops.argtypes = [vp, C.c_int, dp]
        L.plb_launch_count.argtypes = [vp]
        L.plb_launch_count.restype = C.c_longlong
        L.plb_last_kernel_ms.argtypes = [vp]
        L.plb_last_kernel_ms.restype = C.c_float
        if not hasattr(L, "plb_set_dense_output"):      # an older build named by PLB_LIB (A/B probes under profiles/)
            _lib = L
            return _lib
        L.plb_set_dense_output.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, vp, C.c_int]
        L.plb_group_create.argtypes = [C.POINTER(ModelDesc), C.c_int, ip, C.POINTER(vp)]
        L.plb_group_destroy.argtypes = [vp]
        L.plb_group_size.argtypes = [vp]
        L.plb_group_handle.argtypes = [vp, C.c_int]
        L.plb_group_handle.restype = vp
        L.plb_group_simulate.argtypes = [vp, C.c_int, dp, C.POINTER(Run), dp, C.POINTER(Opts), C.POINTER(Bounds),
                                         dp, dp, dp, dp, dp, vp, C.c_int, dp, dp, dp, dp, dp, vp]
        L.plb_group_device_summaries.argtypes = [vp, C.c_int, C.POINTER(vp), ip]
        L.plb_group_last_gather_ms.argtypes = [vp]
        L.plb_group_last_gather_ms.restype = C.c_float
        L.plb_launch_count.argtypes = [vp]
        L.plb_launch_count.restype = C.c_longlong
        L.plb_last_kernel_ms.argtypes = [vp]
        L.plb_last_kernel_ms.restype = C.c_float
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("petlion_b200: " + lib().plb_last_error().decode("utf-8", "replace"))
