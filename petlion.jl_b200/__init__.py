"""petlion.jl_b200 -- B200-native batched implicit DAE integrator behind PETLION.jl's
petlion()/simulate()/simulate!() API and its residual/Jacobian callback surface.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI), codegen/ (sympy -> CUDA
constitutive laws) and the host-side mirror of the reference interface (api.py).
"""
from . import _lib, sweep
from ._lib import build
from .api import EXIT_REASONS, Model, Solution, Table, model_key, petlion, rxn_BV, rxn_MHC, simulate, simulate_

LCO = "LCO"
NMC = "NMC"
simulate_bang = simulate_   # Julia's simulate!

__all__ = ["petlion", "simulate", "simulate_", "simulate_bang", "LCO", "NMC", "Model", "Solution", "Table", "build",
           "EXIT_REASONS", "rxn_BV", "rxn_MHC", "model_key", "sweep"]
