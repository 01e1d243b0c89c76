"""B200-native engine for the residual-loss + gradient hot path of
Cc1-Yy/PINN-based-online-PDE-calculator.

Public surface (mirrors pinn_app/software.py of the reference):
    run_pinn_training, adam_optimizer, lbfgs_optimizer, predictF, data_func_create
    PinnEngine / NetworkSpec / compile_equation for the operator-level API.
"""
from .equation import CompiledEquation, EquationError, compile_equation, validate_reference  # noqa: F401
from .engine import NetworkSpec, PinnEngine, load_library, shard_range  # noqa: F401

__all__ = ["CompiledEquation", "EquationError", "compile_equation", "validate_reference",
           "NetworkSpec", "PinnEngine", "load_library", "shard_range"]
