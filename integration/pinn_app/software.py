"""Shim a maintainer would drop over the reference's pinn_app/software.py so that
`from pinn_app.software import run_pinn_training` (callbacks/training.py:6) resolves to the
B200 engine.  See INTEGRATION.md."""
from pinn_based_online_pde_calculator_b200.software import (  # noqa: F401
    adam_optimizer, colloc2D_set, data_func_create, gaussian2D_smooth, lbfgs_optimizer, predictF, run_pinn_training)
