"""A synthetic 'world' for the orchestration pin of run_pinn_training: deterministic stand-ins for the trained
networks, the residual, the loss histories and the sampled points, shared by

  * tests/golden/gen_reference_training_golden.py -- which runs the REFERENCE'S OWN run_pinn_training body
    (pinn_app/software.py:626-1139) on them, and
  * tests/test_host_logic.py -- which runs the B200 driver's run_pinn_training on the same stand-ins

so that everything the two functions DERIVE (the 11 result files, the stage-2 network and loss hyper-parameters, the
epoch counts, the sampling sizes) can be compared number for number.  No training happens on either side."""
import numpy as np

# the reference's __main__ smoke arguments (software.py:1143-1190) with short epoch counts and a small test grid
KW = dict(
    equation="test equation",
    boundary={"bd_x1_min": 0.1, "bd_x1_max": 0.1, "bd_y1_min": 0.0, "bd_y1_max": 1.0, "bd_u1": 1.0,
              "bd_x2_min": 1.0, "bd_x2_max": 1.0, "bd_y2_min": 0.0, "bd_y2_max": 1.0, "bd_u2": 0.0},
    domain={"x_min": 0.1, "x_max": 1.0, "y_min": 0.0, "y_max": 1.0},
    scl=1.0, epsil=1.0,
    sample_points={"n_col": 5200, "n_bd": 1200, "n_add": 300},
    network_size={"depth": 60, "width": 6},
    testing_size={"x": 37, "y": 29},
    epochs={"adam": 40, "lbfgs": 18},
    equation_weight={"f": 0.05, "df": 0.0},
)
N_INFO = 6


def u_field(stage: int, version: int, z) -> np.ndarray:
    """'network output' [N, 1] of the stage's model after `version` optimiser legs"""
    z = np.asarray(z, dtype=np.float64)
    r, t = z[:, 0:1], z[:, 1:2]
    return np.log(r) / np.log(0.1) + (0.03 / stage) * np.exp(-0.7 * version) * np.sin(3.0 * t) * r


def residual_of(u, z) -> np.ndarray:
    """'gov_eqn' [N, 1]: any deterministic function of the output and the position"""
    z = np.asarray(z, dtype=np.float64)
    return 0.3 * np.sin(7.0 * np.asarray(u, dtype=np.float64)) + 0.1 * z[:, 0:1] * np.cos(2.0 * z[:, 1:2])


def loss_rows(stage: int, phase: str, n: int) -> list:
    """n loss_info rows [loss, loss_d, loss_e, bd_left, bd_right, eqn]"""
    off = {"adam": 0.0, "lbfgs": 0.5}[phase]
    rows = []
    for i in range(int(n)):
        base = (1.0 + off) / (stage * (1.0 + 0.1 * i))
        rows.append(np.array([base, 0.7 * base, 0.3 * base, 0.4 * base, 0.3 * base, 0.3 * base + 1e-3 * i]))
    return rows


def initial_loss(stage: int) -> float:
    return 3.5 + stage


def sampled_points(stage: int, call: int, n: int) -> np.ndarray:
    rs = np.random.RandomState(100 * stage + call)
    return rs.uniform(size=(int(n), 2)) * np.array([0.9, 1.0]) + np.array([0.1, 0.0])


def weight_map(version: int, shape) -> np.ndarray:
    return np.ones(shape) * (1.0 + 0.1 * version)
