"""Device-side samplers: Latin-hypercube stratification (pyDOE.lhs,
software.py:553,562) and inverse-CDF cell sampling (colloc2D_set, software.py:87-136)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d", [(1, 1), (7, 2), (10007, 2), (1 << 20, 3)])
def test_lhs_is_a_latin_hypercube(n, d):
    from pinn_based_online_pde_calculator_b200.engine import sample_lhs_device

    # narrow offset ranges lose strata to the fp32 coordinate grid at n = 2^20 (ulp(0.6) is 25 % of a
    # 2.4e-7 stratum), so the large case uses the unit cube
    lo, hi = ([0.0, 0.0, 0.0][:d], [1.0, 1.0, 1.0][:d]) if n > 100000 else ([-1.0, 0.5, 2.0][:d], [1.0, 0.75, 10.0][:d])
    x = sample_lhs_device(n, lo, hi, seed=1234).cpu().numpy().astype(np.float64)
    assert x.shape == (n, d)
    for j in range(d):
        t = (x[:, j] - lo[j]) / (hi[j] - lo[j])
        assert t.min() >= 0 and t.max() <= 1
        strata = np.clip(np.floor(t * n), 0, n - 1).astype(np.int64)
        # exactly one point per stratum in every dimension; fp32 rounding of (s+U)/n and of the affine
        # map can push a point that was drawn next to a stratum edge across it (measured 8 of 10007)
        counts = np.bincount(strata, minlength=n)
        # at n = 2^20 the fp32 coordinate itself (ulp ~ 5 % of a stratum) moves points across edges
        assert (counts != 1).sum() <= (0.12 if n > 100000 else 0.002) * n and counts.max() <= 2
    if n > 1000 and d >= 2:
        assert abs(np.corrcoef(x[:, 0], x[:, 1])[0, 1]) < 0.05      # independent permutations per dimension
    y = sample_lhs_device(n, lo, hi, seed=99).cpu().numpy()
    if n > 7:
        assert not np.array_equal(x.astype(np.float32), y)


def test_cdf2d_follows_the_cell_distribution():
    from pinn_based_online_pde_calculator_b200.engine import sample_cdf2d_device

    xs, ys = np.linspace(0.0, 2.0, 21), np.linspace(-1.0, 1.0, 11)
    X, Y = np.meshgrid(xs, ys)
    F = np.zeros_like(X)
    F[2:6, 5:15] = 1.0
    F[6:9, 0:4] = 3.0
    n = 400_000
    p = sample_cdf2d_device(n, X, Y, F, seed=7).cpu().numpy()
    ix = np.floor((p[:, 0] - 0.0) / 0.1 + 1e-6).astype(int)
    iy = np.floor((p[:, 1] + 1.0) / 0.2 + 1e-6).astype(int)
    H = np.zeros((10, 20))
    np.add.at(H, (np.clip(iy, 0, 9), np.clip(ix, 0, 19)), 1)
    Fc = F[:-1, :-1]
    assert H[Fc == 0].sum() <= 1e-5 * n                 # zero-mass cells receive no points (fp32 edge rounding aside)
    expect = Fc / Fc.sum() * n
    m = expect > 0
    assert np.abs(H[m] / expect[m] - 1).max() < 0.05
    assert p[:, 0].min() >= 0 and p[:, 0].max() <= 2 and p[:, 1].min() >= -1 and p[:, 1].max() <= 1


def test_run_pinn_training_with_device_sampler(tmp_path):
    from pinn_based_online_pde_calculator_b200.software import run_pinn_training
    from tests.test_gpu_training import KW

    res = run_pinn_training(**dict(KW, network_size={"depth": 32, "width": 3}), epochs={"adam": 250, "lbfgs": 30},
                            output_dir=str(tmp_path / "dev"), sampler="device", stage2=True)
    z = np.load(tmp_path / "dev" / "collocation_point_1.npz")
    assert z["X_col"].shape == (3000 + 1000 + 200 + 1000, 2)
    assert z["X_col"][:, 0].min() >= 0.1 - 1e-6 and z["X_col"][:, 0].max() <= 1 + 1e-6
    assert res["loss_1"][-1, 0] < 0.2 * res["loss_1"][0, 0] and np.isfinite(res["loss_2"]).all()
