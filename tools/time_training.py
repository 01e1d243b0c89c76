"""Wall-clock breakdown of run_pinn_training on the reference's own problem size (needs a B200).
usage: python tools/time_training.py [adam_epochs] [lbfgs_epochs]"""
import collections
import sys
import tempfile
import time

sys.path.insert(0, ".")
from pinn_based_online_pde_calculator_b200 import engine as E  # noqa: E402
from pinn_based_online_pde_calculator_b200.software import run_pinn_training  # noqa: E402
from tests.test_gpu_training import KW  # noqa: E402  (the reference's __main__ smoke arguments, sw:1143-1190)

adam = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
lbfgs = int(sys.argv[2]) if len(sys.argv) > 2 else 500
stats = collections.defaultdict(lambda: [0, 0.0, 0.0, 0])
orig = E.PinnEngine.adam_steps


def timed(self, n_steps, lr, want_rows=True):
    t0 = time.perf_counter()
    r = orig(self, n_steps, lr, want_rows)
    s = stats[id(self)]
    s[0] += n_steps
    s[1] += time.perf_counter() - t0
    s[2] += self.last_ms() * 1e-3
    s[3] += 1
    return r


E.PinnEngine.adam_steps = timed
orig_begin, orig_end = E.PinnEngine.adam_steps_begin, E.PinnEngine.adam_steps_end
t_begin = {}


def timed_begin(self, n_steps, lr):
    t_begin[id(self)] = time.perf_counter()
    return orig_begin(self, n_steps, lr)


def timed_end(self, n_steps):
    r = orig_end(self, n_steps)
    s = stats[id(self)]
    s[0] += n_steps
    s[1] += time.perf_counter() - t_begin[id(self)]   # wall from the enqueue to the rows (host sampling overlaps inside)
    s[2] += self.last_ms() * 1e-3
    s[3] += 1
    return r


E.PinnEngine.adam_steps_begin, E.PinnEngine.adam_steps_end = timed_begin, timed_end
with tempfile.TemporaryDirectory() as d:
    run_pinn_training(**KW, epochs={"adam": 10, "lbfgs": 5}, output_dir=d + "/warm")  # library load, first launches
    stats.clear()
    t0 = time.perf_counter()
    run_pinn_training(**KW, epochs={"adam": adam, "lbfgs": lbfgs}, output_dir=d + "/run")
    dt = time.perf_counter() - t0
print(f"run_pinn_training(adam={adam}, lbfgs={lbfgs}): {dt:.2f} s wall", file=sys.stderr)
for i, (k, (n, wall, gpu, calls)) in enumerate(stats.items()):
    print(f"  engine {i}: {n} Adam steps in {calls} calls: wall {wall:.3f} s ({1e3 * wall / n:.3f} ms/step), GPU {gpu:.3f} s ({1e3 * gpu / n:.3f} ms/step)",
          file=sys.stderr)
