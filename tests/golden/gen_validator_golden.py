"""Generate tests/golden/validator_golden.json by running the REFERENCE's own
validator (pinn_app/callbacks/input_validation.py:19-50) on a corpus of strings.

The function is nested in a Dash callback registration and dash is not installed,
so its source is lifted out with ``ast`` and compiled in isolation; nothing from
the reference is written into the repo except the (input, verdict) pairs.
Run here (the container with /root/reference); the GPU box only reads the JSON.
"""
import ast
import itertools
import json
import os
import random
import re
import string
import sys

REF = "/root/reference/pinn_app/callbacks/input_validation.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "validator_golden.json")

src = open(REF).read()
tree = ast.parse(src)
fn = None
for node in ast.walk(tree):
    if isinstance(node, ast.FunctionDef) and node.name == "on_equation_change":
        fn = node
fn.decorator_list = []
mod = ast.Module(body=[fn], type_ignores=[])
ns = {"re": re, "string": string}
exec(compile(ast.fix_missing_locations(mod), REF, "exec"), ns)
ref_validate = ns["on_equation_change"]

corpus = [
    "", "u_xx + 3*u_yy - 5", "u*u_x", "1/(r**2)*u_tt", "(x+y)*(x-y)", "u_x**2", "-u", "sin(x)", "1e-3", "((x))",
    "t", "z", "pi", "u_rr + 1/r*u_r + 1/(r**2)*u_tt", "u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", "u_y + u*u_x - 0.003183*u_xx",
    "x", "u", "r", "y", "u_x", "u_xy", "u_xyz", "u_", "_x", "xy", "2x", "x2", "x 2", "x + ", "+x", "x+", "x++y", "x**y",
    "x***y", "x*/y", "(x", "x)", "()", "(x)", "(x+y", "(x+(y))", "3.", ".5", ".", "1..2", "1.2.3", "3.14*u", "X", "u_X",
    "u_xx+", "*u", "u_xx u_yy", "u_xx  +\tu_yy", "x^2", "x,y", "u_1", "u_x1", "uu", "u_u", "u_uu", "r**2", "(r**2)",
    "1/(r**2)", "(u_xx+u_yy)*(x-y)/(x+y)", "(x)*(y)", "(1)", "(.5)", "(5.)", "u_ab", "u_zz", "a", "b", "u__x", "u_x_y",
    "0", "00", "1/0", "x/y/u", "x-y-u", "(x-y)-(u)", "x*(y", "x*(y))", "x(y)", "(x)(y)", "2(x)", "(x)2", "u_x(u)",
    "test equation", "u_xx + 2", "u_t - 0.1*(u_xx + u_yy)", "u_t", "e", "exp", "1e3", "1e", "x.y", "x.", ".x", "u.x",
]
rng = random.Random(1234)
alphabet = ["x", "y", "u", "r", "u_x", "u_yy", "u_rt", "t", "1", "2.5", ".5", "3.", "+", "-", "*", "**", "/", "(", ")",
            " ", "e", "_", "u_", "xx", "."]
for _ in range(1500):
    n = rng.randint(1, 9)
    corpus.append("".join(rng.choice(alphabet) for _ in range(n)))
# systematic: all token sequences up to length 4 over a small alphabet
small = ["x", "u_x", "2", "+", "**", "(", ")"]
for n in range(1, 5):
    for tup in itertools.product(small, repeat=n):
        corpus.append("".join(tup))
corpus = sorted(set(corpus))
out = [{"expr": s, "invalid": bool(ref_validate(s))} for s in corpus]
os.makedirs(os.path.dirname(OUT), exist_ok=True)
json.dump(out, open(OUT, "w"), indent=0)
print(len(out), "cases,", sum(o["invalid"] for o in out), "invalid ->", OUT)
