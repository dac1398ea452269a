"""Golden vectors for the reference's tests/test_optimise.py::test_valid_numbers (build container only).

    python tests/golden/make_golden_tnc.py

The phoney 49-point spectrum (S, k) that test hard-codes (tests/test_optimise.py:71-177) is read out of the
reference's test module by evaluating its two `np.array([...])` literals, and the UNMODIFIED reference is run on
it: `min_func` at the test's start point, `minimize(min_func, method="TNC")` exactly as the test calls it
(tests/test_optimise.py:197-203), and the default L-BFGS-B call of `optimise` (SURVEY.md section 8c quotes its
result).  Only data and numerical outputs are stored (tests/golden/tnc.npz).
"""
import ast
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYCURIOUS_REF", "/root/reference")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
sys.modules.setdefault("pycurious.radon", types.ModuleType("pycurious.radon"))
import pycurious                                    # noqa: E402
from scipy.optimize import minimize                 # noqa: E402

src = open(os.path.join(REF, "tests", "test_optimise.py")).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "test_valid_numbers")
vals = {}
for node in ast.walk(fn):
    if isinstance(node, ast.Assign) and len(node.targets) == 1 and getattr(node.targets[0], "id", None) in ("S", "k"):
        vals[node.targets[0].id] = np.array(ast.literal_eval(node.value.args[0]), dtype=np.float64)
S, k = vals["S"], vals["k"]
assert S.size == k.size == 49

raw = np.loadtxt(os.path.join(REF, "tests", "test_mag_data.txt"))
d = raw[:, 2].reshape(305, 305)
grid = pycurious.CurieOptimise(d, raw[:, 0].min(), raw[:, 0].max(), raw[:, 1].min(), raw[:, 1].max())
x0 = np.array([3.0, 1.0, 40.0, 5.0])
sigma_S = np.ones_like(S)
f0 = grid.min_func(x0, k, S, sigma_S)
lower = np.zeros_like(x0)
res_tnc = minimize(grid.min_func, x0, args=(k, S, sigma_S), method="TNC", bounds=list(zip(lower, [None] * 4)))
res_lb = minimize(grid.min_func, x0, args=(k, S, sigma_S), bounds=grid.bounds)
np.savez_compressed(os.path.join(HERE, "tnc.npz"), S=S, k=k, x0=x0, f0=f0, x_tnc=res_tnc.x, f_tnc=res_tnc.fun,
                    x_lbfgsb=res_lb.x, f_lbfgsb=res_lb.fun, nfev_lbfgsb=res_lb.nfev)
print("f0", f0, "TNC", res_tnc.x, res_tnc.fun, "L-BFGS-B", res_lb.x, res_lb.fun, res_lb.nfev)
