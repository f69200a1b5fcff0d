#!/usr/bin/env python3
"""Extract the reference's example problem data into tinympc-matlab_b200/problem_data.json.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference); the JSON it writes
is committed so that tests / bench.py never read /root/reference at run time.

Sources (data only, no code):
  quadrotor : tinympc/TinyMPC/examples/problem_data/quadrotor_20hz_params.hpp:5-37
              (row-major tables, mapped RowMajor in quadrotor_hovering.cpp:35-39)
  rocket    : examples/rocket_landing_constraints.m:17-46 (same numbers as
              tinympc/TinyMPC/examples/problem_data/rocket_landing_params_20hz.hpp:5-29)
  cartpole  : examples/cartpole_example_one_solve.m:13-20
"""
import json, re, sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tinympc-matlab_b200" / "problem_data.json"


def table(text, name):
    m = re.search(name + r"\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}", text, re.S)
    return [float(t.rstrip("f")) for t in re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?f?", m.group(1))]


def rowmajor(vals, r, c):
    return [[vals[i * c + j] for j in range(c)] for i in range(r)]


q = (REF / "tinympc/TinyMPC/examples/problem_data/quadrotor_20hz_params.hpp").read_text()
quad = dict(
    nx=12, nu=4, rho=float(re.search(r"rho_value\s*=\s*([\d.]+)", q).group(1)),
    A=rowmajor(table(q, "Adyn_data"), 12, 12), B=rowmajor(table(q, "Bdyn_data"), 12, 4),
    f=[0.0] * 12, Q=table(q, "Q_data"), R=table(q, "R_data"),
)

r = (REF / "tinympc/TinyMPC/examples/problem_data/rocket_landing_params_20hz.hpp").read_text()
rocket = dict(
    nx=6, nu=3, rho=float(re.search(r"rho_value\s*=\s*([\d.]+)", r).group(1)),
    A=rowmajor(table(r, "Adyn_data"), 6, 6), B=rowmajor(table(r, "Bdyn_data"), 6, 3),
    f=table(r, "fdyn_data"), Q=table(r, "Q_data"), R=table(r, "R_data"),
)

cart = dict(
    nx=4, nu=1, rho=1.0,
    A=[[1.0, 0.01, 0.0, 0.0], [0.0, 1.0, 0.039, 0.0], [0.0, 0.0, 1.002, 0.01], [0.0, 0.0, 0.458, 1.002]],
    B=[[0.0], [0.02], [0.0], [0.067]], f=[0.0] * 4, Q=[10.0, 1.0, 10.0, 1.0], R=[1.0],
)
# sanity: the .m file carries the same cartpole numbers
m = (REF / "examples/cartpole_example_one_solve.m").read_text()
assert "0.458, 1.002" in m and "0.067" in m and "diag([10.0, 1, 10, 1])" in m

OUT.write_text(json.dumps(dict(quadrotor=quad, rocket=rocket, cartpole=cart), indent=1))
print("wrote", OUT)
