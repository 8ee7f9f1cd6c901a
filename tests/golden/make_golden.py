"""Generates tests/golden/naive_kat.json from the UNMODIFIED reference naive path
(oracle/_ref/libnaive_ref.so = /root/reference/src/naive_simulation.cpp compiled by
oracle/Makefile). Run in the authoring container: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle
from nbody_b200 import workloads


def case(name, P, k, dt, steps):
    out, t = oracle.ref_naive_run(P, k, dt, steps)
    return {"name": name, "force_constant": k, "dt": dt, "steps": steps, "time": float(t),
            "particles_in": np.asarray(P, np.float32).tolist(), "particles_out": out.tolist()}


def main():
    cases = []
    P = np.zeros((2, 12), np.float32)
    P[0, 8] = 1.0; P[0, 9] = 1.0
    P[1, 0:3] = (1.0, 0.5, 0.25); P[1, 8] = 2.0; P[1, 9] = 1.0
    cases.append(case("survey_kat_two_particles_repulsive", P, -1.0, 0.001, 1))
    cases.append(case("two_particles_attractive_3_steps", P, 1.0, 0.001, 3))
    cases.append(case("uniform_cube_32_seed42_2_steps", workloads.uniform_cube(32), 1.0, 0.001, 2))
    L = np.zeros((5, 12), np.float32)
    L[:, 0] = [0.1, 0.3, 0.45, 0.7, 0.95]; L[:, 8] = [1, 2, 3, 4, 5]; L[:, 9] = [0.5, 0.4, 0.3, 0.2, 0.1]
    cases.append(case("line_of_5", L, 1.0, 0.001, 4))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "naive_kat.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (unmodified reference naive_simulation.cpp)",
                   "cases": cases}, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
