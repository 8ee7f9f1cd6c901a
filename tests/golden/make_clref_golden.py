"""Generates tests/golden/clref_golden.npz from the REFERENCE'S OWN OpenCL C kernels, compiled for the host from the files
where they lie (oracle/_ref/libclref.so = /root/reference/src/{interaction,field,verify,moment,force}.cl behind
oracle/shim_cl/opencl_c_host.h and oracle/ref_cl_harness.cpp; recipe: oracle/Makefile, target clref).

Per case (the particle sets of tests/golden/fmm_path.npz, so that the GPU parity fixtures are tied to these vectors):
  node_pairs / leaf_pairs   the reference's node (M2L) and leaf (P2P) interaction lists in the order its host loop appends
                            them (src/open_cl_simulation.cpp:242-266), produced by find_interactions (src/interaction.cl:10-100)
  rounds                    launches of find_interactions
  charge, dipole, qcross, qtrace   node_moment_t of every node after compute_moments_from_leafs / _from_nodes (src/moment.cl)
  leaf_force, node_force    the two per-particle force arrays the reference's integration adds (capacity <= 8 only: the near-field
                            kernel reaches at most 8 leaves of a node, src/field.cl:87-102), defects D5 / D7 repaired in the harness
  leaf_force_as_written, node_force_as_written   the same with D5 / D7 left in and a work-group size of 32
plus the nine struct sizes verify.cl reports and pair-force known answers from leaf_moment_field + leaf_field_to_force.

The octree the kernels run on is the oracle's (glade::Orthtree is not in the reference tree); everything else is the reference.
Run in the authoring container (needs /root/reference): python tests/golden/make_clref_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle
from nbody_b200 import workloads
from conftest import sorted_system

CASES = [("uniform", 700, 8), ("plummer", 900, 8), ("two_galaxies", 800, 4), ("plummer", 1500, 32)]


def pair_kats():
    rng = np.random.default_rng(11)
    rows = []
    for _ in range(64):
        pa, pb = rng.random(3).astype(np.float32), rng.random(3).astype(np.float32)
        if rng.random() < 0.25:
            pb = (pa + np.float32(1e-3) * rng.standard_normal(3)).astype(np.float32)  # inside the softening length
        qa, qb = np.float32(0.1 + rng.random()), np.float32(0.1 + rng.random())
        fa, fb = oracle.clref_pair_force(qa, qb, pa, pb)
        rows.append(np.concatenate([[qa, qb], pa, pb, fa, fb]).astype(np.float32))
    return np.stack(rows)


def build_case(kind, n, cap):
    S = sorted_system(workloads.GENERATORS[kind](n), capacity=cap)
    R = oracle.ClRef(S["tree"], S["P"])
    node, leaf, rounds = R.traverse()
    out = {"node_pairs": node, "leaf_pairs": leaf, "rounds": np.uint64(rounds), "geometry": R.geometry()}
    if cap <= 8:
        R0 = oracle.ClRef(S["tree"], S["P"])
        R0.traverse()
        launches0 = R0.moments(repair_d5=False)[0]
        lf0, nf0 = R0.forces(repair_d7=False, node_local_size=32)
        out.update(leaf_force_as_written=lf0, node_force_as_written=nf0, upsweep_launches_as_written=np.uint32(launches0))
    launches, q, d, c, t = R.moments(repair_d5=True)
    out.update(charge=q, dipole=d, qcross=c, qtrace=t, upsweep_launches=np.uint32(launches))
    if cap <= 8:
        lf, nf = R.forces(repair_d7=True)
        out.update(leaf_force=lf, node_force=nf)
    return out


def main():
    if oracle.clref_lib() is None:
        raise SystemExit("oracle/_ref/libclref.so is not built (needs /root/reference)")
    sizes = oracle.clref_type_sizes()
    blob = {"cases": np.array([f"{k}:{n}:{c}" for k, n, c in CASES]), "type_names": np.array(list(sizes.keys())),
            "type_sizes": np.array(list(sizes.values()), np.uint32), "pair_kats": pair_kats()}
    for k, n, c in CASES:
        for name, arr in build_case(k, n, c).items():
            blob[f"{k}_{n}_{c}/{name}"] = arr
    path = os.path.join(HERE, "clref_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
