"""Runs bench.main() (measured arm, one rank) with the device library replaced by a recording stub: see
tests/test_bench_line_assembly.py. Checks bookkeeping only; measures nothing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self: self
import nbody_b200


class StubConfig:
    max_depth = 21


class StubSimulation:
    """Records nothing but the particle array it was given; stats() returns fixed stage times of a 12-level tree."""

    def __init__(self, bounds, particles, dt, **kw):
        self.P = np.asarray(particles, np.float32)
        self.config = StubConfig()

    def step(self):
        return 0.0

    def stats(self):
        return dict(n_particles=self.P.shape[0], n_nodes=100, n_leaves=50, n_levels=12, m2l_entries=10, m2l_interactions=1000,
                    m2l_interactions_low=600, p2p_entries=10, p2p_interactions=100000, near_entries=5, retries=0, device_bytes=1 << 20,
                    ms_total=1.0, ms_sort=0.1, ms_tree=0.05, ms_upsweep=0.05, ms_traverse=0.1, ms_m2l=0.2, ms_l2l=0.05, ms_leaf=0.4, ms_comm=0.0,
                    work_imbalance=0.0, halo_particles=0, imported_nodes=0, migrated_particles=0, ms_import=0.0, ms_halo=0.0, ms_balance=0.0)

    def owned_range(self):
        return 0, self.P.shape[0]

    def owned_particles_into_ptr(self, ptr, capacity):
        pass

    def set_owned_particles_ptr(self, ptr, n):
        pass

    def particles(self):
        return self.P

    def accelerations(self):
        # the direct-field stub below returns ones, so the accuracy check of the line sees "zero error" for unit charge / mass ratios
        return np.ones((self.P.shape[0], 3), np.float32) * (self.P[:, 9] / self.P[:, 8])[:, None]

    def close(self):
        pass


nbody_b200.CudaSimulation = StubSimulation
nbody_b200.direct_field = lambda src, tgt, eps, device=0, repeats=1: (np.ones((len(tgt), 3), np.float64), 1.0)
import bench

sys.argv = ["bench.py", "--n", "65536", "--steps", "2", "--warmup", "1", "--cpu-sample", "256", "--cpu-steps", "1", "--accuracy-targets", "64"]
bench.main()
