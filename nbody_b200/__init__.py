"""nbody_b200 — host-side mirror (Python) of the reference's simulation interface
over the C ABI of the B200 FMM solver (include/nbody_cuda.h).

``CudaSimulation`` has the reference's shape: construct from bounds, particles and
a time step (OpenClSimulation ctor, include/nbody/open_cl_simulation.h:194-198),
``step()`` returns the new time and ``particles()`` returns the state in tree order
(include/nbody/simulation.h:33-34, src/open_cl_simulation.cpp:53-68). The C++ twin
is include/nbody/cuda_simulation.h. There is no CPU path here: importing works
anywhere, but constructing a simulation needs libnbody_cuda.so AND a B200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBODY_CUDA_LIB") or os.path.join(_HERE, "libnbody_cuda.so")  # override: an experimental build of the same library
PARTICLE_FLOATS = 12  # 48-byte AoS record: position[4], velocity[4], mass, charge, pad[2]

KICK_DRIFT, EXPLICIT_EULER = 0, 1
FLAG_KEEP_LISTS, FLAG_NO_INTEGRATE, FLAG_DIRECT, FLAG_CUB_SORT, FLAG_STATIC_PARTITION, FLAG_NO_OVERLAP, FLAG_DIST_SORT = 1, 2, 4, 8, 16, 32, 64
FLAG_PARTITIONED = 128

EXPORTED_SYMBOLS = [
    "nbody_cuda_default_config", "nbody_cuda_tuned_config", "nbody_cuda_create", "nbody_cuda_destroy", "nbody_cuda_set_particles", "nbody_cuda_step",
    "nbody_cuda_num_particles", "nbody_cuda_get_particles", "nbody_cuda_get_permutation", "nbody_cuda_get_accelerations",
    "nbody_cuda_get_keys", "nbody_cuda_get_tree", "nbody_cuda_get_lists", "nbody_cuda_get_expansions", "nbody_cuda_get_stats",
    "nbody_cuda_direct_field", "nbody_cuda_sort_runs", "nbody_cuda_comm_unique_id", "nbody_cuda_create_distributed", "nbody_cuda_owned_range",
    "nbody_cuda_get_owned_particles", "nbody_cuda_set_owned_particles", "nbody_cuda_rebalance",
    "nbody_cuda_set_time_step", "nbody_cuda_get_time_step", "nbody_cuda_next_time_step", "nbody_cuda_get_time",
    "nbody_cuda_checkpoint_save", "nbody_cuda_checkpoint_info", "nbody_cuda_checkpoint_read", "nbody_cuda_checkpoint_write",
    "nbody_cuda_checkpoint_load",
    "nbody_cuda_create_group", "nbody_cuda_group_step", "nbody_cuda_destroy_group",
    "nbody_cuda_last_error",
]
CHECKPOINT_MAGIC = 0x31504B435944424E  # the bytes "NBDYCKP1"


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("bounds", C.c_float * 4), ("time_step", C.c_float), ("force_constant", C.c_float),
                ("softening", C.c_float), ("mac_ratio", C.c_float), ("leaf_capacity", C.c_uint32), ("max_depth", C.c_uint32),
                ("order", C.c_uint32), ("integrator", C.c_uint32), ("flags", C.c_uint32), ("device", C.c_int32),
                ("pool_scale", C.c_float), ("low_order_tau", C.c_float), ("time_step_eta", C.c_float), ("time_step_min", C.c_float),
                ("time_step_max", C.c_float), ("partition_slack_pct", C.c_uint32), ("_reserved", C.c_uint32 * 2)]


class CheckpointHeader(C.Structure):
    """nbody_checkpoint_header (include/nbody_cuda.h): followed in the file by n*48 bytes of particles and n*4 of permutation."""
    _fields_ = [("magic", C.c_uint64), ("version", C.c_uint32), ("header_bytes", C.c_uint32), ("n_particles", C.c_uint64),
                ("steps_done", C.c_uint64), ("time", C.c_float), ("next_time_step", C.c_float), ("last_time_step", C.c_float),
                ("last_acc_max", C.c_float), ("checksum", C.c_uint64), ("config", Config)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_particles", "n_nodes", "n_leaves", "n_levels", "m2l_entries", "m2l_interactions",
                                          "m2l_interactions_low", "p2p_entries", "p2p_interactions", "near_entries", "retries", "device_bytes")] + \
               [(k, C.c_float) for k in ("ms_total", "ms_sort", "ms_tree", "ms_upsweep", "ms_traverse", "ms_m2l", "ms_l2l",
                                         "ms_leaf", "ms_comm", "work_imbalance")] + \
               [(k, C.c_uint64) for k in ("halo_particles", "imported_nodes", "migrated_particles")] + \
               [(k, C.c_float) for k in ("ms_import", "ms_halo", "ms_balance", "_pad")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class NbodyCudaError(RuntimeError):
    pass


_lib = None


def build_library(force=False, verbose=False):
    import importlib
    return importlib.import_module("nbody_b200.build").build(force=force, verbose=verbose)


def load_library():
    """Load libnbody_cuda.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NbodyCudaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.nbody_cuda_default_config.argtypes = [C.POINTER(Config)]
    L.nbody_cuda_default_config.restype = None
    L.nbody_cuda_tuned_config.argtypes = [C.POINTER(Config)]
    L.nbody_cuda_tuned_config.restype = None
    L.nbody_cuda_create.argtypes = [C.POINTER(Config), vp, u64, C.POINTER(vp)]
    L.nbody_cuda_destroy.argtypes = [vp]
    L.nbody_cuda_destroy.restype = None
    L.nbody_cuda_set_particles.argtypes = [vp, vp, u64]
    L.nbody_cuda_step.argtypes = [vp, C.POINTER(C.c_float)]
    L.nbody_cuda_num_particles.argtypes = [vp]
    L.nbody_cuda_num_particles.restype = u64
    L.nbody_cuda_get_particles.argtypes = [vp, vp, u64]
    L.nbody_cuda_get_permutation.argtypes = [vp, vp, u64]
    L.nbody_cuda_get_accelerations.argtypes = [vp, vp, u64]
    L.nbody_cuda_get_keys.argtypes = [vp, vp, u64]
    L.nbody_cuda_get_tree.argtypes = [vp, C.POINTER(u32), u32] + [vp] * 9
    L.nbody_cuda_get_lists.argtypes = [vp, C.POINTER(u64), vp, C.POINTER(u64), vp]
    L.nbody_cuda_get_expansions.argtypes = [vp, vp, vp, u64]
    L.nbody_cuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.nbody_cuda_direct_field.argtypes = [C.c_int, vp, u64, vp, u64, C.c_float, vp, C.POINTER(C.c_float), u32]
    L.nbody_cuda_sort_runs.argtypes = [C.c_int, vp, u64, vp, C.c_int, vp, vp]
    L.nbody_cuda_comm_unique_id.argtypes = [vp]
    L.nbody_cuda_create_distributed.argtypes = [C.POINTER(Config), vp, u64, u64, u64, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.nbody_cuda_owned_range.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.nbody_cuda_get_owned_particles.argtypes = [vp, vp, u64]
    L.nbody_cuda_set_owned_particles.argtypes = [vp, vp, u64]
    L.nbody_cuda_set_time_step.argtypes = [vp, C.c_float]
    L.nbody_cuda_get_time_step.argtypes = [vp] + [C.POINTER(C.c_float)] * 3
    L.nbody_cuda_next_time_step.argtypes = [C.POINTER(Config), C.c_float]
    L.nbody_cuda_next_time_step.restype = C.c_float
    L.nbody_cuda_get_time.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u64)]
    L.nbody_cuda_checkpoint_save.argtypes = [vp, C.c_char_p]
    L.nbody_cuda_checkpoint_info.argtypes = [C.c_char_p, C.POINTER(CheckpointHeader)]
    L.nbody_cuda_checkpoint_read.argtypes = [C.c_char_p, vp, vp, u64]
    L.nbody_cuda_checkpoint_write.argtypes = [C.c_char_p, C.POINTER(CheckpointHeader), vp, vp]
    L.nbody_cuda_checkpoint_load.argtypes = [C.c_char_p, C.POINTER(Config), C.POINTER(vp)]
    L.nbody_cuda_create_group.argtypes = [C.POINTER(Config), vp, u64, C.c_int, C.POINTER(vp)]
    L.nbody_cuda_group_step.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_float)]
    L.nbody_cuda_destroy_group.argtypes = [C.POINTER(vp), C.c_int]
    L.nbody_cuda_destroy_group.restype = None
    L.nbody_cuda_last_error.restype = C.c_char_p
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NbodyCudaError(f"nbody_cuda error {rc}: {load_library().nbody_cuda_last_error().decode()}")


def default_config(**overrides):
    cfg = Config()
    load_library().nbody_cuda_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if k == "bounds":
            for i, b in enumerate(list(v)[:4]):
                cfg.bounds[i] = b
        elif not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k}")
        else:
            setattr(cfg, k, v)
    return cfg


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def ncoef(order):
    return (order + 1) * (order + 2) * (order + 3) // 6


class CudaSimulation:
    """``Simulation<float, float4>`` on one B200 (or one rank of a multi-GPU run).

    particles: float32 [N, 12] boundary records (see nbody_b200.workloads)."""

    def __init__(self, bounds, particles, time_step, log=None, *, _distributed=None, _checkpoint=None, **config):
        self._h = C.c_void_p()
        self._lib = load_library()
        if _checkpoint is not None:  # see from_checkpoint()
            self._log = log
            hdr = checkpoint_info(_checkpoint)
            if config or bounds is not None or time_step is not None:
                base = {k: getattr(hdr.config, k) for k, _ in Config._fields_ if k not in ("_reserved", "bounds", "abi_version")}
                base["device"] = -1
                base.update(config)
                if time_step is not None:
                    base["time_step"] = time_step
                b = list(bounds) + [0.0] * (4 - len(bounds)) if bounds is not None else list(hdr.config.bounds)
                self.config = default_config(bounds=b, **base)
                cfgp = C.byref(self.config)
            else:
                self.config, cfgp = hdr.config, None
            _check(self._lib.nbody_cuda_checkpoint_load(os.fsencode(_checkpoint), cfgp, C.byref(self._h)))
            self.n = int(self._lib.nbody_cuda_num_particles(self._h))
            self.time = self.sim_time()[0]
            return
        particles = np.ascontiguousarray(particles, np.float32)
        if particles.ndim != 2 or particles.shape[1] != PARTICLE_FLOATS:
            raise ValueError("particles must be float32 [N, 12]")
        self.config = default_config(bounds=list(bounds) + [0.0] * (4 - len(bounds)), time_step=time_step, **config)
        self._log = log
        if _distributed is None:
            _check(self._lib.nbody_cuda_create(C.byref(self.config), _ptr(particles), particles.shape[0], C.byref(self._h)))
        else:
            import torch  # noqa: F401  (see comm_unique_id)
            d = _distributed
            uid = np.frombuffer(d["unique_id"], np.uint8).copy()
            _check(self._lib.nbody_cuda_create_distributed(C.byref(self.config), _ptr(particles), particles.shape[0],
                                                           d["n_global"], d["global_offset"], d["rank"], d["world"], _ptr(uid),
                                                           C.byref(self._h)))
        self.n = int(self._lib.nbody_cuda_num_particles(self._h))
        self.time = 0.0

    # --- the reference interface -------------------------------------------------
    def step(self):
        if self._log is not None:
            self._log.write(f"Starting a new step (t={self.time}).\n")
        t = C.c_float()
        _check(self._lib.nbody_cuda_step(self._h, C.byref(t)))
        self.time = t.value
        self.n = int(self._lib.nbody_cuda_num_particles(self._h))  # partitioned mode: a rank's particle count changes from step to step
        if self._log is not None:
            self._log.write("Step finished.\n")
        return t.value

    def particles(self, out=None):
        if out is None:
            out = np.empty((self.n, PARTICLE_FLOATS), np.float32)
        _check(self._lib.nbody_cuda_get_particles(self._h, _ptr(out), out.shape[0]))
        return out

    # --- variable time step and checkpoints -------------------
    @classmethod
    def from_checkpoint(cls, path, log=None, *, bounds=None, time_step=None, **config):
        """Continue the run stored in `path` (save_checkpoint). Without overrides the stored configuration is used."""
        return cls(bounds, None, time_step, log, _checkpoint=path, **config)

    def save_checkpoint(self, path):
        _check(self._lib.nbody_cuda_checkpoint_save(self._h, os.fsencode(path)))

    def set_time_step(self, dt):
        _check(self._lib.nbody_cuda_set_time_step(self._h, dt))

    def time_step(self):
        """{'next': dt of the next step(), 'last': dt of the last one, 'acc_max': max |a| of the last step (eta > 0)}"""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(self._lib.nbody_cuda_get_time_step(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"next": a.value, "last": b.value, "acc_max": c.value}

    def sim_time(self):
        t, k = C.c_float(), C.c_uint64()
        _check(self._lib.nbody_cuda_get_time(self._h, C.byref(t), C.byref(k)))
        return t.value, k.value

    # --- extras --------------------------------------------------------------------
    def set_particles(self, particles):
        particles = np.ascontiguousarray(particles, np.float32)
        _check(self._lib.nbody_cuda_set_particles(self._h, _ptr(particles), particles.shape[0]))

    def set_particles_ptr(self, ptr, n):
        _check(self._lib.nbody_cuda_set_particles(self._h, C.c_void_p(ptr), n))

    def particles_into_ptr(self, ptr, n):
        _check(self._lib.nbody_cuda_get_particles(self._h, C.c_void_p(ptr), n))

    def permutation(self):
        out = np.empty(self.n, np.uint32)
        _check(self._lib.nbody_cuda_get_permutation(self._h, _ptr(out), self.n))
        return out

    def accelerations(self):
        out = np.empty((self.n, 3), np.float32)
        _check(self._lib.nbody_cuda_get_accelerations(self._h, _ptr(out), self.n))
        return out

    def keys(self):
        out = np.empty(self.n, np.uint64)
        _check(self._lib.nbody_cuda_get_keys(self._h, _ptr(out), self.n))
        return out

    def tree(self):
        m = C.c_uint32()
        _check(self._lib.nbody_cuda_get_tree(self._h, C.byref(m), 0, *([None] * 9)))
        m = m.value
        t = {"depth": np.empty(m, np.uint32), "prefix": np.empty(m, np.uint64), "leaf_index": np.empty(m, np.uint32),
             "leaf_count": np.empty(m, np.uint32), "has_children": np.empty(m, np.uint8), "child_off": np.empty((m, 9), np.uint32),
             "parent_off": np.empty(m, np.int32), "sibling": np.empty(m, np.uint32), "geom": np.empty((m, 4), np.float32)}
        mm = C.c_uint32()
        _check(self._lib.nbody_cuda_get_tree(self._h, C.byref(mm), m, *[_ptr(t[k]) for k in
                                             ("depth", "prefix", "leaf_index", "leaf_count", "has_children", "child_off",
                                              "parent_off", "sibling", "geom")]))
        return t

    def lists(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._lib.nbody_cuda_get_lists(self._h, C.byref(a), None, C.byref(b), None))
        m2l = np.empty((a.value, 2), np.uint32)
        p2p = np.empty((b.value, 2), np.uint32)
        _check(self._lib.nbody_cuda_get_lists(self._h, C.byref(a), _ptr(m2l), C.byref(b), _ptr(p2p)))
        return m2l, p2p

    def expansions(self):
        m = C.c_uint32()
        _check(self._lib.nbody_cuda_get_tree(self._h, C.byref(m), 0, *([None] * 9)))
        nc = ncoef(self.config.order)
        M = np.empty((m.value, nc), np.float32)
        L = np.empty((m.value, nc), np.float32)
        _check(self._lib.nbody_cuda_get_expansions(self._h, _ptr(M), _ptr(L), M.size))
        return M, L

    def stats(self):
        st = Stats()
        _check(self._lib.nbody_cuda_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def owned_particles_into_ptr(self, ptr, capacity):
        _check(self._lib.nbody_cuda_get_owned_particles(self._h, C.c_void_p(ptr), capacity))

    def set_owned_particles_ptr(self, ptr, n):
        _check(self._lib.nbody_cuda_set_owned_particles(self._h, C.c_void_p(ptr), n))

    def owned_particles(self):
        first, count = self.owned_range()
        out = np.empty((count, PARTICLE_FLOATS), np.float32)
        _check(self._lib.nbody_cuda_get_owned_particles(self._h, _ptr(out), count))
        return out

    def owned_range(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._lib.nbody_cuda_owned_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.nbody_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Member(CudaSimulation):
    """One rank of a VirtualGroup: answers the per-rank calls (its own particles, in tree order)."""

    def __init__(self, lib, handle, config):
        self._lib, self._h, self.config, self._log = lib, C.c_void_p(handle), config, None
        self.n = int(lib.nbody_cuda_num_particles(self._h))
        self.time = 0.0

    def step(self):
        raise NbodyCudaError("a member of a virtual group steps with VirtualGroup.step()")

    def close(self):
        self._h = C.c_void_p()  # owned by the group


class VirtualGroup:
    """The partitioned multi-GPU scheme (FLAG_PARTITIONED: own particles + locally essential tree) with `world` ranks inside one
    process on ONE GPU (nbody_cuda_create_group): same kernels and phases, the exchanges are device copies. The parity tests use
    it to run 2, 4 and 8 ranks on a one-GPU box. particles(), permutation(), keys(), accelerations() concatenate the members in rank
    order, which is the global tree order."""

    def __init__(self, bounds, particles, time_step, world, **config):
        self._lib = load_library()
        particles = np.ascontiguousarray(particles, np.float32)
        if particles.ndim != 2 or particles.shape[1] != PARTICLE_FLOATS:
            raise ValueError("particles must be float32 [N, 12]")
        config["flags"] = int(config.get("flags", 0)) | FLAG_PARTITIONED
        self.config = default_config(bounds=list(bounds) + [0.0] * (4 - len(bounds)), time_step=time_step, **config)
        self.world = world
        self._hs = (C.c_void_p * world)()
        _check(self._lib.nbody_cuda_create_group(C.byref(self.config), _ptr(particles), particles.shape[0], world, self._hs))
        self.members = [_Member(self._lib, self._hs[r], self.config) for r in range(world)]
        self.n = particles.shape[0]
        self.time = 0.0

    def step(self):
        t = C.c_float()
        _check(self._lib.nbody_cuda_group_step(self._hs, self.world, C.byref(t)))
        self.time = t.value
        for m in self.members:
            m.n = int(self._lib.nbody_cuda_num_particles(m._h))
            m.time = t.value
        return t.value

    def counts(self):
        return [m.n for m in self.members]

    def particles(self):
        return np.concatenate([m.particles() for m in self.members])

    def permutation(self):
        return np.concatenate([m.permutation() for m in self.members])

    def keys(self):
        return np.concatenate([m.keys() for m in self.members])

    def accelerations(self):
        return np.concatenate([m.accelerations() for m in self.members])

    def stats(self):
        return [m.stats() for m in self.members]

    def close(self):
        if getattr(self, "_hs", None) is not None and self._hs[0]:
            self._lib.nbody_cuda_destroy_group(self._hs, self.world)
            self._hs = None
            for m in self.members:
                m._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def direct_field(src_posq, tgt_pos4, softening=0.01, device=-1, repeats=1):
    """All-pairs field of sources (x,y,z,q) on targets (x,y,z,*) with the tiled P2P kernel.
    Returns (field [n_tgt,3] float32, kernel milliseconds)."""
    L = load_library()
    src = np.ascontiguousarray(src_posq, np.float32)
    tgt = np.ascontiguousarray(tgt_pos4, np.float32)
    out = np.empty((tgt.shape[0], 3), np.float32)
    ms = C.c_float()
    _check(L.nbody_cuda_direct_field(device, _ptr(src), src.shape[0], _ptr(tgt), tgt.shape[0], softening, _ptr(out), C.byref(ms),
                                     repeats))
    return out, ms.value


def sort_runs(keys, bound, device=-1):
    """The device side of the distributed sort on one GPU (nbody_cuda_sort_runs): slices [bound[r], bound[r+1]) of `keys` are
    radix-sorted separately, then merged pairwise. Returns (sorted keys, input index per output position)."""
    L = load_library()
    k = np.ascontiguousarray(keys, np.uint64)
    b = np.ascontiguousarray(bound, np.uint32)
    ko, io = np.empty_like(k), np.empty(k.shape[0], np.uint32)
    _check(L.nbody_cuda_sort_runs(device, _ptr(k), k.shape[0], _ptr(b), b.shape[0] - 1, _ptr(ko), _ptr(io)))
    return ko, io


def next_time_step(acc_max, **config):
    """The variable-time-step rule on its own (host arithmetic): nbody_cuda_next_time_step."""
    cfg = config.pop("config", None) or default_config(**config)
    return float(load_library().nbody_cuda_next_time_step(C.byref(cfg), acc_max))


def checkpoint_info(path):
    hdr = CheckpointHeader()
    _check(load_library().nbody_cuda_checkpoint_info(os.fsencode(path), C.byref(hdr)))
    return hdr


def checkpoint_read(path):
    """(header, particles float32 [n,12], orig_index uint32 [n]) of a checkpoint file; host only, checksum verified."""
    hdr = checkpoint_info(path)
    n = int(hdr.n_particles)
    P = np.empty((n, PARTICLE_FLOATS), np.float32)
    orig = np.empty(n, np.uint32)
    _check(load_library().nbody_cuda_checkpoint_read(os.fsencode(path), _ptr(P), _ptr(orig), n))
    return hdr, P, orig


def checkpoint_write(path, particles, orig_index=None, *, time=0.0, steps_done=0, next_time_step=None, config=None):
    """Write a checkpoint from host arrays (host only): the way to start a run from an arbitrary saved state."""
    P = np.ascontiguousarray(particles, np.float32)
    if P.ndim != 2 or P.shape[1] != PARTICLE_FLOATS:
        raise ValueError("particles must be float32 [N, 12]")
    hdr = CheckpointHeader()
    hdr.config = config if config is not None else default_config()
    hdr.n_particles, hdr.steps_done, hdr.time = P.shape[0], steps_done, time
    hdr.next_time_step = hdr.config.time_step if next_time_step is None else next_time_step
    o = None if orig_index is None else np.ascontiguousarray(orig_index, np.uint32)
    if o is not None and o.shape != (P.shape[0],):
        raise ValueError("orig_index must be uint32 [N]")
    _check(load_library().nbody_cuda_checkpoint_write(os.fsencode(path), C.byref(hdr), _ptr(P), _ptr(o)))


def comm_unique_id():
    import torch  # noqa: F401  (loads PyTorch's bundled libnccl first so the process holds a single NCCL)
    uid = np.zeros(128, np.uint8)
    _check(load_library().nbody_cuda_comm_unique_id(_ptr(uid)))
    return uid.tobytes()
