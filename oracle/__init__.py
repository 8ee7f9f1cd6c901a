"""ctypes front end of the CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this package (see oracle/oracle.cpp header).
``liboracle.so`` is the restatement; ``_ref/libnaive_ref.so`` is the UNMODIFIED
reference naive path (/root/reference/src/naive_simulation.cpp) and
``_ref/libclref.so`` the reference's OpenCL C device kernels compiled for the host
(oracle/ref_cl_harness.cpp), both built by oracle/Makefile from the sources where
they lie.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None
_CLREF = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


def build():
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_morton_keys.argtypes = [C.c_uint64, f32p, C.c_uint32, f32p, u64p]
        L.orc_sort_keys.argtypes = [C.c_uint64, u64p, u64p, u32p]
        L.orc_tree_build.argtypes = [C.c_uint64, u64p, f32p, C.c_uint32, C.c_uint32]
        L.orc_tree_build.restype = C.c_void_p
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_tree_num_nodes.argtypes = [C.c_void_p]
        L.orc_tree_num_nodes.restype = C.c_uint32
        L.orc_tree_get.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        L.orc_traverse.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_get_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_traverse_rounds.argtypes = [C.c_void_p]
        L.orc_traverse_rounds.restype = C.c_uint64
        L.orc_direct_field.argtypes = [C.c_uint64, f32p, C.c_uint64, C.c_void_p, C.c_double, f64p, C.c_void_p, C.c_int]
        L.orc_ncoef.argtypes = [C.c_uint32]
        L.orc_ncoef.restype = C.c_uint32
        L.orc_fmm_field.argtypes = [C.c_void_p, C.c_uint64, f32p, C.c_uint32, C.c_double, f64p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int]
        L.orc_set_low_order_tau.argtypes = [C.c_double]
        L.orc_low_count.restype = C.c_uint64
        L.orc_all_count.restype = C.c_uint64
        L.orc_naive_step_as_written.argtypes = [C.c_uint64, f32p, C.c_float, C.c_float, C.c_uint32]
        L.orc_naive_step_as_written.restype = C.c_float
        L.orc_direct_step.argtypes = [C.c_uint64, f32p, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_int, C.c_int]
        L.orc_direct_step.restype = C.c_float
        L.orc_next_time_step.argtypes = [C.c_float] * 6
        L.orc_next_time_step.restype = C.c_float
        L.orc_direct_step_adaptive.argtypes = [C.c_uint64, f32p] + [C.c_float] * 7 + [C.c_uint32, C.c_int, C.c_int, f32p, f32p]
        L.orc_direct_step_adaptive.restype = C.c_float
        _LIB = L
    return _LIB


def ref_lib():
    """The unmodified reference naive path, or None when it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libnaive_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src"):
                build()
            else:
                return None
        R = C.CDLL(path)
        R.ref_naive_run.argtypes = [C.c_uint64, f32p, C.c_float, C.c_float, C.c_uint32]
        R.ref_naive_run.restype = C.c_float
        R.ref_particle_size.restype = C.c_uint32
        _REF = R
    return _REF


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def morton_keys(pos, bounds):
    pos = np.ascontiguousarray(pos, np.float32)
    keys = np.empty(pos.shape[0], np.uint64)
    lib().orc_morton_keys(pos.shape[0], pos, pos.shape[1], np.asarray(bounds, np.float32)[:3].copy(), keys)
    return keys


def sort_keys(keys):
    keys = np.ascontiguousarray(keys, np.uint64)
    out = np.empty_like(keys)
    perm = np.empty(keys.shape[0], np.uint32)
    lib().orc_sort_keys(keys.shape[0], keys, out, perm)
    return out, perm


class Tree:
    """Reference-contract octree in DFS pre-order."""

    def __init__(self, sorted_keys, bounds, capacity=8, max_depth=21):
        sorted_keys = np.ascontiguousarray(sorted_keys, np.uint64)
        self._h = lib().orc_tree_build(sorted_keys.shape[0], sorted_keys, np.asarray(bounds, np.float32)[:3].copy(),
                                       capacity, max_depth)
        m = self.num_nodes = lib().orc_tree_num_nodes(self._h)
        self.depth = np.empty(m, np.uint32)
        self.prefix = np.empty(m, np.uint64)
        self.leaf_index = np.empty(m, np.uint32)
        self.leaf_count = np.empty(m, np.uint32)
        self.has_children = np.empty(m, np.uint8)
        self.child_off = np.empty((m, 9), np.uint32)
        self.parent_off = np.empty(m, np.int32)
        self.sibling = np.empty(m, np.uint32)
        self.geom = np.empty((m, 4), np.float32)
        lib().orc_tree_get(self._h, *[_ptr(a) for a in (self.depth, self.prefix, self.leaf_index, self.leaf_count,
                                                         self.has_children, self.child_off, self.parent_off,
                                                         self.sibling, self.geom)])
        self.m2l = self.p2p = None

    def traverse(self, mac_ratio=0.5):
        a, b = C.c_uint64(), C.c_uint64()
        lib().orc_traverse(self._h, mac_ratio, C.byref(a), C.byref(b))
        self.m2l = np.empty((a.value, 2), np.uint32)
        self.p2p = np.empty((b.value, 2), np.uint32)
        lib().orc_get_lists(self._h, _ptr(self.m2l), _ptr(self.p2p))
        self.rounds = lib().orc_traverse_rounds(self._h)
        return self.m2l, self.p2p

    def fmm_field(self, posq_sorted, order, eps, want_phi=False, want_expansions=False, low_order_tau=0.0):
        """FP64 FMM over the traversal's lists. low_order_tau > 0 mirrors the product's adaptive-order
        M2L (pairs with ext2 < tau*d2 at order-1); self.low_fraction reports how many pairs that was."""
        lib().orc_set_low_order_tau(float(np.float32(low_order_tau)))
        posq = np.ascontiguousarray(posq_sorted, np.float32)
        n = posq.shape[0]
        g = np.empty((n, 3), np.float64)
        phi = np.empty(n, np.float64) if want_phi else None
        nc = lib().orc_ncoef(order)
        M = np.empty((self.num_nodes, nc), np.float64) if want_expansions else None
        L = np.empty((self.num_nodes, nc), np.float64) if want_expansions else None
        lib().orc_fmm_field(self._h, n, posq, order, eps, g, _ptr(phi), _ptr(M), _ptr(L), 1)
        self.low_count = lib().orc_low_count()
        self.low_fraction = self.low_count / max(1, lib().orc_all_count())
        lib().orc_set_low_order_tau(0.0)
        out = [g]
        if want_phi:
            out.append(phi)
        if want_expansions:
            out += [M, L]
        return out[0] if len(out) == 1 else tuple(out)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tree_free(self._h)
            self._h = None


def direct_field(posq, targets=None, eps=0.01, want_phi=False, threads=None):
    posq = np.ascontiguousarray(posq, np.float32)
    if targets is not None:
        targets = np.ascontiguousarray(targets, np.uint32)
    nt = posq.shape[0] if targets is None else targets.shape[0]
    g = np.empty((nt, 3), np.float64)
    phi = np.empty(nt, np.float64) if want_phi else None
    lib().orc_direct_field(posq.shape[0], posq, nt, _ptr(targets), eps, g, _ptr(phi), threads or os.cpu_count() or 1)
    return (g, phi) if want_phi else g


def naive_step_as_written(particles12, force_constant, dt, steps=1):
    P = np.ascontiguousarray(particles12, np.float32).copy()
    t = lib().orc_naive_step_as_written(P.shape[0], P, force_constant, dt, steps)
    return P, t


def direct_step(particles12, G=1.0, eps=0.01, dt=1e-3, steps=1, integrator=0, threads=None):
    P = np.ascontiguousarray(particles12, np.float32).copy()
    t = lib().orc_direct_step(P.shape[0], P, G, eps, dt, steps, integrator, threads or os.cpu_count() or 1)
    return P, t


def next_time_step(eta, length, acc_max, dt0, dt_min=0.0, dt_max=0.0):
    return float(lib().orc_next_time_step(eta, length, acc_max, dt0, dt_min, dt_max))


def direct_step_adaptive(particles12, G=1.0, eps=0.01, dt0=1e-3, eta=0.1, dt_min=0.0, dt_max=0.0, steps=1, integrator=0, length=None,
                         threads=None):
    """Variable-step FP64 direct-sum run: returns (particles, final time, dt per step, max |a| per step)."""
    P = np.ascontiguousarray(particles12, np.float32).copy()
    dts = np.zeros(steps, np.float32)
    amax = np.zeros(steps, np.float32)
    t = lib().orc_direct_step_adaptive(P.shape[0], P, G, eps, eps if length is None else length, dt0, eta, dt_min, dt_max, steps,
                                       integrator, threads or os.cpu_count() or 1, dts, amax)
    return P, t, dts, amax


def ref_naive_run(particles12, force_constant, dt, steps=1):
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libnaive_ref.so is not built")
    P = np.ascontiguousarray(particles12, np.float32).copy()
    t = R.ref_naive_run(P.shape[0], P, force_constant, dt, steps)
    return P, t


# ---- the reference's OpenCL C kernels on the host (oracle/_ref/libclref.so) ----------------------------------------------
def clref_lib():
    """The reference's device kernels compiled for the host (oracle/ref_cl_harness.cpp), or None when never built."""
    global _CLREF
    if _CLREF is None:
        path = os.path.join(_HERE, "_ref", "libclref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src"):
                build()
            else:
                return None
        R = C.CDLL(path)
        R.clref_type_sizes.argtypes = [u32p]
        R.clref_create.argtypes = [C.c_uint32, u32p, u64p, u32p, u32p, u8p, u32p, i32p, u32p, f32p, C.c_uint32, f32p]
        R.clref_create.restype = C.c_void_p
        R.clref_free.argtypes = [C.c_void_p]
        R.clref_node_geometry.argtypes = [C.c_void_p, f32p]
        R.clref_compute_moments.argtypes = [C.c_void_p, C.c_int]
        R.clref_compute_moments.restype = C.c_uint32
        R.clref_get_moments.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p]
        R.clref_traverse.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        R.clref_get_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        R.clref_forces.argtypes = [C.c_void_p, C.c_int, C.c_uint32, f32p, f32p]
        R.clref_integrate.argtypes = [C.c_void_p, f32p, f32p, C.c_float, C.c_void_p]
        R.clref_pair_force.argtypes = [C.c_float, C.c_float, f32p, f32p, f32p, f32p]
        R.clref_can_approx.argtypes = [C.c_uint64, f32p, f32p, f32p, f32p, u8p]
        R.clref_step.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_void_p]
        R.clref_step.restype = C.c_uint64
        _CLREF = R
    return _CLREF


# names of the nine sizes verify.cl reports, in its index order (include/nbody/device/types.h:18-27)
CLREF_TYPE_NAMES = ("leaf_t", "node_t", "leaf_value_t", "node_value_t", "leaf_moment_t", "node_moment_t", "leaf_field_t",
                    "node_field_t", "interaction_t")


def clref_type_sizes():
    out = np.zeros(9, np.uint32)
    clref_lib().clref_type_sizes(out)
    return dict(zip(CLREF_TYPE_NAMES, (int(v) for v in out)))


def clref_pair_force(qa, qb, pa, pb):
    """Force on a and on b from the reference's leaf_moment_field + leaf_field_to_force (src/field.cl:17-32, src/force.cl:4-10)."""
    fa, fb = np.zeros(3, np.float32), np.zeros(3, np.float32)
    clref_lib().clref_pair_force(qa, qb, np.asarray(pa, np.float32)[:3].copy(), np.asarray(pb, np.float32)[:3].copy(), fa, fb)
    return fa, fb


def clref_can_approx(pos_a, dim_a, pos_b, dim_b):
    """find_interactions' can_approx (src/interaction.cl:64-82) for n pairs of cells given by lower corner [n,3] and edge [n]."""
    n = len(dim_a)
    pa = np.zeros((n, 4), np.float32); pa[:, :3] = pos_a
    pb = np.zeros((n, 4), np.float32); pb[:, :3] = pos_b
    out = np.zeros(n, np.uint8)
    clref_lib().clref_can_approx(n, pa, np.ascontiguousarray(dim_a, np.float32), pb, np.ascontiguousarray(dim_b, np.float32), out)
    return out.astype(bool)


class ClRef:
    """The reference's kernels run on one octree: `tree` is an oracle.Tree, `particles12_sorted` the particle records in
    tree order. The octree itself is the oracle's (glade is absent); everything computed ON it comes from the reference's
    own kernel sources."""

    def __init__(self, tree, particles12_sorted, bounds=(1.0, 1.0, 1.0)):
        R = clref_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libclref.so is not built")
        self.tree = tree
        P = np.ascontiguousarray(particles12_sorted, np.float32)
        self.n = P.shape[0]
        self._h = R.clref_create(tree.num_nodes, tree.depth, tree.prefix, tree.leaf_index, tree.leaf_count, tree.has_children,
                                 np.ascontiguousarray(tree.child_off.reshape(-1)), tree.parent_off, tree.sibling,
                                 np.asarray(bounds, np.float32)[:3].copy(), self.n, P)

    def geometry(self):
        g = np.empty((self.tree.num_nodes, 4), np.float32)
        clref_lib().clref_node_geometry(self._h, g)
        return g

    def traverse(self):
        """(node interactions = M2L pairs, leaf interactions = P2P pairs, rounds), in the order the reference's host loop
        appends them (src/open_cl_simulation.cpp:242-266)."""
        a, b, r = C.c_uint64(), C.c_uint64(), C.c_uint64()
        clref_lib().clref_traverse(self._h, C.byref(a), C.byref(b), C.byref(r))
        node = np.empty((a.value, 2), np.uint32)
        leaf = np.empty((b.value, 2), np.uint32)
        clref_lib().clref_get_lists(self._h, _ptr(node), _ptr(leaf))
        return node, leaf, r.value

    def moments(self, repair_d5=True):
        """compute_moments_from_leafs + the upsweep (src/moment.cl); returns (launches, charge, dipole, cross, trace)."""
        m = self.tree.num_nodes
        launches = clref_lib().clref_compute_moments(self._h, 1 if repair_d5 else 0)
        q = np.empty(m, np.float32)
        d, c, t = (np.empty((m, 4), np.float32) for _ in range(3))
        clref_lib().clref_get_moments(self._h, q, d, c, t)
        return launches, q, d, c, t

    def forces(self, repair_d7=True, node_local_size=None):
        """(leaf forces, node forces) per particle: the two arrays the reference's integration adds
        (src/open_cl_simulation.cpp:589-590). Call traverse() and moments() first."""
        if node_local_size is None:
            node_local_size = 1 << max(5, int(np.ceil(np.log2(max(self.n, 1)))))  # every leaf of every target node is reached
        lf, nf = np.empty((self.n, 4), np.float32), np.empty((self.n, 4), np.float32)
        clref_lib().clref_forces(self._h, 1 if repair_d7 else 0, node_local_size, lf, nf)
        return lf[:, :3].copy(), nf[:, :3].copy()

    def integrate(self, leaf_forces, node_forces, dt):
        lf = np.zeros((self.n, 4), np.float32); lf[:, :3] = leaf_forces
        nf = np.zeros((self.n, 4), np.float32); nf[:, :3] = node_forces
        out = np.empty((self.n, 12), np.float32)
        clref_lib().clref_integrate(self._h, lf, nf, dt, _ptr(out))
        return out

    def step(self, dt, repair=True, node_local_size=32):
        """moments + traversal + fields + forces + integration in one call; returns (particles [n,12] in the same order, interactions)."""
        out = np.empty((self.n, 12), np.float32)
        pairs = clref_lib().clref_step(self._h, 1 if repair else 0, node_local_size, dt, _ptr(out))
        return out, int(pairs)

    def __del__(self):
        if getattr(self, "_h", None) and _CLREF is not None:
            _CLREF.clref_free(self._h)
            self._h = None
