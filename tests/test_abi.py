"""CPU tests of the boundary: the C-ABI library loads and exports every symbol that
include/nbody_cuda.h declares; no compute call is made without a GPU, and the product
refuses to run (never falls back to a CPU path) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nbody_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nbody_cuda.h")).read()
    return sorted(set(re.findall(r"\b(nbody_cuda_[a-z_0-9]+)\s*\(", src)))


def test_header_and_python_mirror_agree():
    assert declared_symbols() == sorted(nbody_b200.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(nbody_b200.LIB_PATH):
        nbody_b200.build_library()
    lib = C.CDLL(nbody_b200.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_struct_layouts():
    assert C.sizeof(nbody_b200.Config) == 4 + 16 + 4 * 4 + 4 * 6 + 4 + 7 * 4   # 92 bytes, mirrors nbody_cuda_config
    assert C.sizeof(nbody_b200.Stats) == 12 * 8 + 10 * 4 + 3 * 8 + 4 * 4
    cfg = nbody_b200.default_config()
    assert cfg.abi_version == 1 and cfg.leaf_capacity == 8 and cfg.order == 4 and cfg.max_depth == 21
    assert abs(cfg.softening - 0.01) < 1e-9 and abs(cfg.mac_ratio - 0.5) < 1e-9 and cfg.force_constant == 1.0
    assert list(cfg.bounds)[:3] == [1.0, 1.0, 1.0]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    P = np.zeros((4, 12), np.float32)
    P[:, 8] = 1; P[:, 9] = 1
    with pytest.raises(nbody_b200.NbodyCudaError) as e:
        nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(nbody_b200.NbodyCudaError):
        nbody_b200.direct_field(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32))
    with pytest.raises(nbody_b200.NbodyCudaError) as e:
        nbody_b200.sort_runs(np.arange(8, dtype=np.uint64), [0, 4, 8])
    assert "no CPU fallback" in str(e.value)


def test_sort_runs_argument_validation():
    for keys, bound in ((np.arange(8), [1, 4, 8]), (np.arange(8), [0, 4, 7]), (np.arange(8), [0, 6, 4, 8]), (np.arange(8), [0] + [8] * 17)):
        with pytest.raises(nbody_b200.NbodyCudaError):
            nbody_b200.sort_runs(np.asarray(keys, np.uint64), bound)


def test_argument_validation_without_gpu():
    lib = nbody_b200.load_library()
    h = C.c_void_p()
    cfg = nbody_b200.default_config()
    P = np.zeros((4, 12), np.float32)
    assert lib.nbody_cuda_create(C.byref(cfg), None, 4, C.byref(h)) == 1          # NULL particles
    cfg.order = 7
    assert lib.nbody_cuda_create(C.byref(cfg), P.ctypes.data_as(C.c_void_p), 4, C.byref(h)) == 1
    assert b"order" in lib.nbody_cuda_last_error()
    cfg = nbody_b200.default_config(leaf_capacity=0)
    assert lib.nbody_cuda_create(C.byref(cfg), P.ctypes.data_as(C.c_void_p), 4, C.byref(h)) == 1
    cfg = nbody_b200.default_config()
    assert lib.nbody_cuda_create(C.byref(cfg), P.ctypes.data_as(C.c_void_p), 0, C.byref(h)) == 1
    assert lib.nbody_cuda_num_particles(None) == 0


def test_oracle_is_not_linked_into_the_product():
    # the product library must not depend on anything under oracle/
    import subprocess
    out = subprocess.run(["ldd", nbody_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "naive_ref" not in out
    # ... and nothing outside tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may use it
    for top in ("nbody_b200", "include", "examples", "tools", "script"):
        for root, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", ".sh")):
                    txt = open(os.path.join(root, f)).read()
                    assert "import oracle" not in txt and "liboracle" not in txt and "from oracle" not in txt, os.path.join(root, f)


def test_plain_c99_client_compiles_and_runs(tmp_path):
    """include/nbody_cuda.h is C, not C++: a -std=c99 -pedantic -Werror client links against the library and runs every
    entry point that needs no device (examples/abi_host_check.c)."""
    import subprocess
    if not os.path.exists(nbody_b200.LIB_PATH):
        nbody_b200.build_library()
    exe = str(tmp_path / "abi_host_check")
    libdir = os.path.dirname(nbody_b200.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_host_check.c"), "-L" + libdir, "-lnbody_cuda", "-lm",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe, str(tmp_path / "c.ckp")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "abi_host_check: ok" in r.stdout, r.stdout + r.stderr


CPP_CLIENT = r"""
#include <iostream>
#include <sstream>
#include "nbody/cuda_simulation.h"
int main() {
	using S = nbody::CudaSimulation;
	static_assert(sizeof(S::Particle) == 48 && alignof(S::Particle) == 16, "the reference's Particle layout");
	std::vector<S::Particle> p(4, S::Particle(S::Vector(), S::Vector(), 1.0f, 1.0f));
	std::ostringstream log;
	int failures = 0;
	try {                                          // every method of the wrapper is instantiated; none can run without a device
		S sim({1.0f, 1.0f, 1.0f, 0.0f}, p, 1e-3f, log);
		sim.step(); sim.particles(); sim.permutation(); sim.stats(); sim.setTimeStep(1e-3f); sim.timeStep(); sim.time();
		sim.stepsDone(); sim.saveCheckpoint("x.ckp");
		std::uint64_t a, b; sim.ownedRange(a, b); sim.ownedParticles();
	} catch (const std::runtime_error& e) { ++failures; std::cout << e.what() << "\n"; }
	try { S sim(std::string("/nonexistent.ckp"), log); } catch (const std::runtime_error&) { ++failures; }
	try {
		std::uint8_t id[128] = {0};
		S sim({1.0f, 1.0f, 1.0f, 0.0f}, p, 1e-3f, log, 8, 0, 0, 2, id);
	} catch (const std::runtime_error&) { ++failures; }
	std::cout << "failures " << failures << "\n";
	return 0;
}
"""


def test_cpp_wrapper_compiles_and_fails_loudly_without_a_device(tmp_path):
    """include/nbody/cuda_simulation.h (the drop-in for OpenClSimulation) and the demo driver compile with -Wall -Wextra
    -Werror; without a GPU every constructor throws std::runtime_error (the reference's error channel), nothing falls back."""
    import subprocess
    import torch
    libdir = os.path.dirname(nbody_b200.LIB_PATH)
    src = tmp_path / "client.cpp"
    src.write_text(CPP_CLIENT)
    common = ["-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-lnbody_cuda",
              "-Wl,-rpath," + libdir]
    subprocess.check_call(["g++", str(src), "-o", str(tmp_path / "client")] + common)
    subprocess.check_call(["g++", os.path.join(ROOT, "examples", "nbody_main.cpp"), "-o", str(tmp_path / "nbody_main")] + common)
    if torch.cuda.is_available():
        return
    r = subprocess.run([str(tmp_path / "client")], capture_output=True, text=True, timeout=60, cwd=str(tmp_path))
    assert r.returncode == 0 and "failures 3" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([str(tmp_path / "nbody_main"), "--n", "100", "--steps", "1", "--quiet", "--csv", "none"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


REFERENCE_LAYOUT_CLIENT = r"""
// the REFERENCE's own headers (include/nbody/simulation.h, include/nbody/device/types.h) next to the C ABI's record
#include <cstddef>
#include <vector>
#include "nbody/device/types.h"
#include "nbody/simulation.h"
#include "nbody_cuda.h"
using RefSim = nbody::Simulation<nbody::device::scalar_t, nbody::device::vector_t>;
using P = RefSim::Particle;
static_assert(sizeof(P) == sizeof(nbody_particle) && sizeof(P) == 48, "size");
static_assert(alignof(P) == 16, "alignment");
static_assert(offsetof(P, position) == offsetof(nbody_particle, position), "position");
static_assert(offsetof(P, velocity) == offsetof(nbody_particle, velocity), "velocity");
static_assert(offsetof(P, mass) == offsetof(nbody_particle, mass), "mass");
static_assert(offsetof(P, charge) == offsetof(nbody_particle, charge), "charge");
int main() {
	// the cast INTEGRATION.md section 2 uses, on values: a reference Particle read through the boundary record
	std::vector<P> v(3, P({1.0f, 2.0f, 3.0f, 0.0f}, {4.0f, 5.0f, 6.0f, 0.0f}, 7.0f, 8.0f));
	const nbody_particle* q = reinterpret_cast<const nbody_particle*>(v.data());
	for (int i = 0; i < 3; ++i)
		if (q[i].position[0] != 1.0f || q[i].position[2] != 3.0f || q[i].velocity[1] != 5.0f || q[i].mass != 7.0f || q[i].charge != 8.0f) return 1;
	return 0;
}
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="needs the reference's headers (/root/reference)")
def test_reference_particle_type_is_the_boundary_record(tmp_path):
    """The reference's Simulation<float, vector_t>::Particle, compiled from the reference's OWN headers (with the 6-typedef
    CL/cl2.hpp shim of oracle/shim), has the size, alignment and field offsets of nbody_particle: the reinterpret_cast a
    maintainer's binding uses is layout-exact."""
    import subprocess
    src = tmp_path / "layout.cpp"
    src.write_text(REFERENCE_LAYOUT_CLIENT)
    exe = str(tmp_path / "layout")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-Wno-invalid-offsetof", "-I" + os.path.join(ROOT, "oracle", "shim"),
                           "-I/root/reference/include", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    assert subprocess.run([exe]).returncode == 0


def test_every_entry_point_rejects_null_arguments():
    """A binding written in another language passes NULL sooner or later: every entry point answers with an error code (or does nothing)
    instead of crashing. None of these paths touches a device."""
    L = nbody_b200.load_library()
    N = None
    f32 = C.c_float
    calls = {
        "get_particles": lambda: L.nbody_cuda_get_particles(N, N, 0),
        "get_permutation": lambda: L.nbody_cuda_get_permutation(N, N, 0),
        "get_accelerations": lambda: L.nbody_cuda_get_accelerations(N, N, 0),
        "get_keys": lambda: L.nbody_cuda_get_keys(N, N, 0),
        "get_tree": lambda: L.nbody_cuda_get_tree(N, N, 0, N, N, N, N, N, N, N, N, N),
        "get_lists": lambda: L.nbody_cuda_get_lists(N, N, N, N, N),
        "get_expansions": lambda: L.nbody_cuda_get_expansions(N, N, N, 0),
        "get_stats": lambda: L.nbody_cuda_get_stats(N, N),
        "set_time_step": lambda: L.nbody_cuda_set_time_step(N, f32(1e-3)),
        "get_time_step": lambda: L.nbody_cuda_get_time_step(N, N, N, N),
        "get_time": lambda: L.nbody_cuda_get_time(N, N, N),
        "checkpoint_save": lambda: L.nbody_cuda_checkpoint_save(N, b"/tmp/never_written.ckp"),
        "checkpoint_load": lambda: L.nbody_cuda_checkpoint_load(N, N, N),
        "checkpoint_info": lambda: L.nbody_cuda_checkpoint_info(N, N),
        "checkpoint_read": lambda: L.nbody_cuda_checkpoint_read(N, N, N, 0),
        "checkpoint_write": lambda: L.nbody_cuda_checkpoint_write(N, N, N, N),
        "owned_range": lambda: L.nbody_cuda_owned_range(N, N, N),
        "get_owned_particles": lambda: L.nbody_cuda_get_owned_particles(N, N, 0),
        "set_owned_particles": lambda: L.nbody_cuda_set_owned_particles(N, N, 0),
        "set_particles": lambda: L.nbody_cuda_set_particles(N, N, 0),
        "step": lambda: L.nbody_cuda_step(N, N),
        "group_step": lambda: L.nbody_cuda_group_step(N, 2, N),
        "create_group": lambda: L.nbody_cuda_create_group(N, N, 0, 2, N),
        "create_distributed": lambda: L.nbody_cuda_create_distributed(N, N, 0, 0, 0, 0, 1, N, N),
        "comm_unique_id": lambda: L.nbody_cuda_comm_unique_id(N),
        "sort_runs": lambda: L.nbody_cuda_sort_runs(0, N, 0, N, 1, N, N),
        "direct_field": lambda: L.nbody_cuda_direct_field(0, N, 0, N, 0, f32(0.01), N, N, 1),
        "rebalance": lambda: L.nbody_cuda_rebalance(2, N, N, f32(0.5), N),
        "create": lambda: L.nbody_cuda_create(N, N, 0, N),
    }
    for name, call in calls.items():
        assert call() != 0, name
        assert len(L.nbody_cuda_last_error()) > 0, name
    # the void ones, and the rule that returns a value
    L.nbody_cuda_destroy(N); L.nbody_cuda_destroy_group(N, 2); L.nbody_cuda_default_config(N); L.nbody_cuda_tuned_config(N)
    assert L.nbody_cuda_num_particles(N) == 0 and L.nbody_cuda_next_time_step(N, f32(1.0)) == 0.0
    exercised = set(calls) | {"destroy", "destroy_group", "default_config", "tuned_config", "num_particles", "next_time_step", "last_error"}
    assert {s.replace("nbody_cuda_", "") for s in nbody_b200.EXPORTED_SYMBOLS} == exercised
