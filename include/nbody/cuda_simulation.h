// nbody::CudaSimulation — the B200 FMM solver behind the reference's Simulation API.
//
// Drop-in for nbody::OpenClSimulation (include/nbody/open_cl_simulation.h:19-203):
// same base class instantiation Simulation<float, 16-byte float4>, same constructor
// shape (bounds, particles, timeStep, log), same step()/particles() contract
// (src/open_cl_simulation.cpp:53-106: step() is blocking and returns the new time,
// particles() returns the state in tree order). Header-only: it only forwards to the C
// ABI in include/nbody_cuda.h (link with -lnbody_cuda); failures become
// std::runtime_error like the reference's (src/open_cl_simulation.cpp:630,648,767).
#pragma once

#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../nbody_cuda.h"
#include "simulation.h"

namespace nbody {
namespace device {

using scalar_t = float;

// 16-byte aligned 4-float vector with array access; behaves like the reference's
// device::vector_t (include/nbody/device/types.h:53-74) without the OpenCL headers.
class alignas(16) vector_t final {
public:
	vector_t() : _v{0.0f, 0.0f, 0.0f, 0.0f} {}
	vector_t(std::initializer_list<scalar_t> list) : _v{0.0f, 0.0f, 0.0f, 0.0f} {
		unsigned i = 0;
		for (scalar_t x : list) {
			if (i < 4) _v[i++] = x;
		}
	}
	scalar_t& operator[](unsigned index) { return _v[index]; }
	scalar_t const& operator[](unsigned index) const { return _v[index]; }

private:
	scalar_t _v[4];
};

}  // namespace device

class CudaSimulation final : public Simulation<device::scalar_t, device::vector_t> {
public:
	// Mirrors OpenClSimulation(bounds, particles, timeStep, log); `config` (optional) overrides the
	// reference's compile-time constants (MAC ratio, softening, capacity, ...), see nbody_cuda_config.
	// Without `config`: the reference's constants with the B200-tuned node capacity (nbody_cuda_tuned_config: 48 instead of the
	// reference's provisional 8; same accuracy, 2.7x the speed). Pass nbody_cuda_default_config() for capacity 8.
	// NOTE on the integrator: the default is kick-drift (x += v_new dt), NaiveSimulation's rule (src/naive_simulation.cpp:28-42);
	// OpenClSimulation integrates x += v_old dt (src/open_cl_simulation.cpp:602-607): set config->integrator = NBODY_EXPLICIT_EULER
	// to reproduce that trajectory (the two differ by a dt^2 per step).
	CudaSimulation(device::vector_t bounds, std::vector<Particle> particles, Scalar timeStep, std::ostream& log,
	               const nbody_cuda_config* config = nullptr)
	    : _log(log), _time(0.0f) {
		static_assert(sizeof(Particle) == sizeof(nbody_particle), "Particle must be the 48-byte boundary record");
		nbody_cuda_config cfg;
		if (config) cfg = *config; else nbody_cuda_tuned_config(&cfg);
		for (int k = 0; k < 4; ++k) cfg.bounds[k] = bounds[k];
		cfg.time_step = timeStep;
		_log << "Creating the CUDA simulation (" << particles.size() << " particles).\n";
		check(nbody_cuda_create(&cfg, reinterpret_cast<const nbody_particle*>(particles.data()), particles.size(), &_sim));
	}
	// Continue a run from a checkpoint written by saveCheckpoint() (the reference cannot resume: its particles.csv drops
	// velocities, masses and charges, src/main.cpp:88-95). `config` (optional) replaces the stored configuration.
	CudaSimulation(const std::string& checkpointPath, std::ostream& log, const nbody_cuda_config* config = nullptr)
	    : _log(log), _time(0.0f) {
		_log << "Restoring the CUDA simulation from " << checkpointPath << ".\n";
		check(nbody_cuda_checkpoint_load(checkpointPath.c_str(), config, &_sim));
		check(nbody_cuda_get_time(_sim, &_time, nullptr));
	}
	// One rank of a multi-GPU run (one process per GPU, Morton-range partition). Without `config`: the
	// partitioned scheme (NBODY_FLAG_PARTITIONED: the rank holds only its own particles and imports a locally essential tree;
	// particles() / permutation() then return THIS rank's particles, in tree order; ranks in rank order = the global tree order).
	// `particles` is THIS rank's contiguous slice of the global set, `globalOffset` the index of its first particle and
	// `uniqueId` the 128 bytes nbody_cuda_comm_unique_id() returned on rank 0. Collective: every rank constructs at once.
	CudaSimulation(device::vector_t bounds, std::vector<Particle> particles, Scalar timeStep, std::ostream& log,
	               std::uint64_t globalCount, std::uint64_t globalOffset, int rank, int world, const std::uint8_t (&uniqueId)[128],
	               const nbody_cuda_config* config = nullptr)
	    : _log(log), _time(0.0f) {
		nbody_cuda_config cfg;
		if (config) cfg = *config; else { nbody_cuda_tuned_config(&cfg); cfg.flags |= NBODY_FLAG_PARTITIONED; }
		for (int k = 0; k < 4; ++k) cfg.bounds[k] = bounds[k];
		cfg.time_step = timeStep;
		_log << "Creating rank " << rank << " of " << world << " of the CUDA simulation (" << particles.size() << " of " << globalCount << " particles).\n";
		check(nbody_cuda_create_distributed(&cfg, reinterpret_cast<const nbody_particle*>(particles.data()), particles.size(), globalCount,
		                                    globalOffset, rank, world, uniqueId, &_sim));
	}
	CudaSimulation(const CudaSimulation&) = delete;
	CudaSimulation& operator=(const CudaSimulation&) = delete;
	~CudaSimulation() override { nbody_cuda_destroy(_sim); }

	Scalar step() override {
		_log << "Starting a new step (t=" << _time << ").\n";
		check(nbody_cuda_step(_sim, &_time));
		_log << "Step finished.\n";
		return _time;
	}

	// Tree (Morton / DFS leaf) order, like OpenClSimulation::particles(); see permutation().
	std::vector<Particle> particles() const override {
		const std::uint64_t n = nbody_cuda_num_particles(_sim);
		std::vector<Particle> out(n, Particle(Vector(), Vector(), 0.0f, 0.0f));
		check(nbody_cuda_get_particles(_sim, reinterpret_cast<nbody_particle*>(out.data()), n));
		return out;
	}

	// permutation()[i] = index in the constructor's vector of the particle now at position i
	// (the reference loses particle identity across steps).
	std::vector<std::uint32_t> permutation() const {
		std::vector<std::uint32_t> p(nbody_cuda_num_particles(_sim));
		check(nbody_cuda_get_permutation(_sim, p.data(), p.size()));
		return p;
	}

	// Multi-GPU: the slice [first, first + count) of the global tree-ordered particle array this rank owns after the last step, and
	// just those particles (partitioned scheme: the same as particles(); replicated scheme: particles() returns the whole state).
	void ownedRange(std::uint64_t& first, std::uint64_t& count) const { check(nbody_cuda_owned_range(_sim, &first, &count)); }
	std::vector<Particle> ownedParticles() const {
		std::uint64_t first = 0, count = 0;
		ownedRange(first, count);
		std::vector<Particle> out(count, Particle(Vector(), Vector(), 0.0f, 0.0f));
		check(nbody_cuda_get_owned_particles(_sim, reinterpret_cast<nbody_particle*>(out.data()), count));
		return out;
	}

	void saveCheckpoint(const std::string& path) const { check(nbody_cuda_checkpoint_save(_sim, path.c_str())); }

	// Variable time step (the reference's TODO:2): with nbody_cuda_config::time_step_eta > 0 the library picks every
	// step from the largest acceleration of the step before; setTimeStep() lets the caller drive it instead.
	void setTimeStep(Scalar dt) { check(nbody_cuda_set_time_step(_sim, dt)); }
	Scalar timeStep() const {
		Scalar dt = 0.0f;
		check(nbody_cuda_get_time_step(_sim, &dt, nullptr, nullptr));
		return dt;
	}
	Scalar time() const { return _time; }
	std::uint64_t stepsDone() const {
		std::uint64_t k = 0;
		check(nbody_cuda_get_time(_sim, nullptr, &k));
		return k;
	}

	nbody_cuda_stats stats() const {
		nbody_cuda_stats s;
		check(nbody_cuda_get_stats(_sim, &s));
		return s;
	}

private:
	static void check(int rc) {
		if (rc != NBODY_OK) throw std::runtime_error(std::string("nbody_cuda: ") + nbody_cuda_last_error());
	}

	std::ostream& _log;
	nbody_cuda_sim* _sim = nullptr;
	Scalar _time;
};

}  // namespace nbody
