// Test infrastructure (NOT product code). C-ABI harness around the UNMODIFIED
// reference CPU path /root/reference/src/naive_simulation.cpp:7-46, compiled
// where it lies by oracle/Makefile into oracle/_ref/libnaive_ref.so.
// Used (a) to pin oracle/oracle.cpp's restatement, (b) as bench.py's
// cpu_baseline / --impl reference arm (kind "reference").
#include "nbody/naive_simulation.h"
#include <cstdint>
#include <cstring>

using Sim = nbody::NaiveSimulation;
static_assert(sizeof(Sim::Particle) == 48, "Particle is 48 B AoS (SURVEY 3.2)");

extern "C" {

// particles: n records of 12 floats {pos[4], vel[4], mass, charge, pad, pad}, in/out.
// Runs `steps` calls of NaiveSimulation::step(); returns the last returned time.
float ref_naive_run(std::uint64_t n, float* particles, float force_constant,
                    float dt, std::uint32_t steps) {
	std::vector<Sim::Particle> ps;
	ps.reserve(n);
	for (std::uint64_t i = 0; i < n; ++i) {
		const float* r = particles + 12 * i;
		ps.push_back(Sim::Particle({r[0], r[1], r[2], r[3]}, {r[4], r[5], r[6], r[7]}, r[8], r[9]));
	}
	Sim sim(ps, force_constant, dt);
	float t = 0.0f;
	for (std::uint32_t s = 0; s < steps; ++s) t = sim.step();
	std::vector<Sim::Particle> out = sim.particles();
	for (std::uint64_t i = 0; i < n; ++i) {
		float* r = particles + 12 * i;
		for (int k = 0; k < 4; ++k) { r[k] = out[i].position[k]; r[4 + k] = out[i].velocity[k]; }
		r[8] = out[i].mass; r[9] = out[i].charge;
	}
	return t;
}

std::uint32_t ref_particle_size(void) { return (std::uint32_t) sizeof(Sim::Particle); }

}
