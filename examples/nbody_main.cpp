// Demo driver with the behaviour of the reference's src/main.cpp:16-116: generate a random
// cube of particles, construct the simulation, loop step() until maxTime and append one
// CSV row per step ("time,x0,y0,z0,x1,...", src/main.cpp:88-95) to particles.csv.
// Unlike the reference it takes its parameters from the command line and a fixed seed.
//   nbody_main [--n N] [--steps S] [--dt DT] [--seed SEED] [--csv FILE|none] [--csv-max K] [--quiet]
//              [--distribution uniform|plummer|two-galaxies]   initial condition (default uniform, the reference's)
//              [--capacity C]         octree node capacity (default: the B200 tuning, 48; the reference's constant is 8)
//              [--eta ETA]            variable time step (nbody_cuda_config::time_step_eta; 0 = the reference's fixed step)
//              [--checkpoint FILE]    write the full state after the last step
//              [--restart FILE]       continue from a checkpoint instead of generating particles (appends to the CSV)
// Build: g++ -std=c++14 -O2 -Iinclude examples/nbody_main.cpp -Lnbody_b200 -lnbody_cuda -Wl,-rpath,$PWD/nbody_b200 -o nbody_main
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <string>

#include "nbody/cuda_simulation.h"

using Simulation = nbody::CudaSimulation;

// A Plummer sphere (Aarseth, Henon & Wielen 1974) of `count` equal-mass particles, total mass `mass`, scale radius a = 1/32, truncated
// at rmax, in virial equilibrium for G = 1 — the model of nbody_b200/workloads.py, drawn from this driver's own generator.
template <typename Rng>
static void plummer(std::vector<Simulation::Particle>& out, std::uint64_t count, Rng& rng, const float (&centre)[3], float rmax, float mass,
                    float bulk_vx) {
	std::uniform_real_distribution<double> uni(0.0, 1.0);
	const double a = 1.0 / 32.0;
	auto direction = [&](double (&d)[3]) {
		const double theta = 2.0 * M_PI * uni(rng), c = 2.0 * uni(rng) - 1.0, s = std::sqrt(std::max(0.0, 1.0 - c * c));
		d[0] = s * std::cos(theta); d[1] = s * std::sin(theta); d[2] = c;
	};
	for (std::uint64_t i = 0; i < count; ++i) {
		double r;
		do {  // radius by inversion of the cumulative mass profile, redrawn beyond the truncation radius
			const double u = std::min(std::max(uni(rng), 1e-7), 1.0 - 1e-7);
			r = a / std::sqrt(std::pow(u, -2.0 / 3.0) - 1.0);
		} while (r > rmax);
		double q, y;
		do { q = uni(rng); y = 0.1 * uni(rng); } while (y >= q * q * std::pow(1.0 - q * q, 3.5));  // speed / escape speed by rejection
		const double v = q * std::sqrt(2.0 * mass) * std::pow(r * r + a * a, -0.25);
		double dp[3], dv[3];
		direction(dp); direction(dv);
		Simulation::Vector position = {(float) (centre[0] + r * dp[0]), (float) (centre[1] + r * dp[1]), (float) (centre[2] + r * dp[2]), 0.0f};
		Simulation::Vector velocity = {(float) (v * dv[0]) + bulk_vx, (float) (v * dv[1]), (float) (v * dv[2]), 0.0f};
		const float m = mass / (float) count;
		out.push_back(Simulation::Particle(position, velocity, m, m));
	}
}

// The initial conditions: the random cube of src/main.cpp:23-66 (positions in the bounds, speed 0.1 in a random direction, masses in
// [1, 10]), a Plummer sphere, or two Plummer spheres of n/2 approaching each other (BASELINE configs 3 and 4).
static std::unique_ptr<Simulation> generate(std::uint64_t n, std::uint64_t seed, float dt, float eta, const std::string& distribution,
                                            unsigned capacity, std::ostream& log) {
	std::mt19937_64 rng(seed);
	std::uniform_real_distribution<float> uni(0.0f, 1.0f);
	Simulation::Vector bounds = {1.0f, 1.0f, 1.0f, 0.0f};
	const float velocityMax = 0.1f, massRange[2] = {1.0f, 10.0f};
	std::cout << "Generating particles (" << distribution << ").\n";
	std::vector<Simulation::Particle> particles;
	particles.reserve(n);
	double total_mass = 0;
	nbody_cuda_config cfg;
	nbody_cuda_tuned_config(&cfg);
	if (capacity) cfg.leaf_capacity = capacity;
	cfg.time_step_eta = eta;
	if (distribution == "plummer" || distribution == "two-galaxies") {
		if (distribution == "plummer") {
			const float c[3] = {0.5f, 0.5f, 0.5f};
			plummer(particles, n, rng, c, 0.45f, 1.0f, 0.0f);
		} else {
			const float c0[3] = {0.3f, 0.5f, 0.5f}, c1[3] = {0.7f, 0.5f, 0.5f};
			plummer(particles, n / 2, rng, c0, 0.28f, 0.5f, 0.05f);
			plummer(particles, n - n / 2, rng, c1, 0.28f, 0.5f, -0.05f);
		}
		cfg.force_constant = 1.0f;  // the models are in virial equilibrium for G = 1, total mass 1
		return std::unique_ptr<Simulation>(new Simulation(bounds, particles, dt, log, &cfg));
	}
	for (std::uint64_t i = 0; i < n; ++i) {
		Simulation::Vector position = {bounds[0] * uni(rng), bounds[1] * uni(rng), bounds[2] * uni(rng), 0.0f};
		const float theta = 2.0f * (float) M_PI * uni(rng), phi = std::acos(2.0f * (uni(rng) - 0.5f));
		Simulation::Vector velocity = {velocityMax * std::sin(phi) * std::cos(theta), velocityMax * std::sin(phi) * std::sin(theta),
		                               velocityMax * std::cos(phi), 0.0f};
		const float f = uni(rng), mass = massRange[0] * (1.0f - f) + massRange[1] * f;
		total_mass += mass;
		particles.push_back(Simulation::Particle(position, velocity, mass, mass));  // gravity: charge = mass
	}
	cfg.force_constant = (float) (1.0 / total_mass);  // keeps the free-fall time of the cube of order 1 (see nbody_b200/workloads.py)
	return std::unique_ptr<Simulation>(new Simulation(bounds, particles, dt, log, &cfg));
}

int main(int argc, char** argv) {
	std::uint64_t n = 1000000, seed = 42, csv_max = 1000;
	unsigned steps = 10;
	float dt = 0.001f, eta = 0.0f;
	std::string csv = "particles.csv", checkpoint, restart, distribution = "uniform";
	unsigned capacity = 0;
	bool quiet = false;
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : "0"; };
		if (a == "--n") n = std::strtoull(next(), nullptr, 10);
		else if (a == "--steps") steps = (unsigned) std::strtoul(next(), nullptr, 10);
		else if (a == "--dt") dt = std::strtof(next(), nullptr);
		else if (a == "--seed") seed = std::strtoull(next(), nullptr, 10);
		else if (a == "--csv") csv = next();
		else if (a == "--csv-max") csv_max = std::strtoull(next(), nullptr, 10);
		else if (a == "--quiet") quiet = true;
		else if (a == "--eta") eta = std::strtof(next(), nullptr);
		else if (a == "--checkpoint") checkpoint = next();
		else if (a == "--restart") restart = next();
		else if (a == "--distribution") distribution = next();
		else if (a == "--capacity") capacity = (unsigned) std::strtoul(next(), nullptr, 10);
		else if (a == "--help" || a == "-h") {
			std::cout << "usage: nbody_main [--n N] [--steps K] [--dt DT] [--seed S] [--distribution uniform|plummer|two-galaxies] [--capacity C]\n"
			             "                  [--eta ETA] [--csv FILE|none] [--csv-max ROWS] [--checkpoint FILE] [--restart FILE] [--quiet]\n";
			return 0;
		}
		else { std::cerr << "unknown option " << a << " (--help lists them)\n"; return 2; }
	}
	if (restart.empty() && (n == 0 || n > 0xfffffff0ull)) {
		std::cerr << "--n must be a particle count in [1, 2^32-16)\n";
		return 2;
	}
	if (distribution != "uniform" && distribution != "plummer" && distribution != "two-galaxies") {
		std::cerr << "unknown distribution " << distribution << " (uniform, plummer, two-galaxies)\n";
		return 2;
	}
	try {
		std::ostringstream sink;
		std::ostream& log = quiet ? static_cast<std::ostream&>(sink) : std::cout;
		std::unique_ptr<Simulation> held;
		if (!restart.empty()) {
			held.reset(new Simulation(restart, log));
			std::cout << "Restored " << held->particles().size() << " particles at t=" << held->time() << " (step " << held->stepsDone() << ").\n";
		} else {
			held = generate(n, seed, dt, eta, distribution, capacity, log);
		}
		Simulation& simulation = *held;
		std::ofstream dataFile;
		if (csv != "none") dataFile.open(csv, restart.empty() ? std::ios::out : std::ios::app);
		std::cout << "Starting simulation.\n";
		Simulation::Scalar time = 0.0f;
		for (unsigned s = 0; s < steps; ++s) {
			time = simulation.step();
			if (dataFile.is_open()) {
				dataFile << time;
				std::uint64_t k = 0;
				for (const Simulation::Particle& p : simulation.particles()) {
					if (k++ >= csv_max) break;
					dataFile << "," << p.position[0] << "," << p.position[1] << "," << p.position[2];
				}
				dataFile << "\n";
			}
		}
		if (!checkpoint.empty()) {
			simulation.saveCheckpoint(checkpoint);
			std::cout << "Checkpoint written to " << checkpoint << " (step " << simulation.stepsDone() << ", next dt " << simulation.timeStep() << ").\n";
		}
		const nbody_cuda_stats st = simulation.stats();
		std::cout << "t=" << time << "  last step: " << st.ms_total << " ms, " << st.n_nodes << " nodes, " << st.m2l_interactions
		          << " M2L, " << st.p2p_interactions << " P2P evaluations\n";
	} catch (const std::exception& e) {
		std::cerr << "error: " << e.what() << "\n";
		return 1;
	}
	return 0;
}
