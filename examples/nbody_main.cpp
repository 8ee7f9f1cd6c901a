// Demo driver with the behaviour of the reference's src/main.cpp:16-116: generate a random
// cube of particles, construct the simulation, loop step() until maxTime and append one
// CSV row per step ("time,x0,y0,z0,x1,...", src/main.cpp:88-95) to particles.csv.
// Unlike the reference it takes its parameters from the command line and a fixed seed.
//   nbody_main [--n N] [--steps S] [--dt DT] [--seed SEED] [--csv FILE|none] [--csv-max K] [--quiet]
//              [--eta ETA]            variable time step (nbody_cuda_config::time_step_eta; 0 = the reference's fixed step)
//              [--checkpoint FILE]    write the full state after the last step
//              [--restart FILE]       continue from a checkpoint instead of generating particles (appends to the CSV)
// Build: g++ -std=c++14 -O2 -Iinclude examples/nbody_main.cpp -Lnbody_b200 -lnbody_cuda -Wl,-rpath,$PWD/nbody_b200 -o nbody_main
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

// The random cube of src/main.cpp:23-66 (positions in the bounds, speed 0.1 in a random direction, masses in [1, 10]).
static std::unique_ptr<Simulation> generate(std::uint64_t n, std::uint64_t seed, float dt, float eta, std::ostream& log) {
	std::mt19937_64 rng(seed);
	std::uniform_real_distribution<float> uni(0.0f, 1.0f);
	Simulation::Vector bounds = {1.0f, 1.0f, 1.0f, 0.0f};
	const float velocityMax = 0.1f, massRange[2] = {1.0f, 10.0f};
	std::cout << "Generating particles.\n";
	std::vector<Simulation::Particle> particles;
	particles.reserve(n);
	double total_mass = 0;
	for (std::uint64_t i = 0; i < n; ++i) {
		Simulation::Vector position = {bounds[0] * uni(rng), bounds[1] * uni(rng), bounds[2] * uni(rng), 0.0f};
		const float theta = 2.0f * (float) M_PI * uni(rng), phi = std::acos(2.0f * (uni(rng) - 0.5f));
		Simulation::Vector velocity = {velocityMax * std::sin(phi) * std::cos(theta), velocityMax * std::sin(phi) * std::sin(theta),
		                               velocityMax * std::cos(phi), 0.0f};
		const float f = uni(rng), mass = massRange[0] * (1.0f - f) + massRange[1] * f;
		total_mass += mass;
		particles.push_back(Simulation::Particle(position, velocity, mass, mass));  // gravity: charge = mass
	}
	nbody_cuda_config cfg;
	nbody_cuda_default_config(&cfg);
	cfg.force_constant = (float) (1.0 / total_mass);  // keeps the free-fall time of the cube of order 1 (see nbody_b200/workloads.py)
	cfg.time_step_eta = eta;
	return std::unique_ptr<Simulation>(new Simulation(bounds, particles, dt, log, &cfg));
}

int main(int argc, char** argv) {
	std::uint64_t n = 1000000, seed = 42, csv_max = 1000;
	unsigned steps = 10;
	float dt = 0.001f, eta = 0.0f;
	std::string csv = "particles.csv", checkpoint, restart;
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
		else { std::cerr << "unknown option " << a << "\n"; return 2; }
	}
	try {
		std::ostringstream sink;
		std::ostream& log = quiet ? static_cast<std::ostream&>(sink) : std::cout;
		std::unique_ptr<Simulation> held;
		if (!restart.empty()) {
			held.reset(new Simulation(restart, log));
			std::cout << "Restored " << held->particles().size() << " particles at t=" << held->time() << " (step " << held->stepsDone() << ").\n";
		} else {
			held = generate(n, seed, dt, eta, log);
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
