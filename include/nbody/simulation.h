// nbody::Simulation — the step / particle-state interface every solver implements.
//
// Interface-compatible with the reference's include/nbody/simulation.h:6-37
// (duanebyer/nbody): same template parameters, same nested Particle record, same two
// pure virtuals, so code written against the reference's Simulation compiles against
// this header unchanged. (Written for this project; it also pulls in <vector>, which
// the reference leaves to the includer.)
#pragma once

#include <vector>

namespace nbody {

template <typename TScalar, typename TVector>
class Simulation {
public:
	using Scalar = TScalar;
	using Vector = TVector;

	// One body. With the 16-byte float4 vector this record is 48 bytes, the layout that
	// crosses the C ABI as nbody_particle (include/nbody_cuda.h).
	struct Particle final {
		Vector position;
		Vector velocity;
		Scalar mass;
		Scalar charge;

		Particle(Vector position_, Vector velocity_, Scalar mass_, Scalar charge_)
		    : position(position_), velocity(velocity_), mass(mass_), charge(charge_) {}
	};

	virtual ~Simulation() = default;

	// Advances the system by one time step and returns the new simulation time.
	virtual Scalar step() = 0;
	// Returns a copy of the current particle state.
	virtual std::vector<Particle> particles() const = 0;
};

}  // namespace nbody
