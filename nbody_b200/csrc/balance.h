// Per-step load rebalancing of the Morton-range partition, host side.
//
// Feedback scheme: after every step each rank reports the device time of the stages it runs for its own
// slice only (traversal, M2L, L2L, P2P + integrator). The cost of a particle is taken as constant inside a
// rank's slice, so the cumulative cost over the tree-ordered particle array is piecewise linear; the new
// boundaries are the points where that curve crosses k/world of the total, approached with a damping factor
// so that timing noise does not make the boundaries oscillate. Every rank evaluates this function on the
// same all-gathered numbers, so all ranks obtain the same boundaries without further communication.
// Plain C++ (no CUDA): tests/test_multi_host.py calls it through nbody_cuda_rebalance().
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define NB_BAL_HD __host__ __device__
#else
#define NB_BAL_HD
#endif

namespace nbody {

// part[0..world]: last step's boundaries (part[0] = 0, part[world] = n); work_ms[0..world): per-rank work time.
// target[0..world] receives the wanted boundaries of the next step (before they are snapped to leaf boundaries).
NB_BAL_HD inline void rebalance_boundaries(int world, const uint32_t* part, const float* work_ms, float damping, uint32_t* target) {
	const uint32_t n = part[world];
	target[0] = 0;
	target[world] = n;
	double total = 0.0;
	bool usable = damping > 0.0f;
	for (int r = 0; r < world; ++r) {
		if (!(work_ms[r] > 0.0f) || part[r + 1] < part[r]) usable = false;
		total += work_ms[r];
	}
	if (!usable || !(total > 0.0)) {
		for (int k = 1; k < world; ++k) target[k] = part[k];
		return;
	}
	int r = 0;
	double cum = 0.0;  // cost of ranks 0..r-1
	for (int k = 1; k < world; ++k) {
		const double goal = total * k / world;
		while (r + 1 < world && cum + work_ms[r] < goal) cum += work_ms[r++];
		const double frac = (goal - cum) / work_ms[r];
		const double pos = part[r] + (frac < 0.0 ? 0.0 : frac > 1.0 ? 1.0 : frac) * (double) (part[r + 1] - part[r]);
		double want = part[k] + (double) damping * (pos - part[k]);
		if (want < target[k - 1]) want = target[k - 1];
		if (want > n) want = n;
		target[k] = (uint32_t) (want + 0.5);
	}
}

}  // namespace nbody
