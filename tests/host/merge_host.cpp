// Host build of nbody_b200/csrc/merge_path.h for CPU-side unit tests (g++): runs the distributed sort's merge rounds exactly as
// the kernels k_merge_partition / k_merge_runs (sort.cu) decompose them — tile ranges from the plan, one merge-path search per tile
// in the global arrays, staging, one merge-path search and one serial merge per thread — with loops in place of CTAs and threads.
#include <algorithm>
#include <cstdint>
#include <vector>
#include "../../nbody_b200/csrc/merge_path.h"

using namespace nbody;

static void merge_round(const MergePlan& pl, uint32_t tiles, const uint64_t* kin, const uint32_t* vin, uint64_t* kout, uint32_t* vout) {
	std::vector<uint64_t> sk(kMergeTile);
	std::vector<uint32_t> sv(kMergeTile);
	std::vector<uint32_t> split(tiles);
	for (uint32_t tile = 0; tile < tiles; ++tile) {  // k_merge_partition: one search per tile
		const MergeTileRange r = merge_tile_range(pl, tile);
		split[tile] = merge_path(kin + r.a0, r.na, kin + r.a0 + r.na, r.nb, r.d0);
	}
	for (uint32_t tile = 0; tile < tiles; ++tile) {  // k_merge_runs
		const MergeTileRange r = merge_tile_range(pl, tile);
		const uint32_t i0 = split[tile], i1 = r.d1 == r.na + r.nb ? r.na : split[tile + 1];
		const uint32_t j0 = r.d0 - i0, j1 = r.d1 - i1, ca = i1 - i0, cb = j1 - j0, cnt = ca + cb;
		for (uint32_t t = 0; t < cnt; ++t) {
			const uint32_t src = t < ca ? r.a0 + i0 + t : r.a0 + r.na + j0 + (t - ca);
			sk[t] = kin[src]; sv[t] = vin[src];
		}
		for (uint32_t th = 0; th < (uint32_t) kMergeThreads; ++th) {
			const uint32_t dd = th * kMergeVT < cnt ? th * kMergeVT : cnt;
			const uint32_t i = merge_path(sk.data(), ca, sk.data() + ca, cb, dd), j = dd - i;
			uint64_t rk[kMergeVT]; uint32_t rv[kMergeVT];
			merge_serial(sk.data(), sv.data(), ca, sk.data() + ca, sv.data() + ca, cb, i, j, rk, rv);
			for (int u = 0; u < kMergeVT; ++u)
				if (dd + u < cnt) { kout[r.a0 + r.d0 + dd + u] = rk[u]; vout[r.a0 + r.d0 + dd + u] = rv[u]; }
		}
	}
}

extern "C" {
int merge_tile_size() { return kMergeTile; }

// keys/vals hold `nruns` sorted runs with boundaries bound[0..nruns]; merges them in place (ping-pong through scratch).
// Returns the number of rounds.
int merge_all_runs(uint64_t* keys, uint32_t* vals, uint64_t n, const uint32_t* bound, int nruns) {
	std::vector<uint64_t> k2(n);
	std::vector<uint32_t> v2(n);
	uint64_t* kin = keys; uint64_t* kout = k2.data();
	uint32_t* vin = vals; uint32_t* vout = v2.data();
	MergePlan pl{};
	pl.nruns = nruns;
	for (int r = 0; r <= nruns; ++r) pl.bound[r] = bound[r];
	int rounds = 0;
	while (pl.nruns > 1) {
		const uint32_t tiles = merge_plan_tiles(pl);
		merge_round(pl, tiles, kin, vin, kout, vout);
		std::swap(kin, kout); std::swap(vin, vout);
		pl = merge_plan_next(pl);
		++rounds;
	}
	if (kin != keys) { std::copy(kin, kin + n, keys); std::copy(vin, vin + n, vals); }
	return rounds;
}
}
