// Stable two-way merge of sorted (key, value) runs by merge path, shared between the device kernel (sort.cu: k_merge_runs)
// and the CPU test (tests/host/merge_host.cpp), so that the index arithmetic is checked without a GPU.
//
// Used by the distributed sort (NBODY_FLAG_DIST_SORT, comm.cu): every rank radix-sorts only its own slice of the previous
// step's tree order, the sorted runs are all-gathered, and log2(ranks) rounds of pairwise merges produce the global order.
// Run r holds the particles with previous indices [part[r], part[r+1]), so "run A before run B on equal keys" is exactly the
// tie rule of a stable sort of all keys: the result is bit-identical to the replicated radix sort and to the oracle's
// std::stable_sort.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define NB_MP_HD __host__ __device__ __forceinline__
#else
#define NB_MP_HD inline
#endif

namespace nbody {

constexpr int kMergeThreads = 256;
constexpr int kMergeVT = 8;                               // outputs per thread
constexpr int kMergeTile = kMergeThreads * kMergeVT;      // outputs per CTA
constexpr int kMergeMaxRuns = 16;

// Number of elements of A among the first d outputs of the stable merge of A (na elements) and B (nb); A wins ties.
template <typename K>
NB_MP_HD uint32_t merge_path(const K* A, uint32_t na, const K* B, uint32_t nb, uint32_t d) {
	uint32_t lo = d > nb ? d - nb : 0u, hi = d < na ? d : na;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (A[mid] <= B[d - 1u - mid]) lo = mid + 1u;  // A[mid] precedes B[d-1-mid]: it is among the first d
		else hi = mid;
	}
	return lo;
}

// The next kMergeVT outputs of the merge of A[0..ca) and B[0..cb) from cursor (i, j); slots past the end are left untouched.
template <typename K, typename V>
NB_MP_HD void merge_serial(const K* A, const V* VA, uint32_t ca, const K* B, const V* VB, uint32_t cb, uint32_t i, uint32_t j, K (&rk)[kMergeVT],
                           V (&rv)[kMergeVT]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int u = 0; u < kMergeVT; ++u) {
		const bool ha = i < ca, hb = j < cb;
		if (!ha && !hb) break;
		const K ka = A[ha ? i : 0u], kb = B[hb ? j : 0u];
		const bool take_a = ha && (!hb || ka <= kb);
		rk[u] = take_a ? ka : kb;
		rv[u] = take_a ? VA[i] : VB[hb ? j : 0u];
		if (take_a) ++i; else ++j;
	}
}

// One round of pairwise merges over the whole array: run boundaries bound[0..nruns], pair p = runs (2p, 2p+1) (the last run of
// an odd count is copied), tiles numbered pair by pair: pair p owns tiles [tile_first[p], tile_first[p+1]).
struct MergePlan {
	uint32_t bound[kMergeMaxRuns + 1];
	uint32_t tile_first[kMergeMaxRuns / 2 + 1];
	int nruns, npairs;
};

inline uint32_t merge_plan_tiles(MergePlan& pl) {
	pl.npairs = (pl.nruns + 1) / 2;
	uint32_t t = 0;
	for (int p = 0; p < pl.npairs; ++p) {
		pl.tile_first[p] = t;
		const int e = 2 * p + 2 < pl.nruns ? 2 * p + 2 : pl.nruns;
		const uint32_t len = pl.bound[e] - pl.bound[2 * p];
		t += (len + kMergeTile - 1) / kMergeTile;
	}
	pl.tile_first[pl.npairs] = t;
	return t;
}

// The plan of the next round: every pair has become one run.
inline MergePlan merge_plan_next(const MergePlan& pl) {
	MergePlan nx{};
	nx.nruns = pl.npairs;
	for (int p = 0; p < pl.npairs; ++p) nx.bound[p] = pl.bound[2 * p];
	nx.bound[nx.nruns] = pl.bound[pl.nruns];
	return nx;
}

// What tile `tile` of the plan works on: inputs A = [a0, a0+na), B = [a0+na, a0+na+nb), output diagonals [d0, d1) of the pair,
// written to [a0+d0, a0+d1).
struct MergeTileRange { uint32_t a0, na, nb, d0, d1; };
NB_MP_HD MergeTileRange merge_tile_range(const MergePlan& pl, uint32_t tile) {
	int p = 0;
	while (p + 1 < pl.npairs && tile >= pl.tile_first[p + 1]) ++p;
	const int m = 2 * p + 1 < pl.nruns ? 2 * p + 1 : pl.nruns, e = 2 * p + 2 < pl.nruns ? 2 * p + 2 : pl.nruns;
	MergeTileRange r;
	r.a0 = pl.bound[2 * p];
	r.na = pl.bound[m] - r.a0;
	r.nb = pl.bound[e] - pl.bound[m];
	r.d0 = (tile - pl.tile_first[p]) * (uint32_t) kMergeTile;
	const uint32_t len = r.na + r.nb;
	r.d1 = r.d0 + (uint32_t) kMergeTile < len ? r.d0 + (uint32_t) kMergeTile : len;
	return r;
}

}  // namespace nbody
