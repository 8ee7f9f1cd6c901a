// Stage 1a: stable LSD radix sort of (63-bit Morton key, 32-bit index) pairs, hand-written.
// 8 passes of 8 bits. Each pass: (1) per-tile digit histograms, (2) one exclusive scan over the
// digit-major histogram table, (3) stable scatter — every warp ranks its elements with
// __match_any_sync (equal digits inside a warp-row keep lane order, rows keep row order, warps and
// tiles keep index order), so the permutation is identical to a stable CPU sort (the oracle's
// std::stable_sort): ties between equal keys keep their previous order.
// HBM traffic per pass: 8 B (histogram read) + 12 B read + 12 B write per element; the scatter sorts each tile by
// digit in shared memory first, so its global writes are contiguous runs.
#include <utility>

#include "common.cuh"
#include "merge_path.h"

namespace nbody {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                          // elements per thread
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 elements per CTA
constexpr int kSortWarps = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(uint64_t n, const uint64_t* __restrict__ keys, int shift, uint32_t nblocks,
                                                            uint32_t* __restrict__ hist) {
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t base = (uint64_t) blockIdx.x * kSortTile;
#pragma unroll
	for (int j = 0; j < kSortItems; ++j) {
		const uint64_t i = base + (uint64_t) j * kSortThreads + threadIdx.x;
		if (i < n) atomicAdd(&h[(uint32_t) (keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[(size_t) threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major: one scan gives every tile's base
}

// ---- exclusive scan of a uint32 array (three launches: tile sums, scan of the sums, apply) ----
constexpr int kScanTile = 2048;  // 256 threads x 8

__device__ __forceinline__ uint32_t block_scan_256(uint32_t v, uint32_t* ws, uint32_t& total) {
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (unsigned) d) inc += t;
	}
	if (lane == 31) ws[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t x = lane < 8 ? ws[lane] : 0u;
#pragma unroll
		for (int d = 1; d < 8; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= (unsigned) d) x += t;
		}
		if (lane < 8) ws[lane] = x;
	}
	__syncthreads();
	total = ws[7];
	const uint32_t base = w ? ws[w - 1] : 0u;
	__syncthreads();
	return base + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_sums(uint32_t n, const uint32_t* __restrict__ in, uint32_t* __restrict__ sums) {
	__shared__ uint32_t ws[8];
	const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * 8;
	uint32_t s = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) if (base + k < n) s += in[base + k];
	uint32_t total;
	block_scan_256(s, ws, total);
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_scan_top(uint32_t m, uint32_t* sums) {  // single CTA: exclusive scan of the tile sums in place
	__shared__ uint32_t ws[8];
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < m; b0 += 256) {
		const uint32_t i = b0 + threadIdx.x;
		const uint32_t v = i < m ? sums[i] : 0u;
		uint32_t total;
		const uint32_t ex = block_scan_256(v, ws, total);
		if (i < m) sums[i] = carry + ex;
		carry += total;
	}
}
__global__ void __launch_bounds__(256) k_scan_apply(uint32_t n, uint32_t* data, const uint32_t* __restrict__ sums) {
	__shared__ uint32_t ws[8];
	const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * 8;
	uint32_t v[8], s = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) { v[k] = base + k < n ? data[base + k] : 0u; s += v[k]; }
	uint32_t total;
	uint32_t run = sums[blockIdx.x] + block_scan_256(s, ws, total);
#pragma unroll
	for (int k = 0; k < 8; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
}

// ---- stable scatter ----
// The tile is first sorted by digit inside shared memory (rank = elements of the same digit that precede it in the
// tile), then written out slot by slot: consecutive threads hold consecutive slots of one digit run, so the global
// writes are contiguous runs (16 elements = 128 B of keys on average) instead of one sector per lane.
struct ScatterSmem {
	uint64_t keys[kSortTile];
	uint32_t vals[kSortTile];
	uint32_t whist[kSortWarps][256];  // per warp: running count per digit, then exclusive base inside the tile
	uint32_t gdelta[256];             // global position of tile slot s holding digit d: gdelta[d] + s
	uint32_t ws[8];
};

__global__ void __launch_bounds__(kSortThreads, 3) k_sort_scatter(uint64_t n, const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int shift,
                                                               uint32_t nblocks, const uint32_t* __restrict__ offsets) {
	extern __shared__ __align__(16) unsigned char sort_smem_raw[];
	ScatterSmem& S = *reinterpret_cast<ScatterSmem*>(sort_smem_raw);
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	for (int k = threadIdx.x; k < kSortWarps * 256; k += kSortThreads) (&S.whist[0][0])[k] = 0;
	const uint32_t gb = offsets[(size_t) threadIdx.x * nblocks + blockIdx.x];
	__syncthreads();
	// warp w owns the contiguous range [w*512, (w+1)*512) of the tile, visited row by row (32 consecutive elements per row)
	const uint64_t tbase = (uint64_t) blockIdx.x * kSortTile;
	const uint64_t wbase = tbase + (uint64_t) w * (32 * kSortItems);
	uint64_t key[kSortItems];
	uint32_t rank2[kSortItems / 2];  // two 16-bit ranks per register (a rank is < 4096)
	// all key loads first (independent requests in flight), then the ranking loop, which is serial by nature
#pragma unroll
	for (int j = 0; j < kSortItems; ++j) {
		const uint64_t i = wbase + 32 * j + lane;
		key[j] = i < n ? keys_in[i] : ~0ull;
	}
#pragma unroll
	for (int j = 0; j < kSortItems; ++j) {
		const uint64_t i = wbase + 32 * j + lane;
		const bool ok = i < n;
		const uint32_t d = (uint32_t) (key[j] >> shift) & 255u;
		const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 256u + lane);  // out-of-range lanes match nobody
		const uint32_t before = S.whist[w][d];
		const uint32_t r = before + __popc(peers & ((1u << lane) - 1u));
		if (j & 1) rank2[j >> 1] |= r << 16; else rank2[j >> 1] = r;
		__syncwarp();
		if (ok && (peers >> lane) == 1u) S.whist[w][d] = before + __popc(peers);  // the highest peer lane updates the count
		__syncwarp();
	}
	__syncthreads();
	{  // thread = digit: exclusive prefix over the warps, then over the digits of the tile
		uint32_t run = 0;
#pragma unroll
		for (int ww = 0; ww < kSortWarps; ++ww) run += S.whist[ww][threadIdx.x];
		uint32_t total;
		const uint32_t ex = block_scan_256(run, S.ws, total);
		run = ex;
#pragma unroll
		for (int ww = 0; ww < kSortWarps; ++ww) { const uint32_t c = S.whist[ww][threadIdx.x]; S.whist[ww][threadIdx.x] = run; run += c; }
		S.gdelta[threadIdx.x] = gb - ex;
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < kSortItems; ++j) {
		const uint64_t i = wbase + 32 * j + lane;
		if (i < n) {
			const uint32_t d = (uint32_t) (key[j] >> shift) & 255u;
			const uint32_t slot = S.whist[w][d] + ((j & 1) ? rank2[j >> 1] >> 16 : rank2[j >> 1] & 0xffffu);
			S.keys[slot] = key[j];
			S.vals[slot] = vals_in[i];
		}
	}
	__syncthreads();
	const uint32_t nv = (uint32_t) (n - tbase < (uint64_t) kSortTile ? n - tbase : (uint64_t) kSortTile);
#pragma unroll
	for (int j = 0; j < kSortItems; ++j) {
		const uint32_t slot = j * kSortThreads + threadIdx.x;
		if (slot < nv) {
			const uint64_t k = S.keys[slot];
			const uint32_t pos = S.gdelta[(uint32_t) (k >> shift) & 255u] + slot;
			keys_out[pos] = k;
			vals_out[pos] = S.vals[slot];
		}
	}
}

size_t own_sort_temp_bytes(uint64_t n) {
	const uint64_t nblocks = (n + kSortTile - 1) / kSortTile;
	const uint64_t hist = 256 * nblocks;
	const uint64_t sums = (hist + kScanTile - 1) / kScanTile;
	return (size_t) (hist + sums + 64) * sizeof(uint32_t);
}

// Sorts the `n` (key, index) pairs starting at element `first` of (keys[0], idx[0]) ascending by the low 63 key bits; the result
// ends in the same range of keys[0] / idx[0] (8 passes = even). keys[1] / idx[1] are scratch over the same range.
void launch_own_sort_range(Sim& s, uint64_t first, uint64_t n) {
	if (n == 0) return;
	const uint32_t nblocks = (uint32_t) ((n + kSortTile - 1) / kSortTile);
	const uint32_t hist_n = 256u * nblocks;
	const uint32_t nsums = (hist_n + kScanTile - 1) / kScanTile;
	uint32_t* hist = static_cast<uint32_t*>(s.sort_tmp);
	uint32_t* sums = hist + hist_n;
	uint64_t* kin = s.keys[0] + first; uint64_t* kout = s.keys[1] + first;
	uint32_t* vin = s.idx[0] + first; uint32_t* vout = s.idx[1] + first;
	cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(ScatterSmem));
	for (int pass = 0; pass < 8; ++pass) {
		const int shift = 8 * pass;
		k_sort_hist<<<nblocks, kSortThreads, 0, s.stream>>>(n, kin, shift, nblocks, hist);
		k_scan_sums<<<nsums, 256, 0, s.stream>>>(hist_n, hist, sums);
		k_scan_top<<<1, 256, 0, s.stream>>>(nsums, sums);
		k_scan_apply<<<nsums, 256, 0, s.stream>>>(hist_n, hist, sums);
		k_sort_scatter<<<nblocks, kSortThreads, sizeof(ScatterSmem), s.stream>>>(n, kin, vin, kout, vout, shift, nblocks, hist);
		uint64_t* tk = kin; kin = kout; kout = tk;
		uint32_t* tv = vin; vin = vout; vout = tv;
	}
}

void launch_own_sort(Sim& s) { launch_own_sort_range(s, 0, s.n); }

// Exclusive scan of data[0..n) in place (three launches); the scratch is the radix sort's histogram area, free outside the sort.
void launch_exclusive_scan(Sim& s, uint32_t* data, uint32_t n) {
	if (n == 0) return;
	const uint32_t nsums = (n + kScanTile - 1) / kScanTile;
	uint32_t* sums = static_cast<uint32_t*>(s.sort_tmp);
	k_scan_sums<<<nsums, 256, 0, s.stream>>>(n, data, sums);
	k_scan_top<<<1, 256, 0, s.stream>>>(nsums, sums);
	k_scan_apply<<<nsums, 256, 0, s.stream>>>(n, data, sums);
}

// ---- distributed sort (NBODY_FLAG_DIST_SORT): pairwise stable merges of the all-gathered, slice-wise sorted runs ----
// merge_path.h has the plan and the index arithmetic (checked on the CPU by tests/test_merge_host.py). Two kernels per round:
// k_merge_partition finds every tile's split point with one merge-path search in global memory per thread (all tiles at once: one
// latency chain of ~24 dependent loads for the whole round instead of one per CTA); k_merge_runs then streams: one CTA per tile of
// kMergeTile outputs stages its inputs in shared memory with coalesced loads, every thread merges kMergeVT outputs from its own
// split point, and the merged tile goes back through shared memory so that the global writes are contiguous.
// HBM traffic: 12 B read + 12 B written per element and round.
__global__ void __launch_bounds__(256) k_merge_partition(const MergePlan pl, uint32_t tiles, const uint64_t* __restrict__ kin, uint32_t* __restrict__ split) {
	const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
	if (tile >= tiles) return;
	const MergeTileRange r = merge_tile_range(pl, tile);
	split[tile] = merge_path(kin + r.a0, r.na, kin + r.a0 + r.na, r.nb, r.d0);
}

__global__ void __launch_bounds__(kMergeThreads) k_merge_runs(const MergePlan pl, const uint32_t* __restrict__ split, const uint64_t* __restrict__ kin,
                                                              const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout, uint32_t* __restrict__ vout) {
	__shared__ uint64_t sk[kMergeTile];
	__shared__ uint32_t sv[kMergeTile];
	const MergeTileRange r = merge_tile_range(pl, blockIdx.x);
	// this tile's split, and the next tile's (the end of the A run when this is the pair's last tile)
	const uint32_t i0 = split[blockIdx.x], i1 = r.d1 == r.na + r.nb ? r.na : split[blockIdx.x + 1];
	const uint32_t j0 = r.d0 - i0, j1 = r.d1 - i1;
	const uint32_t ca = i1 - i0, cb = j1 - j0, cnt = ca + cb;  // cnt = d1 - d0 <= kMergeTile
	for (uint32_t t = threadIdx.x; t < cnt; t += kMergeThreads) {
		const uint32_t src = t < ca ? r.a0 + i0 + t : r.a0 + r.na + j0 + (t - ca);
		sk[t] = kin[src];
		sv[t] = vin[src];
	}
	__syncthreads();
	const uint32_t dd = min(threadIdx.x * (uint32_t) kMergeVT, cnt);
	const uint32_t i = merge_path(sk, ca, sk + ca, cb, dd), j = dd - i;
	uint64_t rk[kMergeVT];
	uint32_t rv[kMergeVT];
	merge_serial(sk, sv, ca, sk + ca, sv + ca, cb, i, j, rk, rv);
	__syncthreads();
#pragma unroll
	for (int u = 0; u < kMergeVT; ++u)
		if (dd + u < cnt) { sk[dd + u] = rk[u]; sv[dd + u] = rv[u]; }
	__syncthreads();
	for (uint32_t t = threadIdx.x; t < cnt; t += kMergeThreads) {
		kout[r.a0 + r.d0 + t] = sk[t];
		vout[r.a0 + r.d0 + t] = sv[t];
	}
}

__global__ void k_iota(uint64_t n, uint32_t* __restrict__ idx) {
	const uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
	if (i < n) idx[i] = (uint32_t) i;
}

// (keys[0], idx[0]) hold `nruns` sorted runs with boundaries bound[0..nruns] (bound[nruns] = n): merge them into one sorted
// sequence. Rounds ping-pong between the two buffers; the buffer pointers are swapped at the end if needed, so the result is
// in keys[0] / idx[0] like every other sort path.
void launch_merge_runs(Sim& s, const uint32_t* bound, int nruns) {
	MergePlan pl{};
	pl.nruns = nruns;
	for (int r = 0; r <= nruns; ++r) pl.bound[r] = bound[r];
	int from = 0;
	while (pl.nruns > 1) {
		const uint32_t tiles = merge_plan_tiles(pl);
		uint32_t* split = static_cast<uint32_t*>(s.sort_tmp);  // the radix sort's histogram scratch (256 words per 4096 elements) is free by now
		if (tiles) {
			k_merge_partition<<<(tiles + 255) / 256, 256, 0, s.stream>>>(pl, tiles, s.keys[from], split);
			k_merge_runs<<<tiles, kMergeThreads, 0, s.stream>>>(pl, split, s.keys[from], s.idx[from], s.keys[from ^ 1], s.idx[from ^ 1]);
		}
		from ^= 1;
		pl = merge_plan_next(pl);
	}
	if (from) { std::swap(s.keys[0], s.keys[1]); std::swap(s.idx[0], s.idx[1]); }
}

}  // namespace nbody

using namespace nbody;

extern "C" int nbody_cuda_sort_runs(int device, const uint64_t* keys, uint64_t n, const uint32_t* bound, int nruns, uint64_t* keys_out,
                                    uint32_t* index_out) {
	if (!keys || !bound || !keys_out || !index_out || n == 0 || n > 0xfffffff0ull || nruns < 1 || nruns > kMergeMaxRuns) {
		set_error("bad argument");
		return NBODY_ERR_INVALID;
	}
	if (bound[0] != 0 || bound[nruns] != n) { set_error("run boundaries must start at 0 and end at n"); return NBODY_ERR_INVALID; }
	for (int r = 0; r < nruns; ++r)
		if (bound[r] > bound[r + 1]) { set_error("run boundaries must not decrease"); return NBODY_ERR_INVALID; }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: this library has no CPU fallback"); return NBODY_ERR_CUDA; }
	if (device >= 0) NB_CUDA_CHECK(cudaSetDevice(device));
	Sim s{};
	s.n = n;
	auto release = [&]() {
		for (int k = 0; k < 2; ++k) { if (s.keys[k]) cudaFree(s.keys[k]); if (s.idx[k]) cudaFree(s.idx[k]); }
		if (s.sort_tmp) cudaFree(s.sort_tmp);
		if (s.stream) cudaStreamDestroy(s.stream);
	};
	auto fail = [&](const char* what) { set_error(what); release(); return NBODY_ERR_CUDA; };
	if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream creation failed");
	for (int k = 0; k < 2; ++k)
		if (cudaMalloc((void**) &s.keys[k], n * sizeof(uint64_t)) != cudaSuccess || cudaMalloc((void**) &s.idx[k], n * sizeof(uint32_t)) != cudaSuccess)
			return fail("device allocation failed");
	if (cudaMalloc(&s.sort_tmp, own_sort_temp_bytes(n)) != cudaSuccess) return fail("device allocation failed");
	uint64_t* k0 = s.keys[0];
	uint32_t* i0 = s.idx[0];
	if (cudaMemcpyAsync(k0, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream) != cudaSuccess) return fail("upload failed");
	k_iota<<<(unsigned) ((n + 255) / 256), 256, 0, s.stream>>>(n, i0);
	for (int r = 0; r < nruns; ++r) launch_own_sort_range(s, bound[r], bound[r + 1] - bound[r]);  // what each rank does for its slice
	launch_merge_runs(s, bound, nruns);                                                          // what every rank does after the all-gather
	if (cudaMemcpyAsync(keys_out, s.keys[0], n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
	    cudaMemcpyAsync(index_out, s.idx[0], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
	    cudaStreamSynchronize(s.stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
		return fail("distributed-sort pipeline failed on the device");
	release();
	return NBODY_OK;
}

