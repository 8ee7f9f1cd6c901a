// Stage 1: Morton keys, device radix sort, state permutation and the linear
// octree, all on the device with no host round trip.
//
// Replaces the reference's host-side glade::Orthtree build and per-step
// re-bucketing (src/open_cl_simulation.cpp:15-51,108-124) and the per-step
// upload of the whole leaf/node arrays (:128-135). The octree obeys the contract
// the reference consumes (split above `capacity`, all 8 children
// exist, contiguous particle range per node) but is stored LEVEL-MAJOR with the
// 8 children of a node contiguous, which is what the traversal / M2L kernels
// want; nbody_cuda_get_tree() re-emits it in the reference's DFS pre-order.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace nbody {

// ---------------------------------------------------------------------------
// AoS48 boundary records <-> SoA float4 planes
// ---------------------------------------------------------------------------
__global__ void k_import(uint64_t n, const float4* __restrict__ aos, float4* __restrict__ posq, float4* __restrict__ velm,
                         uint32_t* __restrict__ orig, uint32_t first_id) {
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
		const float4 p = aos[3 * i], v = aos[3 * i + 1], mq = aos[3 * i + 2];
		posq[i] = make_float4(p.x, p.y, p.z, mq.y);
		velm[i] = make_float4(v.x, v.y, v.z, mq.x);
		if (orig) orig[i] = first_id + (uint32_t) i;
	}
}
__global__ void k_export(uint64_t n, float4* __restrict__ aos, const float4* __restrict__ posq, const float4* __restrict__ velm) {
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
		const float4 p = posq[i], v = velm[i];
		aos[3 * i] = make_float4(p.x, p.y, p.z, 0.0f);
		aos[3 * i + 1] = make_float4(v.x, v.y, v.z, 0.0f);
		aos[3 * i + 2] = make_float4(v.w, p.w, 0.0f, 0.0f);
	}
}
__global__ void k_keys(uint64_t n, const float4* __restrict__ posq, float sx, float sy, float sz, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx);
__global__ void k_gather(uint64_t n, const uint32_t* __restrict__ idx, const float4* __restrict__ posq_in, const float4* __restrict__ velm_in,
                         const uint32_t* __restrict__ orig_in, float4* __restrict__ posq_out, float4* __restrict__ velm_out, uint32_t* __restrict__ orig_out);
static inline int grid_for(uint64_t n, int block) {
	const uint64_t want = (n + block - 1) / block;
	const uint64_t cap = (uint64_t) kNumSM * 16;
	return (int) (want < 1 ? 1 : (want > cap ? cap : want));
}
void launch_import(Sim& s, const nbody_particle* aos_dev, uint64_t n) {
	k_import<<<grid_for(n, 256), 256, 0, s.stream>>>(n, (const float4*) aos_dev, s.posq[0], s.velm[0], s.orig[0], 0u);
}
// partitioned mode: identities start at `first_id` (the rank's offset into the constructor's global array); keep_ids leaves orig alone
void launch_import_ids(Sim& s, const nbody_particle* aos_dev, uint64_t n, uint32_t first_id, bool keep_ids) {
	k_import<<<grid_for(n, 256), 256, 0, s.stream>>>(n, (const float4*) aos_dev, s.posq[0], s.velm[0], keep_ids ? nullptr : s.orig[0], first_id);
}
void launch_keys(Sim& s, const float4* pos, uint64_t n) {  // keys[0] = Morton keys of pos[0..n), idx[0] = 0..n-1
	const float sx = 2097152.0f / s.cfg.bounds[0], sy = 2097152.0f / s.cfg.bounds[1], sz = 2097152.0f / s.cfg.bounds[2];
	if (n) k_keys<<<grid_for(n, 256), 256, 0, s.stream>>>(n, pos, sx, sy, sz, s.keys[0], s.idx[0]);
}
void launch_gather(Sim& s, uint64_t n) {  // (posq, velm, orig)[1][i] = (posq, velm, orig)[0][idx[0][i]]
	if (n) k_gather<<<grid_for(n, 256), 256, 0, s.stream>>>(n, s.idx[0], s.posq[0], s.velm[0], s.orig[0], s.posq[1], s.velm[1], s.orig[1]);
}
void launch_export(Sim& s, nbody_particle* aos_dev, uint64_t n) {
	k_export<<<grid_for(n, 256), 256, 0, s.stream>>>(n, (float4*) aos_dev, s.posq[0], s.velm[0]);
}

// ---------------------------------------------------------------------------
// Morton keys. 21 bits per dimension, digit = x | y<<1 | z<<2, one FP32 multiply
// per coordinate (no FMA contraction) so the host oracle classifies identically.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread3(uint32_t v) {
	uint64_t x = v & 0x1fffffu;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}
__device__ __forceinline__ uint32_t compact3(uint64_t x) {
	x &= 0x1249249249249249ull;
	x = (x | x >> 2) & 0x10c30c30c30c30c3ull;
	x = (x | x >> 4) & 0x100f00f00f00f00full;
	x = (x | x >> 8) & 0x1f0000ff0000ffull;
	x = (x | x >> 16) & 0x1f00000000ffffull;
	x = (x | x >> 32) & 0x1fffffull;
	return (uint32_t) x;
}
__device__ __forceinline__ uint32_t quantise(float x, float scale) {
	const float v = fminf(fmaxf(__fmul_rn(x, scale), 0.0f), 2097151.0f);
	return __float2uint_rz(v);
}
__global__ void k_keys(uint64_t n, const float4* __restrict__ posq, float sx, float sy, float sz, uint64_t* __restrict__ keys,
                       uint32_t* __restrict__ idx) {
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
		const float4 p = posq[i];
		keys[i] = spread3(quantise(p.x, sx)) | spread3(quantise(p.y, sy)) << 1 | spread3(quantise(p.z, sz)) << 2;
		idx[i] = (uint32_t) i;
	}
}
// the same for the elements [first, first + count) only (distributed sort)
__global__ void k_keys_range(uint64_t first, uint64_t count, const float4* __restrict__ posq, float sx, float sy, float sz,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
	for (uint64_t t = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; t < count; t += (uint64_t) gridDim.x * blockDim.x) {
		const uint64_t i = first + t;
		const float4 p = posq[i];
		keys[i] = spread3(quantise(p.x, sx)) | spread3(quantise(p.y, sy)) << 1 | spread3(quantise(p.z, sz)) << 2;
		idx[i] = (uint32_t) i;
	}
}
__global__ void k_gather(uint64_t n, const uint32_t* __restrict__ idx, const float4* __restrict__ posq_in,
                         const float4* __restrict__ velm_in, const uint32_t* __restrict__ orig_in, float4* __restrict__ posq_out,
                         float4* __restrict__ velm_out, uint32_t* __restrict__ orig_out) {
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t j = idx[i];
		posq_out[i] = posq_in[j];
		velm_out[i] = velm_in[j];
		orig_out[i] = orig_in[j];
	}
}

// Distributed runs gather in two halves: positions (+ identity) right after the sort, velocities only before the leaf
// kernel, because the other ranks' velocities of the previous step may still be travelling on the second stream (comm.cu).
__global__ void k_gather_pos(uint64_t n, const uint32_t* __restrict__ idx, const float4* __restrict__ posq_in,
                             const uint32_t* __restrict__ orig_in, float4* __restrict__ posq_out, uint32_t* __restrict__ orig_out) {
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t j = idx[i];
		posq_out[i] = posq_in[j];
		orig_out[i] = orig_in[j];
	}
}
__global__ void k_gather_vel(uint64_t n, const uint32_t* __restrict__ idx, const float4* __restrict__ velm_in, float4* __restrict__ velm_out) {
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) velm_out[i] = velm_in[idx[i]];
}
void launch_gather_velocities(Sim& s) {
	if (!s.comm) return;  // single GPU: k_gather already moved them
	k_gather_vel<<<grid_for(s.n, 256), 256, 0, s.stream>>>(s.n, s.idx[0], s.velm[0], s.velm[1]);
}

size_t own_sort_temp_bytes(uint64_t n);  // sort.cu
void launch_own_sort(Sim& s);            // sort.cu

size_t sort_temp_bytes(uint64_t n) {
	size_t bytes = 0;
	cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
	cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int64_t) n, 0, 63);
	return bytes > own_sort_temp_bytes(n) ? bytes : own_sort_temp_bytes(n);
}

// keys of state order -> stable radix sort -> gather the state into sorted order (posq[1], velm[1], orig[1]).
// After this call keys[0] holds the sorted keys.
int launch_keys_sort_permute(Sim& s) {
	const uint64_t n = s.n;
	const float sx = 2097152.0f / s.cfg.bounds[0], sy = 2097152.0f / s.cfg.bounds[1], sz = 2097152.0f / s.cfg.bounds[2];
	if (s.comm && (s.cfg.flags & NBODY_FLAG_DIST_SORT) && !(s.cfg.flags & NBODY_FLAG_CUB_SORT)) {
		// Distributed sort: the state is in the previous step's tree order and this rank's slice of it is [first, first + count).
		// Keys and radix sort for that slice only (idx = index into the whole state array), all-gather of the sorted runs,
		// pairwise stable merges. Run r holds lower previous indices than run r + 1, so the merged permutation is the stable
		// sort of all keys, bit for bit what the replicated sort below computes (tests/test_merge_host.py).
		uint64_t first = 0, count = 0;
		comm_own_slice(s, &first, &count);
		if (count) k_keys_range<<<grid_for(count, 256), 256, 0, s.stream>>>(first, count, s.posq[0], sx, sy, sz, s.keys[0], s.idx[0]);
		launch_own_sort_range(s, first, count);
		uint32_t bound[kMaxRanks + 1];
		int nruns = 0;
		const int rc = comm_sort_exchange(s, bound, &nruns);
		if (rc) return rc;
		launch_merge_runs(s, bound, nruns);
		const int rcw = comm_wait_positions(s);  // the other ranks' positions may still be travelling on the second stream
		if (rcw) return rcw;
		k_gather_pos<<<grid_for(n, 256), 256, 0, s.stream>>>(n, s.idx[0], s.posq[0], s.orig[0], s.posq[1], s.orig[1]);
		return NBODY_OK;
	}
	k_keys<<<grid_for(n, 256), 256, 0, s.stream>>>(n, s.posq[0], sx, sy, sz, s.keys[0], s.idx[0]);
	if (s.cfg.flags & NBODY_FLAG_CUB_SORT) {
		// comparison path only: CUB's onesweep sort, the bar the hand-written sort is measured against
		cub::DoubleBuffer<uint64_t> k(s.keys[0], s.keys[1]);
		cub::DoubleBuffer<uint32_t> v(s.idx[0], s.idx[1]);
		size_t bytes = s.sort_tmp_bytes;
		NB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(s.sort_tmp, bytes, k, v, (int64_t) n, 0, 63, s.stream));
		if (k.Current() != s.keys[0]) { std::swap(s.keys[0], s.keys[1]); }
		if (v.Current() != s.idx[0]) { std::swap(s.idx[0], s.idx[1]); }
	} else {
		launch_own_sort(s);
	}
	if (s.comm) k_gather_pos<<<grid_for(n, 256), 256, 0, s.stream>>>(n, s.idx[0], s.posq[0], s.orig[0], s.posq[1], s.orig[1]);
	else k_gather<<<grid_for(n, 256), 256, 0, s.stream>>>(n, s.idx[0], s.posq[0], s.velm[0], s.orig[0], s.posq[1], s.velm[1], s.orig[1]);
	return NBODY_OK;
}

// ---------------------------------------------------------------------------
// Linear octree, one level per pair of launches, sizes read from the device
// control block (the host never learns the node count during a step).
// ---------------------------------------------------------------------------
__global__ void k_tree_init(Ctrl* c, uint32_t n, float bx, float by, float bz, float4* geom, uint2* info, uint32_t* nbegin,
                            uint32_t* nparent, uint64_t* nkey, uint32_t* p2p_head) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		for (int l = 0; l < kNumLevels + 2; ++l) c->level_off[l] = l == 0 ? 0u : 1u;
		c->n_nodes = 1; c->status = 0; c->scan_ticket = 0; c->n_levels = 1;
		c->gq_count[0] = c->gq_count[1] = 0; c->items_count[0] = c->items_count[1] = 0;
		c->seg_cursor = 0; c->near_cursor[0] = c->near_cursor[1] = 0; c->p2p_cursor = 0; c->m2l_cursor = 0;
		c->stat_m2l_inter = c->stat_m2l_low = c->stat_p2p_entries = c->stat_p2p_inter = c->stat_near = c->stat_leaves = 0;
		for (int k = 0; k < 4; ++k) c->work_ticket[k] = 0;
		c->acc_max2_bits = 0;
		c->n_leaf_items = 0;
		c->part[0] = 0;
		for (int k = 1; k <= kMaxRanks; ++k) c->part[k] = n;  // single GPU: rank 0 owns everything (k_partition overwrites this)
		geom[0] = make_float4(__fadd_rn(__fmul_rn(0.0f, bx), __fmul_rn(bx, 0.5f)), __fadd_rn(__fmul_rn(0.0f, by), __fmul_rn(by, 0.5f)),
		                      __fadd_rn(__fmul_rn(0.0f, bz), __fmul_rn(bz, 0.5f)), bx);
		info[0] = make_uint2(0u, n);
		nbegin[0] = 0; nparent[0] = 0; nkey[0] = 0; p2p_head[0] = 0xffffffffu;
	}
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (unsigned) d) inc += t;
	}
	if (lane == 31) warp_sums[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t ws = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0u;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d);
			if (lane >= (unsigned) d) ws += t;
		}
		warp_sums[lane] = ws;  // inclusive over warps
	}
	__syncthreads();
	total = warp_sums[(blockDim.x >> 5) - 1];
	const uint32_t base = w == 0 ? 0u : warp_sums[w - 1];
	__syncthreads();
	return base + inc - v;
}

// Does a node of level `l` with `count` particles and key prefix `key` split? More than `cap` particles — or, in partitioned mode, a
// cell that holds particles of several ranks (it contains a splitter key strictly inside its key range) whose GLOBAL count exceeds
// cap (LetCtrl::force, summed over the ranks before the build): every rank's tree is then the global octree restricted to the
// cells that hold its own particles.
__device__ __forceinline__ bool node_splits(uint32_t count, uint32_t cap, uint64_t key, int l, const LetCtrl* __restrict__ lc) {
	if (count > cap) return true;
	if (lc == nullptr || count == 0u) return false;
	const int sh = 3 * (kMaxDepth - l);
	for (int b = 1; b < lc->world; ++b)
		if (lc->force[b][l] > cap && (lc->split[b] >> sh) == (key >> sh)) return true;
	return false;
}

// Pass A of level `l`: count the nodes that split in each tile; the last block to finish
// turns the tile counts into exclusive offsets (a block-wide scan) and publishes the size of level l+1.
// A level without nodes (below the deepest one) costs one launch that only clears the split count.
__global__ void __launch_bounds__(256) k_level_count(Ctrl* c, int l, uint32_t cap, uint32_t max_depth, uint32_t max_nodes,
                                                      const uint2* __restrict__ info, const uint64_t* __restrict__ nkey,
                                                      const LetCtrl* __restrict__ lc, uint32_t* scan_sums) {
	__shared__ uint32_t warp_sums[32];
	__shared__ bool last;
	const uint32_t lo = c->level_off[l], hi = c->level_off[l + 1];
	const uint32_t nl = hi - lo;
	if (nl == 0 || (uint32_t) l >= max_depth) {  // uniform over the grid: nothing can split here
		if (blockIdx.x == 0 && threadIdx.x == 0) scan_sums[gridDim.x] = 0;
		return;
	}
	const uint32_t tile = (nl + gridDim.x - 1) / gridDim.x;
	const uint32_t t0 = lo + blockIdx.x * tile, t1 = min(hi, t0 + tile);
	uint32_t cnt = 0;
	for (uint32_t i = t0 + threadIdx.x; i < t1; i += blockDim.x) cnt += node_splits(info[i].y, cap, lc ? nkey[i] : 0ull, l, lc) ? 1u : 0u;
	uint32_t total;
	block_exclusive_scan(cnt, warp_sums, total);
	if (threadIdx.x == 0) {
		scan_sums[blockIdx.x] = total;
		__threadfence();
		last = atomicAdd(&c->scan_ticket, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (!last) return;
	__threadfence();
	// exclusive scan of the gridDim.x tile counts by the whole block: thread t owns counts [t*per, (t+1)*per)
	constexpr uint32_t kPerMax = (kScanBlocks + 255) / 256;
	const uint32_t per = (gridDim.x + blockDim.x - 1) / blockDim.x;
	uint32_t v[kPerMax], mine = 0;
#pragma unroll
	for (uint32_t k = 0; k < kPerMax; ++k) {
		const uint32_t b = threadIdx.x * per + k;
		v[k] = (k < per && b < gridDim.x) ? ((volatile uint32_t*) scan_sums)[b] : 0u;
		mine += v[k];
	}
	uint32_t run;
	uint32_t ex = block_exclusive_scan(mine, warp_sums, run);
#pragma unroll
	for (uint32_t k = 0; k < kPerMax; ++k) {
		const uint32_t b = threadIdx.x * per + k;
		if (k < per && b < gridDim.x) scan_sums[b] = ex;
		ex += v[k];
	}
	if (threadIdx.x == 0) {
		uint32_t next = hi + 8u * run;
		if (next > max_nodes || next < hi) { atomicOr(&c->status, kOvfNodes); run = 0; next = hi; }
		scan_sums[gridDim.x] = run;
		for (int k = l + 2; k < kNumLevels + 2; ++k) c->level_off[k] = next;
		c->n_nodes = next;
		if (run) c->n_levels = l + 2;
		c->scan_ticket = 0;
	}
}

// Pass B of level `l`: every splitting node gets its 8 children at level_off[l+1] + 8*rank,
// rank = its position among the splitting nodes of the level (deterministic, Morton order).
// Eight lanes per node, lane k = child k: the seven child boundaries are found by seven concurrent binary
// searches in the sorted keys (one thread searching them one after the other made every level cost 60-90 us of
// dependent-load latency, even the top ones with a handful of nodes) and the eight child records are written by
// eight consecutive lanes.
constexpr int kSplitNodes = 256 / 8;  // nodes per block iteration
__global__ void __launch_bounds__(256) k_level_split(Ctrl* c, int l, uint32_t cap, uint32_t max_depth, float bx, float by, float bz,
                                                      const uint64_t* __restrict__ keys, float4* geom, uint2* info, uint32_t* nbegin,
                                                      uint32_t* nparent, uint64_t* nkey, uint32_t* p2p_head,
                                                      const LetCtrl* __restrict__ lc, const uint32_t* __restrict__ scan_sums) {
	__shared__ uint32_t warp_sums[32];
	if ((uint32_t) l >= max_depth || scan_sums[gridDim.x] == 0) return;
	const uint32_t lo = c->level_off[l], hi = c->level_off[l + 1];
	const uint32_t nl = hi - lo;
	const uint32_t tile = (nl + gridDim.x - 1) / gridDim.x;
	const uint32_t t0 = lo + blockIdx.x * tile, t1 = min(hi, t0 + tile);
	uint32_t running = scan_sums[blockIdx.x];
	const int shift = 3 * (kMaxDepth - 1 - l);
	const float sc = __int_as_float((127 - (l + 1)) << 23);  // 2^-(l+1), exact
	const float dx = __fmul_rn(bx, sc), dy = __fmul_rn(by, sc), dz = __fmul_rn(bz, sc);
	const uint32_t k = threadIdx.x & 7u, slot = threadIdx.x >> 3;
	for (uint32_t base = t0; base < t1; base += kSplitNodes) {  // uniform trip count inside the block
		const uint32_t i = base + slot;
		uint2 nf = make_uint2(0u, 0u);
		if (i < t1) nf = info[i];
		const bool split = i < t1 && node_splits(nf.y, cap, lc ? nkey[i] : 0ull, l, lc);
		uint32_t total;
		uint32_t rank = block_exclusive_scan(split && k == 0u ? 1u : 0u, warp_sums, total);
		rank = __shfl_sync(0xffffffffu, rank, (threadIdx.x & 31u) & ~7u);  // the node's lane 0 holds its rank
		uint32_t b = 0, e = 0, end_k = 0;
		uint64_t pk = 0;
		if (split) {
			b = nbegin[i]; e = b + nf.y; pk = nkey[i];
			// first particle whose digit at this level exceeds k
			end_k = e;
			if (k < 7u) {
				const uint64_t bound = pk | (uint64_t) (k + 1u) << shift;
				uint32_t a = b, z = e;
				while (a < z) { const uint32_t m = a + ((z - a) >> 1); if (keys[m] < bound) a = m + 1; else z = m; }
				end_k = a;
			}
		}
		uint32_t prev = __shfl_up_sync(0xffffffffu, end_k, 1);  // child k starts where child k-1 ends
		if (k == 0u) prev = b;
		if (split) {
			const uint32_t cb = hi + 8u * (running + rank);
			if (k == 0u) info[i] = make_uint2(cb, nf.y);
			const uint64_t pp = l == 0 ? 0ull : pk >> (shift + 3);  // parent's digits, last one in bits 0..2
			const uint32_t pix = compact3(pp), piy = compact3(pp >> 1), piz = compact3(pp >> 2);
			const uint32_t cid = cb + k;
			const uint32_t ix = pix << 1 | (k & 1u), iy = piy << 1 | (k >> 1 & 1u), iz = piz << 1 | (k >> 2 & 1u);
			geom[cid] = make_float4(__fadd_rn(__fmul_rn((float) ix, dx), __fmul_rn(dx, 0.5f)),
			                        __fadd_rn(__fmul_rn((float) iy, dy), __fmul_rn(dy, 0.5f)),
			                        __fadd_rn(__fmul_rn((float) iz, dz), __fmul_rn(dz, 0.5f)), dx);
			info[cid] = make_uint2(0u, end_k - prev);
			nbegin[cid] = prev;
			nparent[cid] = i;
			nkey[cid] = pk | (uint64_t) k << shift;
			p2p_head[cid] = 0xffffffffu;
		}
		running += total;
	}
}

// The level loops of a step stop at Sim::depth_bound (last step's depth + 1) instead of max_depth: on the Plummer benchmark the tree
// has 12 levels of 21, and the launches for the empty ones were a third of all launches of a step. If a node at the bound still wants
// to split, kOvfDepth makes the host re-run the step unbounded.
__global__ void k_level_check(Ctrl* c, int l, uint32_t cap, const uint2* __restrict__ info, const uint64_t* __restrict__ nkey,
                              const LetCtrl* __restrict__ lc) {
	const uint32_t lo = c->level_off[l], hi = c->level_off[l + 1];
	bool any = false;
	for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x)
		any |= node_splits(info[i].y, cap, nkey[i], l, lc);
	if (any) atomicOr(&c->status, kOvfDepth);
}

void launch_tree_build(Sim& s) {
	const nbody_cuda_config& cf = s.cfg;
	const int bound = s.depth_bound < (int) cf.max_depth ? s.depth_bound : (int) cf.max_depth;
	k_tree_init<<<1, 32, 0, s.stream>>>(s.ctrl, (uint32_t) s.n, cf.bounds[0], cf.bounds[1], cf.bounds[2], s.geom, s.info, s.nbegin,
	                                     s.nparent, s.nkey, s.p2p_head);
	for (int l = 0; l < bound; ++l) {
		k_level_count<<<kScanBlocks, 256, 0, s.stream>>>(s.ctrl, l, cf.leaf_capacity, cf.max_depth, s.max_nodes, s.info, s.nkey, s.let_ctrl,
		                                                  s.scan_sums);
		k_level_split<<<kScanBlocks, 256, 0, s.stream>>>(s.ctrl, l, cf.leaf_capacity, cf.max_depth, cf.bounds[0], cf.bounds[1],
		                                                  cf.bounds[2], s.keys[0], s.geom, s.info, s.nbegin, s.nparent, s.nkey,
		                                                  s.p2p_head, s.let_ctrl, s.scan_sums);
	}
	if (bound < (int) cf.max_depth) k_level_check<<<kNumSM, 256, 0, s.stream>>>(s.ctrl, bound, cf.leaf_capacity, s.info, s.nkey, s.let_ctrl);
}

}  // namespace nbody
