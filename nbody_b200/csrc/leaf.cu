// Stage 5: near field. One warp per block of <= kLeafG targets of one leaf: the leaf's P2P
// source list is expanded into a per-warp shared-memory tile of source particles (x,y,z,q);
// every lane owns one source of a 32-source row and applies it to all targets of the block
// (target coordinates are broadcast from shared memory, the 3 x G partial accelerations live
// in registers), so all 32 lanes do useful work whatever the leaf population is. One
// transposing shuffle reduction per block, then the epilogue adds the far field (L2P) and
// applies the integrator, so accelerations never make a round trip through a per-interaction
// buffer.
//
// Replaces src/field.cl:49-148 (8x8 work-group per leaf interaction writing one
// 16-byte slot per (leaf, partner leaf, interaction)), src/force.cl:21-81 (slot
// reductions), the CPU prefix sums at src/open_cl_simulation.cpp:371-417 and the
// serial host integration loop at :572-616.
// Field of source j on target i: q_j (x_j - x_i) / (|x_j - x_i|^2 + eps^2)^(3/2)
// (src/field.cl:17-32 with FORCE_CONSTANT folded into force_constant).
// 20 flop per evaluation by the SURVEY 8d convention: 3 FADD, 3 FFMA, MUFU.RSQ (2), 3 FMUL, 3 FFMA.
#include <cstdlib>

#include "common.cuh"

namespace nbody {

constexpr int kLeafWarps = 4;
constexpr int kLeafTile = 256;  // source particles per tile; two tiles (8 KB) per warp
constexpr int kLeafChunk = 8;   // node ids per work ticket
#ifndef NBODY_LEAF_G
#define NBODY_LEAF_G 16
#endif
constexpr int kLeafG = NBODY_LEAF_G;  // most targets per block (3 accumulators each, in registers)
#ifndef NBODY_LEAF_ROWS
#define NBODY_LEAF_ROWS 4
#endif
constexpr int kLeafRows = NBODY_LEAF_ROWS;  // 32-source rows in flight per lane (independent dependency chains per target)
constexpr uint32_t kLeafPad = 32u * kLeafRows;  // tiles are padded with zero-charge sources to a multiple of this
#ifndef NBODY_LEAF_MIN_CTAS
#define NBODY_LEAF_MIN_CTAS 4
#endif
constexpr int kLeafMinCtas = NBODY_LEAF_MIN_CTAS;  // resident CTAs per SM the register allocation is held to
static_assert(kLeafG >= 1 && kLeafG <= 16, "the transposing reduction handles up to 16 targets per block");

__device__ __forceinline__ void leaf_cp_async16(void* smem, const void* gmem) {
	const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void leaf_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void leaf_cp_async_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void leaf_cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- tile fill with TMA 1-D bulk copies (default; -DNBODY_LEAF_BULK=0 builds the cp.async row fill it replaced, for A/B runs) ----
// One cp.async.bulk per contiguous run of source particles (adjacent list entries are merged: 57 particles = 910 B per run on the
// Plummer benchmark, tests/tools/p2p_list_structure.py) with completion on a per-warp, per-buffer mbarrier, instead of eight rows of
// per-lane 16-byte cp.async copies driven by a flat-slot -> entry bitmap. Measured (profiles/r02a_call.log, Plummer 2^24, capacity 48):
// leaf kernel 38.41 -> 35.12 ms, 41.4 -> 45.3 % of the FP32 FMA peak, parity suite green.
#ifndef NBODY_LEAF_BULK
#define NBODY_LEAF_BULK 1
#endif
#ifndef NBODY_LEAF_FLAT
#define NBODY_LEAF_FLAT 0   // 1: the leaf's segment chain is read as one flat entry range (tiles run across segment ends); measured, see DESIGN.md section 10
#endif
__device__ __forceinline__ unsigned leaf_smem_addr(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void leaf_mbar_init(uint64_t* bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(leaf_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void leaf_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(leaf_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void leaf_mbar_wait(uint64_t* bar, unsigned parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "LEAF_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	    "@P1 bra LEAF_DONE;\n"
	    "bra LEAF_WAIT;\n"
	    "LEAF_DONE:\n"
	    "}\n" ::"r"(leaf_smem_addr(bar)), "r"(parity)
	    : "memory");
}
__device__ __forceinline__ void leaf_bulk_copy(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(leaf_smem_addr(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(leaf_smem_addr(bar))
	             : "memory");
}
__device__ __forceinline__ void leaf_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <bool SOFT>
__device__ __forceinline__ void p2p_interact(const float4& s, float tx, float ty, float tz, float eps2, float& ax, float& ay, float& az) {
	const float dx = s.x - tx, dy = s.y - ty, dz = s.z - tz;
	const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, eps2)));
	float inv;
	if (SOFT) asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2));  // r2 >= eps^2: a bare MUFU.RSQ, no denormal fix-up
	else inv = r2 > 0.0f ? rsqrtf(r2) : 0.0f;  // eps = 0: coincident points (and i == j) exert no force
	const float inv2 = inv * inv;
	const float w = (s.w * inv) * inv2;
	ax = fmaf(w, dx, ax);
	ay = fmaf(w, dy, ay);
	az = fmaf(w, dz, az);
}

struct LeafArgs {
	Ctrl* c;
	const float4* posq;      // sorted positions of this step (sources and targets)
	const float4* velm_in;   // sorted velocities
	float4* posq_out;        // new state
	float4* velm_out;
	float4* acc;
	const float4* geom;
	const uint2* info;
	const uint32_t* nbegin;
	const uint32_t* p2p_head;
	const Segment* seg;
	const uint2* p2p;        // {first particle, count} per source leaf
	const float* L;
	float eps2, G, dt;
	int integrator, no_integrate;
	int rank;                     // this rank's slice of the tree-ordered particle array: [c->part[rank], c->part[rank+1])
	int reverse;                  // hand the leaves out from the deepest level upwards
	const uint32_t* items;        // the non-empty leaves of the slice in node order (k_leaf_items), or nullptr: tickets are chunks of node ids
	unsigned long long* stat_inter;
	unsigned long long* stat_leaves;
};

// One tile against the G targets of the block: lane = one source of each of kLeafRows rows, the targets are applied
// by falling through a switch (entry point = G), so there is ONE copy of the interaction code whatever G is —
// per-G unrolled loops were tried and lost to instruction-cache misses (warps of one SM work on different G).
// The tile is padded with zero-charge sources up to a whole group of rows.
template <bool SOFT>
__device__ __forceinline__ void tile_rows(unsigned G, const float4* __restrict__ buf, uint32_t nrows, unsigned lane, const float4* __restrict__ tgt,
                                          float eps2, float (&ax)[16], float (&ay)[16], float (&az)[16]) {
	__syncwarp();
#pragma unroll 1
	for (uint32_t r = 0; r < nrows; r += kLeafRows) {
		float4 s[kLeafRows];
#pragma unroll
		for (int u = 0; u < kLeafRows; ++u) s[u] = buf[(r + u) * 32u + lane];
		float4 t = tgt[G - 1u];  // the coordinates of the next target are fetched one block ahead of their use
#define NB_LEAF_T(k)                                                                                        \
	case k + 1:                                                                                                \
		if (k < kLeafG) {                                                                                        \
			const float4 tn = tgt[k > 0 ? k - 1 : 0];                                                              \
			_Pragma("unroll") for (int u = 0; u < kLeafRows; ++u) p2p_interact<SOFT>(s[u], t.x, t.y, t.z, eps2, ax[k], ay[k], az[k]); \
			t = tn;                                                                                                \
		}
		switch (G) {
			NB_LEAF_T(15) NB_LEAF_T(14) NB_LEAF_T(13) NB_LEAF_T(12) NB_LEAF_T(11) NB_LEAF_T(10) NB_LEAF_T(9) NB_LEAF_T(8)
			NB_LEAF_T(7) NB_LEAF_T(6) NB_LEAF_T(5) NB_LEAF_T(4) NB_LEAF_T(3) NB_LEAF_T(2) NB_LEAF_T(1) NB_LEAF_T(0)
			default: break;
		}
#undef NB_LEAF_T
	}
	__syncwarp();
}

// Sum v[k] over the 32 lanes for 16 values at once: each halving step exchanges the half of the values the
// partner lane is responsible for, so 16 shuffles do the work of 80. On return v[0] of lanes 2k and 2k+1 holds
// the total of value k.
__device__ __forceinline__ void transpose_reduce16(float (&v)[16], unsigned lane) {
#pragma unroll
	for (int h = 8; h >= 1; h >>= 1) {  // partner = lane ^ (2h): lanes with that bit set keep the upper h values
		const bool up = (lane & (2u * h)) != 0u;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
			v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * h);
		}
	}
	v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// The compacted leaf list for small trees: every non-empty childless node of this rank's slice, in node (level-major, Morton) order up
// to warp granularity (one atomic per 32 nodes), so neighbouring leaves are still worked on at the same time.
__global__ void __launch_bounds__(256) k_leaf_items(Ctrl* c, int rank, const uint2* __restrict__ info, const uint32_t* __restrict__ nbegin,
                                                     uint32_t* __restrict__ items, uint32_t items_cap) {
	if (c->status) return;
	const uint32_t n_nodes = c->n_nodes, own_first = c->part[rank], own_end = c->part[rank + 1];
	const unsigned lane = threadIdx.x & 31u;
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_nodes; base += stride) {
		const uint32_t node = base + lane;
		bool leaf = false;
		if (node < n_nodes) {
			const uint2 nf = info[node];
			const uint32_t b = nbegin[node];
			leaf = nf.x == 0u && nf.y != 0u && b >= own_first && b < own_end;
		}
		const unsigned m = __ballot_sync(0xffffffffu, leaf);
		uint32_t at = 0;
		if (lane == 0 && m) at = atomicAdd(&c->n_leaf_items, (uint32_t) __popc(m));
		at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
		if (leaf && at < items_cap) items[at] = node;
	}
}

// Leaves are handed out kLeafChunk node ids at a time by an atomic ticket, so the grid is exactly the resident set.
// A leaf's source list is a chain of segments of {first particle, count} entries. The sources stream through two
// 256-particle tiles per warp: the flat particle range of up to 32 entries (one per lane) is cut at exactly one
// tile — an entry may straddle two tiles, `skip` remembers how much of the first entry is already consumed — and
// copied with fully used 16-byte cp.async rows (the flat slot -> entry map is a bitmap of the entries' end
// positions, built with warp OR-reductions) while the previous tile is being evaluated; the entries of the batch
// after that are already in registers, so neither the list walk nor the particle fetch sits on the critical path.
template <int P, bool SOFT>
__global__ void __launch_bounds__(kLeafWarps * 32, kLeafMinCtas) k_leaf(const LeafArgs a) {
	using E = Expansion<P>;
	constexpr int STRIDE = coef_stride(P);
	constexpr uint32_t END = 0xffffffffu;
	__shared__ float4 sbuf[kLeafWarps][2][kLeafTile];
	__shared__ float4 stgt[kLeafWarps][16];
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#if NBODY_LEAF_BULK
	__shared__ __align__(8) uint64_t sbar[kLeafWarps][2];  // one mbarrier per warp and tile buffer; a single arrival (lane 0) + the copied bytes
	__shared__ uint2 sent[kLeafWarps][32];                  // the batch of list entries in flight for the tile after next (cp.async)
	unsigned bar_parity = 0u;                               // bit b: the phase parity the next wait on buffer b expects
	if (lane == 0) {
		leaf_mbar_init(&sbar[w][0], 1u);
		leaf_mbar_init(&sbar[w][1], 1u);
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	leaf_fence_proxy_async();
	__syncwarp();
#endif
	if (a.c->status) return;  // a pool overflowed: the host grows it and re-runs the step; leave the state untouched
	const uint32_t n_nodes = a.c->n_nodes, n_items = a.c->n_leaf_items;
	const uint32_t own_first = a.c->part[a.rank], own_end = a.c->part[a.rank + 1];
	const unsigned le_mask = (2u << lane) - 1u;  // bits 0..lane
	unsigned long long inter = 0, leaves = 0;
#pragma unroll 1
	for (;;) {
		uint32_t chunk = 0;
		if (lane == 0) chunk = atomicAdd(&a.c->work_ticket[3], a.items ? 1u : (uint32_t) kLeafChunk);
		chunk = __shfl_sync(0xffffffffu, chunk, 0);
		if (chunk >= (a.items ? n_items : n_nodes)) break;
		// a.reverse (developer switch NBODY_LEAF_REVERSE=1): tickets run from the END of the level-major node array, so that the deepest
		// levels (the dense regions: full leaves, the longest source lists) start first. Measured slower both on one GPU (34.8 against
		// 34.0 ms at 2^24) and on a rank's 1/8 share (5.28 against 5.20 ms): the kernel has no tail worth removing — run alone on one GPU,
		// the eight shares of an 8-rank run take 35.3 ms together against 34.8 ms for the whole (profiles/r02k_summary.md).
		if (a.reverse) chunk = (n_nodes - 1u - chunk) & ~(uint32_t) (kLeafChunk - 1);
		uint2 nf_l = make_uint2(1u, 0u);
		uint32_t b_l = 0;
		unsigned todo;
		if (a.items) {
			// small problems (a rank's share of a multi-GPU run): one ticket = one leaf of the compacted leaf list (k_leaf_items). Chunks of
			// 8 node ids are sibling groups, i.e. 8 neighbouring leaves of similar weight: with only ~9 such tickets per warp on a 1/8
			// share of the benchmark the kernel ended in a long tail (4.98 ms against 4.40 ms, profiles/r02k_summary.md)
			chunk = a.items[chunk];
			if (lane == 0) { nf_l = a.info[chunk]; b_l = a.nbegin[chunk]; }
			todo = 1u;
		} else {
			if (lane < (unsigned) kLeafChunk && chunk + lane < n_nodes) { nf_l = a.info[chunk + lane]; b_l = a.nbegin[chunk + lane]; }
			// childless, non-empty, and inside this rank's slice
			todo = __ballot_sync(0xffffffffu, nf_l.x == 0u && nf_l.y != 0u && b_l >= own_first && b_l < own_end);
		}
#pragma unroll 1
		while (todo) {
			const int kk = __ffs(todo) - 1;
			todo &= todo - 1u;
			const uint32_t node = chunk + kk;
			const uint32_t nt = __shfl_sync(0xffffffffu, nf_l.y, kk), b = __shfl_sync(0xffffffffu, b_l, kk);
			++leaves;
			const float4 g = a.geom[node];
#if NBODY_LEAF_FLAT
			uint32_t seg_off_l = 0u, seg_beg_l = 0xffffffffu, E = 0u, nseg = 0u;
			for (uint32_t sj = a.p2p_head[node]; sj != END && nseg < 32u; ++nseg) {
				const Segment sgm = a.seg[sj];
				if (lane == nseg) { seg_off_l = sgm.off; seg_beg_l = E; }
				E += sgm.cnt;
				sj = sgm.next;
			}
#endif
			const uint32_t nblk = (nt + kLeafG - 1) / kLeafG, gmax = (nt + nblk - 1) / nblk;  // even blocks of <= kLeafG targets
#pragma unroll 1
			for (uint32_t t0 = 0; t0 < nt; t0 += gmax) {
				// (a warp reduction leaves G in a uniform register: the per-target guards of tile_rows become uniform branches without
				//  reconvergence barriers; ptxas cannot know that a value that came through a shuffle is the same in every lane)
				const unsigned G = __reduce_min_sync(0xffffffffu, min(gmax, nt - t0));
				__syncwarp();
				if (lane < G) stgt[w][lane] = a.posq[b + t0 + lane];
				float ax[16], ay[16], az[16];
#pragma unroll
				for (int k = 0; k < 16; ++k) ax[k] = ay[k] = az[k] = 0.f;
				unsigned long long nsrc = 0;
				// ---- cursor over the segment chain: entry e0 of segment sg, of which `skip` particles are consumed ----
				uint32_t si = a.p2p_head[node], e0 = 0, skip = 0;
				Segment sg; sg.off = 0; sg.cnt = 0; sg.next = END;
#if NBODY_LEAF_FLAT
				auto flat_addr = [&](uint32_t gidx) -> uint32_t {
					uint32_t addr = 0u;
					for (uint32_t j = 0; j < nseg; ++j) {
						const uint32_t bj = __shfl_sync(0xffffffffu, seg_beg_l, j), oj = __shfl_sync(0xffffffffu, seg_off_l, j);
						if (gidx >= bj) addr = oj + (gidx - bj);
					}
					return addr;
				};
				auto fetch = [&](uint2& ent) -> bool {
					if (e0 >= E) return false;
					const uint32_t ad = flat_addr(e0 + lane);
					ent = make_uint2(0u, 0u);
					if (e0 + lane < E) ent = a.p2p[ad];
					return true;
				};
#else
				auto fetch = [&](uint2& ent) -> bool {  // the (up to) 32 entries at the cursor, lane l holds entry l; warp-uniform result
					while (e0 >= sg.cnt) {
						if (si == END) return false;
						sg = a.seg[si]; si = sg.next; e0 = 0;
					}
					ent = make_uint2(0u, 0u);
					if (e0 + lane < sg.cnt) ent = a.p2p[sg.off + e0 + lane];
					return true;
				};
#endif
#if NBODY_LEAF_BULK
				// The same for the batch after next, but through shared memory with cp.async: the entries are not needed before the next
				// tile has been evaluated, and a register-destination load this far ahead made the warp wait for it at the very next
				// instruction that touched its scoreboard (4.8 % of the kernel's stall samples, profiles/r02b_ncu_k_leaf_source.csv.gz).
				auto fetch_async = [&]() -> bool {
#if NBODY_LEAF_FLAT
					if (e0 >= E) return false;
					const uint32_t ad = flat_addr(e0 + lane);
					if (e0 + lane < E) {
						const unsigned sa = leaf_smem_addr(&sent[w][lane]);
						asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(a.p2p + ad) : "memory");
					} else sent[w][lane] = make_uint2(0u, 0u);
#else
					while (e0 >= sg.cnt) {
						if (si == END) return false;
						sg = a.seg[si]; si = sg.next; e0 = 0;
					}
					if (e0 + lane < sg.cnt) {
						const unsigned sa = leaf_smem_addr(&sent[w][lane]);
						asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(a.p2p + (sg.off + e0 + lane)) : "memory");
					} else sent[w][lane] = make_uint2(0u, 0u);
#endif
					leaf_cp_async_commit();
					return true;
				};
#endif
				// Cut one tile off the flat particle range of `ent`, issue its fill and advance the cursor.
				auto stage = [&](const uint2& ent, float4* tile, uint32_t& fill) {
#if NBODY_LEAF_FLAT
					const uint32_t nvalid = min(32u, E - e0);
#else
					const uint32_t nvalid = min(32u, sg.cnt - e0);
#endif
					const uint32_t sk = lane == 0u ? skip : 0u;
					const uint32_t v = lane < nvalid ? ent.y - sk : 0u;
					uint32_t inc = v;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
						if (lane >= (unsigned) d) inc += u;
					}
					const bool full = lane < nvalid && inc <= (uint32_t) kLeafTile;  // entries that end inside the tile: a prefix of the lanes
					const uint32_t nfull = __popc(__ballot_sync(0xffffffffu, full));
					const uint32_t avail = __shfl_sync(0xffffffffu, inc, nvalid - 1u);
					const uint32_t used = __shfl_sync(0xffffffffu, inc, (nfull - 1u) & 31u);  // particles of the whole entries (if any)
					fill = min(avail, (uint32_t) kLeafTile);
					if (nfull < nvalid) skip = (nfull ? 0u : skip) + (fill - (nfull ? used : 0u));  // entry nfull straddles the tile end
					else skip = 0u;
					e0 += nfull;
#if NBODY_LEAF_BULK
					// one bulk copy per contiguous run: lane l heads a run if its entry does not start where lane l-1's ends
					const uint32_t prev_end = __shfl_up_sync(0xffffffffu, ent.x + ent.y, 1);
					const bool head = lane < nvalid && (lane == 0u || ent.x != prev_end);
					const unsigned heads = __ballot_sync(0xffffffffu, head);
					const uint32_t start = inc - v;                                  // flat slot of this entry's first unconsumed particle
					const unsigned above = lane < 31u ? heads >> (lane + 1u) : 0u;   // heads after this lane
					const uint32_t next = above ? lane + (uint32_t) __ffs(above) : nvalid;  // lane of the next run's head (or one past the entries)
					const uint32_t run_end = min(__shfl_sync(0xffffffffu, start, next & 31u), avail);  // (lane `next` == nvalid holds v = 0: its start is avail)
					const uint32_t stop = min(next < nvalid ? run_end : avail, (uint32_t) kLeafTile);
					uint64_t* bar = &sbar[w][tile == sbuf[w][0] ? 0 : 1];
					leaf_fence_proxy_async();   // the generic-proxy reads / padding writes of this buffer's previous use precede the async writes
					__syncwarp();
					if (lane == 0u) leaf_mbar_expect_tx(bar, 16u * fill);
					__syncwarp();
					if (head && start < stop) leaf_bulk_copy(tile + start, a.posq + (ent.x + sk), 16u * (stop - start), bar);
#else
					const uint32_t src0 = ent.x + sk - (inc - v);  // particle index of flat slot f inside entry l: src0 + f
					uint32_t pc = 0;
#pragma unroll
					for (int i = 0; i < kLeafTile / 32; ++i) {
						if (32u * i >= fill) break;
						// bit p of the bitmap: some entry ends at flat position p; entry of slot f = number of ends <= f
						const unsigned word = __reduce_or_sync(0xffffffffu, (full && (inc >> 5) == (uint32_t) i) ? 1u << (inc & 31u) : 0u);
						const uint32_t f = 32u * i + lane;
						const uint32_t e = pc + __popc(word & le_mask);
						const uint32_t s0 = __shfl_sync(0xffffffffu, src0, e & 31u);
						if (f < fill) leaf_cp_async16(tile + f, a.posq + (s0 + f));
						pc += __popc(word);
					}
#endif
					// pad to whole row groups with zero-charge sources (only the last tile of a segment is short)
					for (uint32_t f = fill + lane; f < (fill + kLeafPad - 1u) / kLeafPad * kLeafPad; f += 32u) tile[f] = make_float4(0.f, 0.f, 0.f, 0.f);
#if !NBODY_LEAF_BULK
					leaf_cp_async_commit();
#endif
				};
				uint2 ent_cur = make_uint2(0u, 0u), ent_nxt = make_uint2(0u, 0u);
				uint32_t fill_cur = 0;
				int cur = 0;
#if NBODY_LEAF_BULK
				bool pending = false;  // a batch of entries is in flight into sent[w]
#endif
				bool has_cur = fetch(ent_cur);
				if (has_cur) stage(ent_cur, sbuf[w][0], fill_cur);
				bool has_nxt = has_cur && fetch(ent_nxt);
#pragma unroll 1
				while (has_cur) {
					uint32_t fill_nxt = 0;
#if NBODY_LEAF_BULK
					if (has_nxt) {
						if (pending) { leaf_cp_async_wait0(); ent_nxt = sent[w][lane]; pending = false; }  // issued one tile ago
						stage(ent_nxt, sbuf[w][cur ^ 1], fill_nxt);
					}
					const bool has_nn = has_nxt && fetch_async();  // entries of the batch after next: in flight during the math
					pending = has_nn;
#else
					if (has_nxt) stage(ent_nxt, sbuf[w][cur ^ 1], fill_nxt);
					else leaf_cp_async_commit();  // empty group keeps the wait_group arithmetic uniform
					uint2 ent_nn = make_uint2(0u, 0u);
					const bool has_nn = has_nxt && fetch(ent_nn);  // entries of the batch after next: in flight during the math
#endif
#if NBODY_LEAF_BULK
					leaf_mbar_wait(&sbar[w][cur], (bar_parity >> cur) & 1u);
					bar_parity ^= 1u << cur;
#else
					leaf_cp_async_wait1();
#endif
					tile_rows<SOFT>(G, sbuf[w][cur], (fill_cur + 31u) >> 5, lane, stgt[w], a.eps2, ax, ay, az);
					nsrc += fill_cur;
#if !NBODY_LEAF_BULK
					ent_nxt = ent_nn;
#endif
					fill_cur = fill_nxt; has_cur = has_nxt; has_nxt = has_nn;
					cur ^= 1;
				}
#if !NBODY_LEAF_BULK
				leaf_cp_async_wait0();
#endif
				transpose_reduce16(ax, lane);
				transpose_reduce16(ay, lane);
				transpose_reduce16(az, lane);
				if (lane == 0) inter += nsrc * G;
				const unsigned t = lane >> 1;
				if ((lane & 1u) == 0u && t < G) {
					// far field: L2P of this leaf's local expansion, then the integrator
					float l[E::NC];
					const float4* L4 = reinterpret_cast<const float4*>(a.L + (size_t) node * STRIDE);
#pragma unroll
					for (int q = 0; q < (E::NC + 3) / 4; ++q) {
						const float4 v = L4[q];
						l[4 * q] = v.x;
						if (4 * q + 1 < E::NC) l[4 * q + 1] = v.y;
						if (4 * q + 2 < E::NC) l[4 * q + 2] = v.z;
						if (4 * q + 3 < E::NC) l[4 * q + 3] = v.w;
					}
					const uint32_t i = b + t0 + t;
					const float4 tt = stgt[w][t];
					float fx, fy, fz;
					E::l2p(l, tt.x - g.x, tt.y - g.y, tt.z - g.z, fx, fy, fz);
					const float4 vm = a.velm_in[i];
					const float sc = a.G * tt.w / vm.w;  // a = G q/m * field (src/force.cl:4-10, src/open_cl_simulation.cpp:602-604)
					const float axx = sc * (ax[0] + fx), ayy = sc * (ay[0] + fy), azz = sc * (az[0] + fz);
					a.acc[i] = make_float4(axx, ayy, azz, 0.f);
					if (!a.no_integrate) {
						const float vx = fmaf(axx, a.dt, vm.x), vy = fmaf(ayy, a.dt, vm.y), vz = fmaf(azz, a.dt, vm.z);
						const bool kd = a.integrator == NBODY_KICK_DRIFT;
						a.posq_out[i] = make_float4(fmaf(kd ? vx : vm.x, a.dt, tt.x), fmaf(kd ? vy : vm.y, a.dt, tt.y), fmaf(kd ? vz : vm.z, a.dt, tt.z), tt.w);
						a.velm_out[i] = make_float4(vx, vy, vz, vm.w);
					} else {
						a.posq_out[i] = tt;
						a.velm_out[i] = vm;
					}
				}
			}
		}
	}
	if (lane == 0 && leaves) { atomicAdd(a.stat_inter, inter); atomicAdd(a.stat_leaves, leaves); }
}

template <int P>
static void leaf_t(Sim& s, const LeafArgs& a) {
	int ctas = kLeafMinCtas;
	if (s.cfg.softening > 0.0f) {
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_leaf<P, true>, kLeafWarps * 32, 0);
		k_leaf<P, true><<<kNumSM * (ctas > 0 ? ctas : 1), kLeafWarps * 32, 0, s.stream>>>(a);
	} else {
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_leaf<P, false>, kLeafWarps * 32, 0);
		k_leaf<P, false><<<kNumSM * (ctas > 0 ? ctas : 1), kLeafWarps * 32, 0, s.stream>>>(a);
	}
}

void launch_leaf(Sim& s) {
	LeafArgs a{};
	a.c = s.ctrl; a.posq = s.posq[1]; a.velm_in = s.velm[1]; a.posq_out = s.posq[0]; a.velm_out = s.velm[0]; a.acc = s.acc;
	a.geom = s.geom; a.info = s.info; a.nbegin = s.nbegin; a.p2p_head = s.p2p_head; a.seg = s.pools.seg; a.p2p = s.pools.p2p; a.L = s.L;
	a.eps2 = s.cfg.softening * s.cfg.softening; a.G = s.cfg.force_constant; a.dt = s.dt;
	a.integrator = (int) s.cfg.integrator; a.no_integrate = (s.cfg.flags & NBODY_FLAG_NO_INTEGRATE) ? 1 : 0;
	a.rank = s.rank;
	static const char* env = std::getenv("NBODY_LEAF_REVERSE");
	a.reverse = env ? std::atoi(env) != 0 : 0;
	// ticket granularity: leaves on small trees (fewer than 2^19 nodes at the last step: ~27 chunk tickets per warp), node chunks on
	// large ones (where the compacted list costs 2 %: 34.8 against 34.0 ms at 2^24 on one GPU). NBODY_LEAF_ITEMS=0/1 overrides.
	static const char* env_items = std::getenv("NBODY_LEAF_ITEMS");
	const uint32_t last_nodes = s.steps_done ? s.ctrl_host->n_nodes : (uint32_t) (s.n / 12 + 1);
	const bool items = env_items ? std::atoi(env_items) != 0 : last_nodes < (1u << 19);
	a.items = nullptr;
	if (items && !a.reverse) {
		a.items = s.leaf_items;
		k_leaf_items<<<kNumSM * 4, 256, 0, s.stream>>>(s.ctrl, s.rank, s.info, s.nbegin, s.leaf_items, s.max_nodes);
	}
	a.stat_inter = &s.ctrl->stat_p2p_inter; a.stat_leaves = &s.ctrl->stat_leaves;
	switch (s.cfg.order) {
		case 2: leaf_t<2>(s, a); break;
		case 3: leaf_t<3>(s, a); break;
		case 5: leaf_t<5>(s, a); break;
		default: leaf_t<4>(s, a); break;
	}
}

// ---------------------------------------------------------------------------
// All-pairs tiled P2P (no tree): validation against direct summation on large N
// and the P2P FP32 microbenchmark. Each thread owns 2 targets; sources stream
// through a double-buffered shared-memory tile.
// ---------------------------------------------------------------------------
constexpr int kDirThreads = 256;
constexpr int kDirTile = 1024;
constexpr int kDirTargets = 4;  // targets per thread: one LDS.128 feeds 4 interactions

template <bool SOFT>
__global__ void __launch_bounds__(kDirThreads) k_direct(const float4* __restrict__ src, uint64_t n_src, const float4* __restrict__ tgt,
                                                        uint64_t n_tgt, float eps2, float4* __restrict__ out) {
	__shared__ float4 tile[kDirTile];
	const uint64_t i0 = (uint64_t) blockIdx.x * (kDirTargets * kDirThreads) + threadIdx.x;
	float tx[kDirTargets], ty[kDirTargets], tz[kDirTargets], ax[kDirTargets], ay[kDirTargets], az[kDirTargets];
	// running totals with Kahan compensation across tiles: a plain FP32 sum over 10^7 sources loses ~1e-2
	float sx[kDirTargets], sy[kDirTargets], sz[kDirTargets], cx[kDirTargets], cy[kDirTargets], cz[kDirTargets];
#pragma unroll
	for (int t = 0; t < kDirTargets; ++t) {
		const uint64_t i = i0 + (uint64_t) t * kDirThreads;
		const float4 p = i < n_tgt ? tgt[i] : make_float4(0.f, 0.f, 0.f, 0.f);
		tx[t] = p.x; ty[t] = p.y; tz[t] = p.z; ax[t] = ay[t] = az[t] = 0.f;
		sx[t] = sy[t] = sz[t] = cx[t] = cy[t] = cz[t] = 0.f;
	}
	for (uint64_t base = 0; base < n_src; base += kDirTile) {
		__syncthreads();
		for (int q = threadIdx.x; q < kDirTile; q += kDirThreads) {
			const uint64_t j = base + q;
			tile[q] = j < n_src ? src[j] : make_float4(0.f, 0.f, 0.f, 0.f);  // q = 0 padding exerts no force
		}
		__syncthreads();
#pragma unroll 4
		for (int q = 0; q < kDirTile; ++q) {
			const float4 s = tile[q];
#pragma unroll
			for (int t = 0; t < kDirTargets; ++t) p2p_interact<SOFT>(s, tx[t], ty[t], tz[t], eps2, ax[t], ay[t], az[t]);
		}
#pragma unroll
		for (int t = 0; t < kDirTargets; ++t) {  // fold the tile's partial sum into the compensated total
			float y, u;
			y = __fsub_rn(ax[t], cx[t]); u = __fadd_rn(sx[t], y); cx[t] = __fsub_rn(__fsub_rn(u, sx[t]), y); sx[t] = u; ax[t] = 0.f;
			y = __fsub_rn(ay[t], cy[t]); u = __fadd_rn(sy[t], y); cy[t] = __fsub_rn(__fsub_rn(u, sy[t]), y); sy[t] = u; ay[t] = 0.f;
			y = __fsub_rn(az[t], cz[t]); u = __fadd_rn(sz[t], y); cz[t] = __fsub_rn(__fsub_rn(u, sz[t]), y); sz[t] = u; az[t] = 0.f;
		}
	}
#pragma unroll
	for (int t = 0; t < kDirTargets; ++t) {
		const uint64_t i = i0 + (uint64_t) t * kDirThreads;
		if (i < n_tgt) out[i] = make_float4(sx[t], sy[t], sz[t], 0.f);
	}
}

int direct_field_device(const float4* src, uint64_t n_src, const float4* tgt, uint64_t n_tgt, float eps2, float4* out, cudaStream_t st) {
	const unsigned grid = (unsigned) ((n_tgt + kDirTargets * kDirThreads - 1) / (kDirTargets * kDirThreads));
	if (grid == 0) return NBODY_OK;
	if (eps2 > 0.0f) k_direct<true><<<grid, kDirThreads, 0, st>>>(src, n_src, tgt, n_tgt, eps2, out);
	else k_direct<false><<<grid, kDirThreads, 0, st>>>(src, n_src, tgt, n_tgt, eps2, out);
	return NBODY_OK;
}

__global__ void k_direct_finish(const Ctrl* __restrict__ c, uint64_t n, const float4* __restrict__ field, const float4* __restrict__ posq, const float4* __restrict__ velm,
                                float4* __restrict__ posq_out, float4* __restrict__ velm_out, float4* __restrict__ acc, float G, float dt,
                                int integrator, int no_integrate) {
	if (c->status) return;  // a pool overflowed (the tree build's node pool): the host grows it and re-runs the step from the untouched state
	for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
		const float4 f = field[i], p = posq[i], vm = velm[i];
		const float sc = G * p.w / vm.w;
		const float ax = sc * f.x, ay = sc * f.y, az = sc * f.z;
		acc[i] = make_float4(ax, ay, az, 0.f);
		if (no_integrate) { posq_out[i] = p; velm_out[i] = vm; continue; }
		const float vx = fmaf(ax, dt, vm.x), vy = fmaf(ay, dt, vm.y), vz = fmaf(az, dt, vm.z);
		const bool kd = integrator == NBODY_KICK_DRIFT;
		posq_out[i] = make_float4(fmaf(kd ? vx : vm.x, dt, p.x), fmaf(kd ? vy : vm.y, dt, p.y), fmaf(kd ? vz : vm.z, dt, p.z), p.w);
		velm_out[i] = make_float4(vx, vy, vz, vm.w);
	}
}

// NBODY_FLAG_DIRECT: the whole step by direct summation (state is still Morton-sorted first,
// so particles() keeps the same order contract). Field staged in `acc`, then finished in place.
void launch_direct(Sim& s) {
	const float eps2 = s.cfg.softening * s.cfg.softening;
	float4* field = s.acc;  // staged in place: k_direct_finish reads field[i] before it writes acc[i]
	direct_field_device(s.posq[1], s.n, s.posq[1], s.n, eps2, field, s.stream);
	const uint64_t want = (s.n + 255) / 256;
	k_direct_finish<<<(unsigned) (want > kNumSM * 16 ? kNumSM * 16 : (want ? want : 1)), 256, 0, s.stream>>>(
	    s.ctrl, s.n, field, s.posq[1], s.velm[1], s.posq[0], s.velm[0], s.acc, s.cfg.force_constant, s.dt, (int) s.cfg.integrator,
	    (s.cfg.flags & NBODY_FLAG_NO_INTEGRATE) ? 1 : 0);
}

// Variable time step (nbody_cuda_config::time_step_eta > 0): max |a|^2 over this rank's slice of `acc`, 16 B per particle
// read once (HBM-bound, ~0.05 ms at 2^24). Launched only when the feature is on, so the fixed-step path is unchanged.
// |a|^2 >= 0, so the float ordering equals the ordering of the bit patterns and one atomicMax per block suffices; a NaN
// (bits above +inf) wins the maximum and is rejected by the host rule.
__global__ void __launch_bounds__(256) k_acc_max(Ctrl* c, int rank, const float4* __restrict__ acc) {
	if (c->status) return;
	const uint32_t first = c->part[rank], end = c->part[rank + 1];
	float m = 0.0f;
	bool bad = false;
	for (uint64_t i = (uint64_t) first + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (uint64_t) gridDim.x * blockDim.x) {
		const float4 a = acc[i];
		const float a2 = fmaf(a.z, a.z, fmaf(a.y, a.y, a.x * a.x));
		bad |= !(a2 == a2);
		m = fmaxf(m, a2);  // fmaxf drops NaN operands: `bad` carries them
	}
	uint32_t bits = bad ? 0x7fc00000u : __float_as_uint(m);
	bits = __reduce_max_sync(0xffffffffu, bits);
	__shared__ uint32_t sm[8];
	if ((threadIdx.x & 31u) == 0u) sm[threadIdx.x >> 5] = bits;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t b = sm[0];
		for (int k = 1; k < 8; ++k) b = sm[k] > b ? sm[k] : b;
		if (b) atomicMax(&c->acc_max2_bits, b);
	}
}

void launch_acc_max(Sim& s) { k_acc_max<<<kNumSM * 4, 256, 0, s.stream>>>(s.ctrl, s.rank, s.acc); }

}  // namespace nbody
