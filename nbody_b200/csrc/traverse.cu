// Stage 3: dual-tree traversal on the device.
//
// Restates the reference's traversal rule — src/interaction.cl:22-99 (8x8 child
// pairs of a reducible pair, MAC at :64-82) and the host partition loop
// src/open_cl_simulation.cpp:247-266 — without the device->host->device round trip
// per tree level and without the reference's precomputed-interaction slot buffer
// (src/open_cl_simulation.cpp:371-417, src/field.cl:129-145).
//
// Formulation. The reference refines unordered node pairs; the symmetric closure
// of that recursion is organised here by TARGET: in round r every active target A
// owns a "near list" near(A) = { B : pair (A,B) is reached and must be refined }.
// The 8 children of A inherit near(A) as their common candidate source: each entry
// B contributes its 8 children (or B itself when childless). Each (child, candidate)
// is classified with the reference MAC into
//   M2L   (can_approx)                      -> grouped M2L list (candidate, 8-bit target mask)
//   near  (!can_approx, one side splits)    -> near list of that child for round r+1
//   P2P   (!can_approx, both childless)     -> P2P source list of that child
// A childless target whose near list is not empty is carried to the next round as a
// single-target group (the reference splits only the side that has children).
// Every directed (target, source) pair produced is exactly one direction of one
// unordered pair the reference's recursion produces (tests/test_gpu_parity.py).
//
// Kernel shape: one CTA (8 warps) per group, groups handed out by an atomic ticket.
//   A. the near list is expanded into a candidate table in shared memory (ids, cell
//      geometry, flags) with coalesced loads — once per CTA, shared by the 8 targets;
//   B. warp w classifies (target w, 32 candidates) per step; the outcome is kept as three
//      ballots per (target, batch) in shared memory, so nothing is classified twice;
//   C. warp scans over the ballot popcounts give exact list sizes and write positions;
//   D. one atomic per list reserves space; E. the lists are written in candidate order
//      (deterministic content; only their placement in the pools depends on timing).
#include "common.cuh"

namespace nbody {

constexpr int kTravThreads = 256;
constexpr int kTravEntries = 192;                 // near entries expanded per chunk
constexpr int kTravCand = kTravEntries * 8;       // candidate slots per chunk
constexpr int kTravBatches = kTravCand / 32;      // 64

// MAC, src/interaction.cl:64-82, FP32 round-to-nearest without FMA contraction: bit-identical
// to oracle mac_accept(). For the reference's ratio 0.5 (ratio^2 = 0.25) the IEEE division is
// replaced by an exactly equivalent test: fl(ext2/d2) < 0.25  <=>  ext2/d2 < 0.25 - 2^-27
// (the rounding boundary below 0.25, ties go to the even 0.25)  <=>  d2/4 - ext2 > d2 * 2^-27,
// where d2/4 and d2*2^-27 are exact and the subtraction is exact whenever the outcome is in
// doubt (Sterbenz: ext2 in [d2/8, d2/2]); outside that range the sign is unambiguous.
// Returns 0 = not accepted, 1 = accepted (order P), 2 = accepted and well enough separated for order P-1
// (ext2 < tau * d2, an implementation choice that does not touch the lists; mirrored in FP32 by the oracle).
template <bool QUARTER>
__device__ __forceinline__ unsigned mac_classify(float ax, float ay, float az, float ad, const float4& b, float ratio_sq, float tau) {
	const float dx = __fsub_rn(b.x, ax), dy = __fsub_rn(b.y, ay), dz = __fsub_rn(b.z, az);
	const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
	const float ext = __fadd_rn(ad, b.w);
	const float ext2 = __fmul_rn(__fmul_rn(0.75f, ext), ext);
	bool accept;
	if (QUARTER) accept = __fsub_rn(__fmul_rn(0.25f, d2), ext2) > __fmul_rn(d2, 7.450580596923828125e-9f);  // 2^-27
	else accept = __fdiv_rn(ext2, d2) < ratio_sq;
	return accept ? (ext2 < __fmul_rn(tau, d2) ? 2u : 1u) : 0u;
}

// Seeds of the traversal: near(root) = {root}, the reference's seed interaction {0,0}. Partitioned mode: the roots of the other
// ranks' imported trees as well — every rank's tree is the global octree restricted to its own particles, so the own targets
// against each of these source trees together cover the global recursion exactly once (let.cu).
__global__ void k_traverse_init(Ctrl* c, const uint2* __restrict__ info, uint32_t* near0, Group* q1, const TraverseSeeds seeds) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		uint32_t ncand = 0;
		for (uint32_t k = 0; k < seeds.n; ++k) { near0[k] = seeds.id[k]; ncand += info[seeds.id[k]].x ? 8u : 1u; }
		c->near_cursor[0] = seeds.n;
		c->near_cursor[1] = 0;
		const uint2 r = info[0];
		Group g{};
		if (r.x) { g.first = r.x; g.nt = 8; }
		else { g.first = 0; g.nt = 1; }
		g.n_cand = ncand;
		g.list_off = 0; g.list_cnt = seeds.n;
		q1[0] = g;
		c->gq_count[1] = r.y ? 1u : 0u;
		c->gq_count[0] = 0;
	}
}

__global__ void k_round_prep(Ctrl* c, int r) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		c->near_cursor[r & 1] = 0;
		c->gq_count[(r + 1) & 1] = 0;
		c->work_ticket[2] = 0;
	}
}

struct TraverseArgs {
	Ctrl* c;
	const float4* geom;
	const uint2* info;
	const uint32_t* nbegin;
	int rank;                      // this rank's slice of the tree-ordered particle array is [c->part[rank], c->part[rank+1])
	uint2* near_ref;
	uint32_t* p2p_head;
	const uint32_t* near_in;
	uint32_t* near_out;
	uint64_t near_cap;
	uint2* p2p;
	uint64_t p2p_cap;
	uint32_t* m2l_id;
	uint8_t* m2l_mask;
	uint8_t* m2l_mask_lo;
	uint64_t m2l_cap;
	Segment* seg;
	uint32_t seg_cap;
	const Group* q_in;
	Group* q_out;
	uint32_t gq_cap;
	Group* items8;
	Group* items1;
	uint32_t items_cap;
	float ratio_sq;
	float tau;
	int round;
	// partitioned mode: imported nodes have ids >= imp_base; the lists' imported leaves / nodes are marked for the halo and multipole fetch
	uint32_t imp_base;
	uint32_t* imp_hoff;
	uint32_t* imp_mflag;
};

// Sized so that 4 CTAs fit in the 228 KB of an SM: sizeof(TravSmem) + 1 KB reserve <= 57 KB (static_assert below).
struct TravSmem {
	float4 cgeom[kTravCand];
	uint32_t cid[kTravCand];
	uint8_t cflag[kTravCand];               // bit0: non-empty, bit1: has children
	uint32_t bal[4][8][kTravBatches];       // ballots per (list, target, batch): 0 = M2L, 1 = near, 2 = P2P, 3 = M2L at low order
	uint32_t pre[2][8][kTravBatches];       // exclusive prefix of the near / P2P popcounts over the batches ([0] near, [1] P2P)
	uint32_t uni[kTravBatches], upre[kTravBatches];  // union of the M2L ballots over the targets, and its prefix
	uint32_t chunk_cnt[5][8];               // per-chunk totals: 0 = M2L (per target), 1 = near, 2 = P2P, 3 = next-round candidate slots, 4 = low-order M2L
	uint32_t total[5][8];                   // per-group totals
	uint32_t off[3][8];                     // reserved offsets: [0][0] = M2L group list, [1][t] near, [2][t] P2P
	uint32_t running[3][8];
	uint32_t warp_sums[32];
	uint32_t ncand, u_total, m2l_total, item, ok;
};

static_assert(sizeof(TravSmem) + 1024 <= 233472 / 5, "TravSmem must allow 5 CTAs per SM");

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, unsigned lane, uint32_t& total) {
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (unsigned) d) inc += t;
	}
	total = __shfl_sync(0xffffffffu, inc, 31);
	return inc - v;
}

template <bool QUARTER, bool LET>
__global__ void __launch_bounds__(kTravThreads) k_traverse(const TraverseArgs a) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	TravSmem& S = *reinterpret_cast<TravSmem*>(smem_raw);
	Ctrl* c = a.c;
	const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	const uint32_t n_groups = min(c->gq_count[a.round & 1], a.gq_cap);
	const int out = a.round & 1;
	const uint32_t own_first = c->part[a.rank], own_end = c->part[a.rank + 1];
	for (;;) {
		__syncthreads();
		if (tid == 0) S.item = atomicAdd(&c->work_ticket[2], 1u);
		__syncthreads();
		const uint32_t gi = S.item;
		if (gi >= n_groups) break;
		const Group G = a.q_in[gi];
		const uint32_t nt = G.nt;
		// target data: for nt == 8 warp w owns target w; for nt == 1 every warp works for target 0
		const uint32_t my_t = nt == 8 ? w : 0u;
		const float4 tg = a.geom[G.first + my_t];
		const uint2 ti = a.info[G.first + my_t];
		const uint32_t tb = a.nbegin[G.first + my_t];
		// a target is active when it holds particles of this rank's Morton range
		const bool t_act = ti.y > 0 && tb < own_end && tb + ti.y > own_first, t_ch = ti.x != 0;
		if (tid < 40) { S.total[tid >> 3][tid & 7] = 0; }
		const uint32_t nchunks = (G.list_cnt + kTravEntries - 1) / kTravEntries;
		for (int pass = 0; pass < 2; ++pass) {
			if (pass == 1) {
				// ---- D. reserve exact space: thread t for target t's lists, thread 8 for the group's M2L list ----
				__syncthreads();
				if (tid == 0) S.ok = 1;
				__syncthreads();
				if (tid < nt) {
					const uint32_t nn = S.total[1][tid], np = S.total[2][tid];
					S.off[1][tid] = 0; S.off[2][tid] = 0;
					if (nn) {
						const unsigned long long o = atomicAdd(&c->near_cursor[out], (unsigned long long) nn);
						if (o + nn > a.near_cap) { S.ok = 0; atomicOr(&c->status, kOvfNear); } else S.off[1][tid] = (uint32_t) o;
					}
					if (np) {
						const unsigned long long o = atomicAdd(&c->p2p_cursor, (unsigned long long) np);
						if (o + np > a.p2p_cap) { S.ok = 0; atomicOr(&c->status, kOvfP2P); } else S.off[2][tid] = (uint32_t) o;
					}
					S.running[1][tid] = 0; S.running[2][tid] = 0;
				} else if (tid == 8) {
					const uint32_t nm = S.m2l_total;
					S.off[0][0] = 0; S.running[0][0] = 0;
					if (nm) {
						const unsigned long long o = atomicAdd(&c->m2l_cursor, (unsigned long long) nm);
						if (o + nm > a.m2l_cap) { S.ok = 0; atomicOr(&c->status, kOvfM2L); } else S.off[0][0] = (uint32_t) o;
					}
				}
				__syncthreads();
				if (!S.ok) break;
			} else if (tid == 0) S.m2l_total = 0;
			for (uint32_t ch = 0; ch < nchunks; ++ch) {
				const uint32_t e0 = ch * kTravEntries;
				const uint32_t ne = min((uint32_t) kTravEntries, G.list_cnt - e0);
				if (pass == 0 || nchunks > 1) {
					// ---- A. expand near entries into the candidate table ----
					__syncthreads();
					uint32_t B = 0, nslots = 0;
					uint2 bi = make_uint2(0u, 0u);
					if (tid < ne) { B = a.near_in[G.list_off + e0 + tid]; bi = a.info[B]; nslots = bi.x ? 8u : 1u; }
					uint32_t wtot;
					const uint32_t wex = warp_excl_scan(nslots, lane, wtot);
					if (lane == 31) S.warp_sums[w] = wtot;
					__syncthreads();
					uint32_t base = wex;
					for (unsigned k = 0; k < w; ++k) base += S.warp_sums[k];
					if (tid == 0) { uint32_t t = 0; for (int k = 0; k < kTravThreads / 32; ++k) t += S.warp_sums[k]; S.ncand = t; }
					for (uint32_t k = 0; k < nslots; ++k) S.cid[base + k] = bi.x ? bi.x + k : B;
					__syncthreads();
					const uint32_t ncand = S.ncand;
					for (uint32_t s = tid; s < ncand; s += kTravThreads) {
						const uint32_t id = S.cid[s];
						const uint2 ci = a.info[id];
						S.cgeom[s] = a.geom[id];
						S.cflag[s] = (uint8_t) ((ci.y > 0 ? 1u : 0u) | (ci.x ? 2u : 0u));
					}
					if (tid < 40) S.chunk_cnt[tid >> 3][tid & 7] = 0;
					if (tid < 32 && ncand + tid < kTravCand) S.cflag[ncand + tid] = 0;  // the tail of the last batch: flag 0 = not a candidate
					__syncthreads();
					// ---- B. classify: (target, batch) pairs round-robin over the warps; branch-free inside the loop ----
					// (slots past ncand carry flag 0, the MAC is evaluated unconditionally and masked, counters stay in
					//  registers and reach shared memory once per warp)
					const uint32_t nb = (ncand + 31) / 32;
					uint32_t c_lo = 0, c_nc = 0;
					const uint32_t tgt_id = G.first + (nt == 8 ? w : 0u);
					const uint32_t trow = nt == 8 ? w : 0u;
					if (t_act) {
						for (uint32_t q = w; q < nt * nb; q += 8) {
							const uint32_t b = nt == 8 ? q >> 3 : q;
							const uint32_t s = 32 * b + lane;
							const unsigned fl = S.cflag[s];
							const unsigned cls = mac_classify<QUARTER>(tg.x, tg.y, tg.z, tg.w, S.cgeom[s], a.ratio_sq, a.tau);
							const bool valid = (fl & 1u) != 0;
							const bool accept = valid && cls != 0u && S.cid[s] != tgt_id;
							const bool rest = valid && !accept;
							const bool nearb = rest && (t_ch || (fl & 2u));
							const unsigned m_acc = __ballot_sync(0xffffffffu, accept), m_lo = __ballot_sync(0xffffffffu, accept && cls == 2u);
							const unsigned m_near = __ballot_sync(0xffffffffu, nearb), m_p2p = __ballot_sync(0xffffffffu, rest && !nearb);
							const unsigned m_ch = __ballot_sync(0xffffffffu, nearb && (fl & 2u));
							c_lo += __popc(m_lo);
							c_nc += 8u * __popc(m_ch) + __popc(m_near & ~m_ch);
							if (lane == 0) { S.bal[0][trow][b] = m_acc; S.bal[1][trow][b] = m_near; S.bal[2][trow][b] = m_p2p; S.bal[3][trow][b] = m_lo; }
						}
						if (lane == 0 && (c_lo | c_nc)) { atomicAdd(&S.chunk_cnt[4][trow], c_lo); atomicAdd(&S.chunk_cnt[3][trow], c_nc); }
					} else {
						for (uint32_t q = w; q < nt * nb; q += 8) {  // an empty / foreign target: all-zero ballot rows
							const uint32_t b = nt == 8 ? q >> 3 : q;
							if (lane < 4) S.bal[lane][trow][b] = 0u;
						}
					}
					__syncthreads();
					// ---- C. prefix sums over the batches: 3 lists x nt targets + the M2L union, one warp each ----
					for (uint32_t job = w; job < 3 * nt + 1; job += 8) {
						const bool is_union = job == 3 * nt;
						const uint32_t li = is_union ? 0u : job / nt, t = is_union ? 0u : job - li * nt;
						uint32_t carry = 0;
						for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
							const uint32_t b = b0 + lane;
							uint32_t m = 0;
							if (b < nb) {
								if (is_union) { for (uint32_t tt = 0; tt < nt; ++tt) m |= S.bal[0][tt][b]; S.uni[b] = m; }
								else m = S.bal[li][t][b];
							}
							uint32_t tot;
							const uint32_t ex = warp_excl_scan(__popc(m), lane, tot);
							if (b < nb) { if (is_union) S.upre[b] = carry + ex; else if (li) S.pre[li - 1][t][b] = carry + ex; }
							carry += tot;
						}
						if (lane == 0) { if (is_union) S.u_total = carry; else S.chunk_cnt[li][t] = carry; }
					}
					__syncthreads();
					if (pass == 0) {
						if (tid < 40) S.total[tid >> 3][tid & 7] += S.chunk_cnt[tid >> 3][tid & 7];
						if (tid == 32) S.m2l_total += S.u_total;
					}
				}
				if (pass == 1) {
					// ---- E. write the lists of this chunk in candidate order ----
					const uint32_t ncand = S.ncand, nb = (ncand + 31) / 32;
					for (uint32_t q = w; q < nt * nb; q += 8) {
						const uint32_t t = nt == 8 ? w : 0u, b = nt == 8 ? q >> 3 : q;
						const uint32_t s = 32 * b + lane;
						const uint32_t mn = S.bal[1][t][b], mp = S.bal[2][t][b];
						if (mn >> lane & 1u) a.near_out[S.off[1][t] + S.running[1][t] + S.pre[0][t][b] + __popc(mn & lt_mask)] = S.cid[s];
						if (mp >> lane & 1u) {
							const uint32_t src = S.cid[s];
							const uint2 en = make_uint2(a.nbegin[src], a.info[src].y);
							a.p2p[S.off[2][t] + S.running[2][t] + S.pre[1][t][b] + __popc(mp & lt_mask)] = en;
							if (LET && (en.x & kImported)) a.imp_hoff[en.x & ~kImported] = en.y;  // (only imported leaves carry the tag; every writer stores the same count)
						}
					}
					for (uint32_t b = w; b < nb; b += 8) {
						const uint32_t u = S.uni[b];
						if (u >> lane & 1u) {
							unsigned am = 0, al = 0;
							for (uint32_t tt = 0; tt < nt; ++tt) { am |= (S.bal[0][tt][b] >> lane & 1u) << tt; al |= (S.bal[3][tt][b] >> lane & 1u) << tt; }
							const uint32_t pos = S.off[0][0] + S.running[0][0] + S.upre[b] + __popc(u & lt_mask);
							const uint32_t mid = S.cid[32 * b + lane];
							a.m2l_id[pos] = mid;
							if (LET && mid >= a.imp_base) a.imp_mflag[mid - a.imp_base] = 1u;
							a.m2l_mask[pos] = (uint8_t) am;
							a.m2l_mask_lo[pos] = (uint8_t) al;
						}
					}
					__syncthreads();
					if (tid < nt) { S.running[1][tid] += S.chunk_cnt[1][tid]; S.running[2][tid] += S.chunk_cnt[2][tid]; }
					if (tid == 8) S.running[0][0] += S.u_total;
				}
			}
		}
		__syncthreads();
		if (!S.ok) continue;
		// ---- F. publish: near lists -> next round's groups, P2P segments, the M2L work item, statistics ----
		if (tid < nt) {
			const uint32_t target = G.first + tid;
			const uint2 tin = a.info[target];
			const uint32_t nn = S.total[1][tid], np = S.total[2][tid];
			if (tin.y && (nn || np)) {
				a.near_ref[target] = make_uint2(S.off[1][tid], nn);
				if (nn) {
					const uint32_t qi = atomicAdd(&c->gq_count[(a.round + 1) & 1], 1u);
					if (qi < a.gq_cap) {
						Group g{};
						if (tin.x) { g.first = tin.x; g.nt = 8; } else { g.first = target; g.nt = 1; }
						g.list_off = S.off[1][tid]; g.list_cnt = nn; g.n_cand = S.total[3][tid];
						a.q_out[qi] = g;
					} else atomicOr(&c->status, kOvfGroups);
				}
				if (np) {
					const uint32_t si = atomicAdd(&c->seg_cursor, 1u);
					if (si < a.seg_cap) {
						Segment sg; sg.off = S.off[2][tid]; sg.cnt = np; sg.next = a.p2p_head[target];
						a.seg[si] = sg;
						a.p2p_head[target] = si;
					} else atomicOr(&c->status, kOvfSeg);
				}
			}
		} else if (tid == 32) {
			uint32_t inter = 0, p2pn = 0, nearn = 0, low = 0;
			for (uint32_t t = 0; t < nt; ++t) { inter += S.total[0][t]; p2pn += S.total[2][t]; nearn += S.total[1][t]; low += S.total[4][t]; }
			if (S.m2l_total) {
				const int which = nt == 8 ? 0 : 1;
				const uint32_t ii = atomicAdd(&c->items_count[which], 1u);
				if (ii < a.items_cap) {
					Group it{};
					it.first = G.first; it.nt = nt; it.list_off = S.off[0][0]; it.list_cnt = S.m2l_total;
					(which == 0 ? a.items8 : a.items1)[ii] = it;
				} else atomicOr(&c->status, kOvfItems);
			}
			atomicAdd(&c->stat_m2l_inter, (unsigned long long) inter);
			if (low) atomicAdd(&c->stat_m2l_low, (unsigned long long) low);
			atomicAdd(&c->stat_p2p_entries, (unsigned long long) p2pn);
			atomicAdd(&c->stat_near, (unsigned long long) nearn);
		}
	}
}

void launch_traversal(Sim& s) {
	Pools& p = s.pools;
	const bool quarter = s.cfg.mac_ratio * s.cfg.mac_ratio == 0.25f;
	const bool let = s.let != nullptr && s.imp_hoff != nullptr && s.imp_mflag != nullptr;
	auto kernel = quarter ? (let ? k_traverse<true, true> : k_traverse<true, false>) : (let ? k_traverse<false, true> : k_traverse<false, false>);
	cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(TravSmem));
	k_traverse_init<<<1, 32, 0, s.stream>>>(s.ctrl, s.info, p.near[0], p.gq[1], s.seeds);
	TraverseArgs a{};
	a.c = s.ctrl; a.geom = s.geom; a.info = s.info; a.near_ref = s.near_ref; a.p2p_head = s.p2p_head;
	a.nbegin = s.nbegin; a.rank = s.rank;
	a.near_cap = p.near_cap; a.p2p = p.p2p; a.p2p_cap = p.p2p_cap; a.m2l_id = p.m2l_id; a.m2l_mask = p.m2l_mask; a.m2l_mask_lo = p.m2l_mask_lo; a.m2l_cap = p.m2l_cap;
	a.seg = p.seg; a.seg_cap = p.seg_cap; a.gq_cap = p.gq_cap; a.items8 = p.items[0]; a.items1 = p.items[1]; a.items_cap = p.items_cap;
	a.ratio_sq = s.cfg.mac_ratio * s.cfg.mac_ratio;
	a.tau = s.cfg.order >= 3 ? s.cfg.low_order_tau : 0.0f;  // order P-1 >= 2 only
	a.imp_base = s.let ? s.max_nodes : 0xffffffffu; a.imp_hoff = s.imp_hoff; a.imp_mflag = s.imp_mflag;
	const int rounds = s.trav_bound < (int) s.cfg.max_depth ? s.trav_bound : (int) s.cfg.max_depth;
	for (int r = 1; r <= rounds; ++r) {
		k_round_prep<<<1, 32, 0, s.stream>>>(s.ctrl, r);
		a.round = r;
		a.near_in = p.near[(r - 1) & 1];
		a.near_out = p.near[r & 1];
		a.q_in = p.gq[r & 1];
		a.q_out = p.gq[(r + 1) & 1];
		kernel<<<kNumSM * 5, kTravThreads, sizeof(TravSmem), s.stream>>>(a);
	}
}

}  // namespace nbody
