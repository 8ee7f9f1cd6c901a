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
// unordered pair the reference's recursion produces (tests/test_gpu_lists.py).
//
// One warp per group; lanes = 4 near entries x 8 children. Two passes over the
// candidates: count, allocate exact space with one atomic per list, then write in
// deterministic (list) order.
#include "common.cuh"

namespace nbody {

__device__ __forceinline__ bool mac_accept(float ax, float ay, float az, float ad, const float4& b, float ratio_sq) {
	// FP32, round-to-nearest, no FMA contraction: bit-identical to oracle mac_accept()
	const float dx = __fsub_rn(b.x, ax), dy = __fsub_rn(b.y, ay), dz = __fsub_rn(b.z, az);
	const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
	const float ext = __fadd_rn(ad, b.w);
	const float ext2 = __fmul_rn(__fmul_rn(0.75f, ext), ext);
	return __fdiv_rn(ext2, d2) < ratio_sq;
}

__global__ void k_traverse_init(Ctrl* c, const uint2* __restrict__ info, uint32_t* near0, Group* q1) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		near0[0] = 0;  // near(root) = {root}: the reference's seed interaction {0,0}
		c->near_cursor[0] = 1;
		c->near_cursor[1] = 0;
		const uint2 r = info[0];
		Group g{};
		if (r.x) { g.first = r.x; g.nt = 8; g.n_cand = 8; }
		else { g.first = 0; g.nt = 1; g.n_cand = 1; }
		g.list_off = 0; g.list_cnt = 1;
		q1[0] = g;
		c->gq_count[1] = r.y ? 1u : 0u;
		c->gq_count[0] = 0;
	}
}

__global__ void k_round_prep(Ctrl* c, int r) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		c->near_cursor[r & 1] = 0;
		c->gq_count[(r + 1) & 1] = 0;
	}
}

struct TraverseArgs {
	Ctrl* c;
	const float4* geom;
	const uint2* info;
	uint2* near_ref;
	uint32_t* p2p_head;
	const uint32_t* near_in;
	uint32_t* near_out;
	uint64_t near_cap;
	uint32_t* p2p;
	uint64_t p2p_cap;
	uint32_t* m2l_id;
	uint8_t* m2l_mask;
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
	int round;
};

__global__ void __launch_bounds__(128) k_traverse(const TraverseArgs a) {
	Ctrl* c = a.c;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	const uint32_t n_groups = min(c->gq_count[a.round & 1], a.gq_cap);
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	const int out = a.round & 1;
	for (uint32_t gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gi < n_groups; gi += warps) {
		const Group G = a.q_in[gi];
		// ---- targets: up to 8 siblings, data replicated in every lane's registers ----
		float4 tg0 = make_float4(0.f, 0.f, 0.f, 0.f);
		uint2 ti0 = make_uint2(0u, 0u);
		if (lane < G.nt) { tg0 = a.geom[G.first + lane]; ti0 = a.info[G.first + lane]; }
		float tx[8], ty[8], tz[8], td[8];
		unsigned act = 0, tch = 0;  // bit t: target t is non-empty / has children
#pragma unroll
		for (int t = 0; t < 8; ++t) {
			tx[t] = __shfl_sync(0xffffffffu, tg0.x, t);
			ty[t] = __shfl_sync(0xffffffffu, tg0.y, t);
			tz[t] = __shfl_sync(0xffffffffu, tg0.z, t);
			td[t] = __shfl_sync(0xffffffffu, tg0.w, t);
			const uint32_t cx = __shfl_sync(0xffffffffu, ti0.x, t), cy = __shfl_sync(0xffffffffu, ti0.y, t);
			if (cy) act |= 1u << t;
			if (cx) tch |= 1u << t;
		}
		uint32_t cnt_near[8], cnt_p2p[8], cnt_nc[8];
#pragma unroll
		for (int t = 0; t < 8; ++t) cnt_near[t] = cnt_p2p[t] = cnt_nc[t] = 0;
		uint32_t cnt_m2l = 0;
		uint32_t off_near = 0, off_p2p = 0, off_m2l = 0;  // lane t holds the offsets of target t; m2l in every lane
		bool ok = true;
		unsigned long long m2l_inter = 0;

#pragma unroll 1
		for (int pass = 0; pass < 2; ++pass) {
			uint32_t run_near[8], run_p2p[8];
#pragma unroll
			for (int t = 0; t < 8; ++t) run_near[t] = run_p2p[t] = 0;
			uint32_t run_m2l = 0;
			if (pass == 1 && !ok) break;
#pragma unroll 1
			for (uint32_t base = 0; base < G.list_cnt; base += 4) {
				const uint32_t e = base + (lane >> 3);
				const unsigned k = lane & 7u;
				bool valid = e < G.list_cnt;
				uint32_t B = 0;
				uint2 bi = make_uint2(0u, 0u);
				if (valid) { B = a.near_in[G.list_off + e]; bi = a.info[B]; }
				uint32_t cand = B;
				uint2 ci = bi;
				if (bi.x) { cand = bi.x + k; ci = a.info[cand]; }
				else valid = valid && k == 0;
				valid = valid && ci.y > 0;
				float4 cg = make_float4(0.f, 0.f, 0.f, 1.f);
				if (valid) cg = a.geom[cand];
				const bool cch = ci.x != 0;
				unsigned amask = 0;
#pragma unroll
				for (int t = 0; t < 8; ++t) {
					if (!(act >> t & 1u)) continue;  // warp-uniform
					const bool same = cand == G.first + t;
					const bool accept = valid && !same && mac_accept(tx[t], ty[t], tz[t], td[t], cg, a.ratio_sq);
					const bool nearb = valid && !accept && ((tch >> t & 1u) || cch);
					const bool p2pb = valid && !accept && !nearb;
					if (accept) amask |= 1u << t;
					const unsigned mn = __ballot_sync(0xffffffffu, nearb), mp = __ballot_sync(0xffffffffu, p2pb);
					if (pass == 0) {
						cnt_near[t] += __popc(mn);
						cnt_p2p[t] += __popc(mp);
						const unsigned mc = __ballot_sync(0xffffffffu, nearb && cch);
						cnt_nc[t] += 8u * __popc(mc) + __popc(mn & ~mc);
					} else {
						const uint32_t on = __shfl_sync(0xffffffffu, off_near, t), op = __shfl_sync(0xffffffffu, off_p2p, t);
						if (nearb) a.near_out[on + run_near[t] + __popc(mn & lt_mask)] = cand;
						if (p2pb) a.p2p[op + run_p2p[t] + __popc(mp & lt_mask)] = cand;
						run_near[t] += __popc(mn);
						run_p2p[t] += __popc(mp);
					}
				}
				const unsigned mm = __ballot_sync(0xffffffffu, amask != 0);
				if (pass == 0) cnt_m2l += __popc(mm);
				else {
					if (amask) {
						const uint32_t pos = off_m2l + run_m2l + __popc(mm & lt_mask);
						a.m2l_id[pos] = cand;
						a.m2l_mask[pos] = (uint8_t) amask;
					}
					run_m2l += __popc(mm);
					m2l_inter += __popc(amask);
				}
			}
			if (pass == 0) {
				// ---- exact allocation: lane t allocates for target t, lane 0 for the group's M2L list ----
				uint32_t my_near = 0, my_p2p = 0;
#pragma unroll
				for (int t = 0; t < 8; ++t) if (lane == (unsigned) t) { my_near = cnt_near[t]; my_p2p = cnt_p2p[t]; }
				bool fail = false;
				if (lane < 8 && my_near) {
					const unsigned long long o = atomicAdd(&c->near_cursor[out], (unsigned long long) my_near);
					if (o + my_near > a.near_cap) { fail = true; atomicOr(&c->status, kOvfNear); } else off_near = (uint32_t) o;
				}
				if (lane < 8 && my_p2p) {
					const unsigned long long o = atomicAdd(&c->p2p_cursor, (unsigned long long) my_p2p);
					if (o + my_p2p > a.p2p_cap) { fail = true; atomicOr(&c->status, kOvfP2P); } else off_p2p = (uint32_t) o;
				}
				if (lane == 0 && cnt_m2l) {
					const unsigned long long o = atomicAdd(&c->m2l_cursor, (unsigned long long) cnt_m2l);
					if (o + cnt_m2l > a.m2l_cap) { fail = true; atomicOr(&c->status, kOvfM2L); } else off_m2l = (uint32_t) o;
				}
				off_m2l = __shfl_sync(0xffffffffu, off_m2l, 0);
				ok = !__any_sync(0xffffffffu, fail);
			}
		}
		if (!ok) continue;
		// ---- publish: near lists -> next round's groups, P2P segments, M2L work item ----
		uint32_t my_near = 0, my_p2p = 0, my_nc = 0;
#pragma unroll
		for (int t = 0; t < 8; ++t) if (lane == (unsigned) t) { my_near = cnt_near[t]; my_p2p = cnt_p2p[t]; my_nc = cnt_nc[t]; }
		if (lane < G.nt && (act >> lane & 1u)) {
			const uint32_t target = G.first + lane;
			a.near_ref[target] = make_uint2(off_near, my_near);
			if (my_near) {
				const uint32_t qi = atomicAdd(&c->gq_count[(a.round + 1) & 1], 1u);
				if (qi < a.gq_cap) {
					Group g{};
					if (ti0.x) { g.first = ti0.x; g.nt = 8; } else { g.first = target; g.nt = 1; }
					g.list_off = off_near; g.list_cnt = my_near; g.n_cand = my_nc;
					a.q_out[qi] = g;
				} else atomicOr(&c->status, kOvfGroups);
			}
			if (my_p2p) {
				const uint32_t si = atomicAdd(&c->seg_cursor, 1u);
				if (si < a.seg_cap) {
					Segment sg; sg.off = off_p2p; sg.cnt = my_p2p; sg.next = a.p2p_head[target];
					a.seg[si] = sg;
					a.p2p_head[target] = si;
				} else atomicOr(&c->status, kOvfSeg);
			}
		}
		unsigned long long p2p_total = my_p2p, near_total = my_near;
#pragma unroll
		for (int d = 4; d >= 1; d >>= 1) {
			p2p_total += __shfl_xor_sync(0xffffffffu, p2p_total, d);
			near_total += __shfl_xor_sync(0xffffffffu, near_total, d);
		}
#pragma unroll
		for (int d = 16; d >= 1; d >>= 1) m2l_inter += __shfl_xor_sync(0xffffffffu, m2l_inter, d);
		if (lane == 0) {
			if (cnt_m2l) {
				const int which = G.nt == 8 ? 0 : 1;
				const uint32_t ii = atomicAdd(&c->items_count[which], 1u);
				if (ii < a.items_cap) {
					Group it{};
					it.first = G.first; it.nt = G.nt; it.list_off = off_m2l; it.list_cnt = cnt_m2l;
					(which == 0 ? a.items8 : a.items1)[ii] = it;
				} else atomicOr(&c->status, kOvfItems);
			}
			atomicAdd(&c->stat_m2l_inter, m2l_inter);
			atomicAdd(&c->stat_p2p_entries, p2p_total);
			atomicAdd(&c->stat_near, near_total);
		}
	}
}

void launch_traversal(Sim& s) {
	Pools& p = s.pools;
	k_traverse_init<<<1, 32, 0, s.stream>>>(s.ctrl, s.info, p.near[0], p.gq[1]);
	TraverseArgs a{};
	a.c = s.ctrl; a.geom = s.geom; a.info = s.info; a.near_ref = s.near_ref; a.p2p_head = s.p2p_head;
	a.near_cap = p.near_cap; a.p2p = p.p2p; a.p2p_cap = p.p2p_cap; a.m2l_id = p.m2l_id; a.m2l_mask = p.m2l_mask; a.m2l_cap = p.m2l_cap;
	a.seg = p.seg; a.seg_cap = p.seg_cap; a.gq_cap = p.gq_cap; a.items8 = p.items[0]; a.items1 = p.items[1]; a.items_cap = p.items_cap;
	a.ratio_sq = s.cfg.mac_ratio * s.cfg.mac_ratio;
	for (int r = 1; r <= (int) s.cfg.max_depth; ++r) {
		k_round_prep<<<1, 32, 0, s.stream>>>(s.ctrl, r);
		a.round = r;
		a.near_in = p.near[(r - 1) & 1];
		a.near_out = p.near[r & 1];
		a.q_in = p.gq[r & 1];
		a.q_out = p.gq[(r + 1) & 1];
		k_traverse<<<kNumSM * 8, 128, 0, s.stream>>>(a);
	}
}

}  // namespace nbody
