// Stage 4: far field. M2L over the grouped interaction lists, then the L2L downsweep.
// (Kernel templates; instantiated per expansion order in m2l.cu, m2l_p4.cu and m2l_p5.cu so that the orders compile in parallel:
//  order 5 alone takes ptxas over a minute.)
//
// Replaces the reference's order-0 far field (src/field.cl:153-211: the source
// monopole evaluated once at the target node centre and copied into a 32-byte slot
// per (leaf, interaction), reduced by src/force.cl:52-81 after CPU prefix sums at
// src/open_cl_simulation.cpp:419-484) with order-P Cartesian expansions.
//
// M2L kernel: one CTA per work item = 8 sibling targets (one warp each) sharing one
// candidate list. Candidate geometry + multipoles are staged in shared memory once
// per CTA in chunks of 256 (coalesced 16-byte copies; only the multipole orders 0..P-1 a
// field-only M2L reads are staged: 20/12/4 floats per candidate, a stride that keeps the
// per-lane LDS.128 reads of consecutive slots conflict-free),
// so a multipole is read from L2 once per 8 targets. Each lane owns one candidate at
// a time and keeps its own partial local expansion in registers; one shuffle
// reduction per target at the end, then RED.ADD into L. FP32-FMA bound: per
// (target, source) pair the derivative tensor (~100 flop at P=4) plus 175 FMAs.
#pragma once
#include "common.cuh"

namespace nbody {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
	const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Multipole record in shared memory, read with LDS.128 (records are 16-byte aligned and the
// record stride of 36/20/12 floats keeps consecutive slots of a quarter warp on distinct banks).
struct SmemCoefs {
	const float4* p;
	__device__ __forceinline__ float operator[](int i) const {
		const float4 v = p[i >> 2];
		return (i & 3) == 0 ? v.x : (i & 3) == 1 ? v.y : (i & 3) == 2 ? v.z : v.w;
	}
};

// One (target, source) M2L evaluated at order PE <= P.
template <int P, int PE>
__device__ __forceinline__ void m2l_one(float (&Lacc)[Expansion<P>::NC], const float4& tg, const float4& sg, const float* sM, float eps2) {
	float D[Expansion<PE>::NC];
	Expansion<PE>::derivatives(tg.x - sg.x, tg.y - sg.y, tg.z - sg.z, eps2, D);
	const SmemCoefs M{reinterpret_cast<const float4*>(sM)};
	Expansion<P>::template m2l<1, PE>(Lacc, M, D);
}

// Two independent interactions at order PE: both derivative tensors first, then both contractions, so the
// compiler can interleave two dependency chains (the order-3 tensors are small enough to keep two in registers).
template <int P, int PE>
__device__ __forceinline__ void m2l_two(float (&Lacc)[Expansion<P>::NC], const float4& tg, const float4& ga, const float* Ma, const float4& gb,
                                        const float* Mb, float eps2) {
	float Da[Expansion<PE>::NC], Db[Expansion<PE>::NC];
	Expansion<PE>::derivatives(tg.x - ga.x, tg.y - ga.y, tg.z - ga.z, eps2, Da);
	Expansion<PE>::derivatives(tg.x - gb.x, tg.y - gb.y, tg.z - gb.z, eps2, Db);
	const SmemCoefs A{reinterpret_cast<const float4*>(Ma)}, B{reinterpret_cast<const float4*>(Mb)};
	Expansion<P>::template m2l<1, PE>(Lacc, A, Da);
	Expansion<P>::template m2l<1, PE>(Lacc, B, Db);
}

template <int P, int NT>
struct M2LShared {
	static constexpr int CH = NT == 8 ? 256 : 32;   // candidate slots per chunk = threads per CTA
	static constexpr int MS = coef_stride(P - 1);  // a field-only M2L reads multipole orders 0..P-1 only (|n| >= 1, |n|+|m| <= P)
	float sM[2][CH * MS];
	float4 sgeom[2][CH];
	uint32_t sid[3][CH];                     // ids / masks of chunk k live in ring slot k % 3 (published two chunks ahead)
	uint8_t smask[3][CH], smask_lo[3][CH];
	uint8_t list_hi[NT == 8 ? 8 : 1][CH], list_lo[NT == 8 ? 8 : 1][CH];  // per-warp compacted slot numbers of the two order classes
	uint32_t item;
};

// One CTA per work item: 8 warps for 8 sibling targets (one warp each), or a single-warp CTA for a carried
// target (their lists are short, so a multi-warp CTA would spend its time in barriers),
// items handed out by an atomic ticket. Candidate chunks are double buffered with cp.async:
// while the warps evaluate chunk c, chunk c+1 is in flight (fully coalesced 16-byte copies: thread
// t moves piece t, t+CH, ... of the chunk's records), and the ids of chunk c+2 are already in
// registers, so no address dependency sits between the one barrier per chunk and the math.
// Inside a chunk each warp first compacts the slots its target accepts into two dense lists
// (order P and order P-1), then all 32 lanes work through each list.
template <int P, int NT>
__global__ void __launch_bounds__(NT == 8 ? 256 : 32, NT == 8 ? (P <= 4 ? 2 : 1) : (P <= 4 ? 16 : 8))
k_m2l(Ctrl* __restrict__ c, const Group* __restrict__ items, uint32_t items_cap, const float4* __restrict__ geom,
      const float* __restrict__ M, float* __restrict__ L, const uint32_t* __restrict__ m2l_id, const uint8_t* __restrict__ m2l_mask,
      const uint8_t* __restrict__ m2l_mask_lo, float eps2, uint32_t imp_base, const float* __restrict__ Mimp) {
	using E = Expansion<P>;
	using SH = M2LShared<P, NT>;
	constexpr int CH = SH::CH;
	constexpr int STRIDE = coef_stride(P);
	constexpr int S4 = STRIDE / 4;          // 16-byte pieces per multipole record in global memory
	constexpr int MS = SH::MS, MS4 = MS / 4;  // floats / pieces staged per candidate
	constexpr int PL = P > 2 ? P - 1 : P;  // the low evaluation order
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SH& S = *reinterpret_cast<SH*>(smem_raw);
	const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	const uint32_t n_items = min(c->items_count[NT == 8 ? 0 : 1], items_cap);
	const float4* M4 = reinterpret_cast<const float4*>(M);
	// partitioned mode: ids >= imp_base are nodes of other ranks' trees; the multipoles the lists name were fetched from their owners
	// into Mimp, MS floats (orders 0..P-1, all a field-only M2L reads) per imported node (let.cu)
	const float4* Mi4 = reinterpret_cast<const float4*>(Mimp);
	for (;;) {
		__syncthreads();  // everyone is done with the previous item (and with S.item)
		if (tid == 0) S.item = atomicAdd(&c->work_ticket[NT == 8 ? 0 : 1], 1u);
		__syncthreads();
		const uint32_t it = S.item;
		if (it >= n_items) break;
		const Group G = items[it];
		const uint32_t target = NT == 8 ? G.first + w : G.first;
		const float4 tg = geom[target];
		float Lacc[E::NC];
#pragma unroll
		for (int a = 0; a < E::NC; ++a) Lacc[a] = 0.0f;
		bool any = false;
		const uint32_t nchunks = (G.list_cnt + CH - 1) / CH;
		auto fetch = [&](uint32_t chunk, uint32_t& id, uint32_t& mk) {
			const uint32_t e = chunk * CH + tid;
			id = 0; mk = 0;
			if (chunk < nchunks && e < G.list_cnt) { id = m2l_id[G.list_off + e]; mk = m2l_mask[G.list_off + e] | (uint32_t) m2l_mask_lo[G.list_off + e] << 8; }
		};
		auto publish = [&](uint32_t chunk, uint32_t id, uint32_t mk) {  // my slot's id/masks, at least one barrier before the chunk is issued
			const int r = chunk % 3;
			S.sid[r][tid] = id; S.smask[r][tid] = (uint8_t) mk; S.smask_lo[r][tid] = (uint8_t) (mk >> 8);
		};
		auto issue = [&](uint32_t chunk, int buf) {
			const uint32_t ns = min((uint32_t) CH, G.list_cnt - chunk * CH);
			const uint32_t* ids = S.sid[chunk % 3];
			float4* dst = reinterpret_cast<float4*>(S.sM[buf]);
#pragma unroll
			for (int r = 0; r < MS4; ++r) {
				const uint32_t piece = tid + r * CH, slot = piece / MS4, j = piece - slot * MS4;
				if (slot < ns) {
					const uint32_t id = ids[slot];
					cp_async16(dst + piece, id < imp_base ? M4 + (size_t) id * S4 + j : Mi4 + (size_t) (id - imp_base) * MS4 + j);
				}
			}
			if (tid < ns) cp_async16(&S.sgeom[buf][tid], geom + ids[tid]);
			cp_async_commit();
		};
		uint32_t id_n, mk_n;
		fetch(0, id_n, mk_n);
		publish(0, id_n, mk_n);
		fetch(1, id_n, mk_n);
		publish(1, id_n, mk_n);
		fetch(2, id_n, mk_n);
		__syncthreads();
		issue(0, 0);
		for (uint32_t ch = 0; ch < nchunks; ++ch) {
			const int cur = ch & 1;
			cp_async_wait_all();
			__syncthreads();  // chunk ch has landed for everyone; everyone has finished chunk ch-1; ids of ch+1 are visible
			if (ch + 1 < nchunks) issue(ch + 1, cur ^ 1);
			const uint32_t ns = min((uint32_t) CH, G.list_cnt - ch * CH);
			const uint8_t* mk_acc = S.smask[ch % 3];
			const uint8_t* mk_lo = S.smask_lo[ch % 3];
			// ---- compact the accepted slots of my target into the two order classes ----
			uint32_t cnt_h = 0, cnt_l = 0;
			constexpr int PERW = CH / 32;  // slots per lane
			const uint32_t s0 = 0u;
#pragma unroll
			for (int i = 0; i < PERW; ++i) {
				const uint32_t s = s0 + lane + 32 * i;
				bool acc = false, lo = false;
				if (s < ns) {
					acc = NT == 8 ? (mk_acc[s] >> w & 1u) : true;
					lo = (mk_lo[s] >> (NT == 8 ? w : 0u)) & 1u;
				}
				const bool hi = acc && !lo;
				const unsigned bh = __ballot_sync(0xffffffffu, hi), bl = __ballot_sync(0xffffffffu, lo);
				if (hi) S.list_hi[w][cnt_h + __popc(bh & lt_mask)] = (uint8_t) (s - s0);
				if (lo) S.list_lo[w][cnt_l + __popc(bl & lt_mask)] = (uint8_t) (s - s0);
				cnt_h += __popc(bh); cnt_l += __popc(bl);
			}
			__syncwarp();
			any = any || (cnt_h + cnt_l) != 0;
			// The last 32-wide step of the order-P list would leave lanes idle: fill them with pairs taken from the tail of the
			// order-(P-1) list (evaluating a pair at the higher order costs nothing there and only improves it).
			const uint32_t take = min((32u - (cnt_h & 31u)) & 31u, cnt_l);
			cnt_l -= take;
			const uint32_t tot_h = cnt_h + take;
			auto hi_slot = [&](uint32_t k) -> uint32_t { return k < cnt_h ? S.list_hi[w][k] : S.list_lo[w][cnt_l + (k - cnt_h)]; };
			// order P pairs; the slot number and geometry of the next iteration are fetched before the math of this one
			{
				uint32_t k = lane;
				uint32_t sc = 0; float4 gc = make_float4(0.f, 0.f, 0.f, 0.f);
				if (k < tot_h) { sc = s0 + hi_slot(k); gc = S.sgeom[cur][sc]; }
				while (k < tot_h) {
					const uint32_t kn = k + 32;
					uint32_t sn = 0; float4 gn = gc;
					if (kn < tot_h) { sn = s0 + hi_slot(kn); gn = S.sgeom[cur][sn]; }
					m2l_one<P, P>(Lacc, tg, gc, S.sM[cur] + sc * MS, eps2);
					k = kn; sc = sn; gc = gn;
				}
			}
			// order P-1 pairs: two interactions per lane in flight while whole 64-slot strides remain (their
			// derivative chains and FMA streams interleave), then the remainder one at a time
			uint32_t base = 0;
			if (cnt_l >= 64) {
				uint32_t sa = s0 + S.list_lo[w][lane], sb = s0 + S.list_lo[w][32 + lane];
				float4 ga = S.sgeom[cur][sa], gb = S.sgeom[cur][sb];
				for (; base + 64 <= cnt_l; base += 64) {
					uint32_t san = sa, sbn = sb; float4 gan = ga, gbn = gb;
					if (base + 128 <= cnt_l) {
						san = s0 + S.list_lo[w][base + 64 + lane]; sbn = s0 + S.list_lo[w][base + 96 + lane];
						gan = S.sgeom[cur][san]; gbn = S.sgeom[cur][sbn];
					}
					m2l_two<P, PL>(Lacc, tg, ga, S.sM[cur] + sa * MS, gb, S.sM[cur] + sb * MS, eps2);
					sa = san; sb = sbn; ga = gan; gb = gbn;
				}
			}
			for (uint32_t k = base + lane; k < cnt_l; k += 32) {
				const uint32_t s = s0 + S.list_lo[w][k];
				m2l_one<P, PL>(Lacc, tg, S.sgeom[cur][s], S.sM[cur] + s * MS, eps2);
			}
			// ids of chunk ch+2 become visible at the next barrier (ring slot (ch+2)%3 was last read for chunk ch-1,
			// which every warp finished before this iteration's barrier); the registers then prefetch chunk ch+3
			publish(ch + 2, id_n, mk_n);
			fetch(ch + 3, id_n, mk_n);
		}
		if (!any) continue;  // warp-uniform
		// warp reduction (transposing: 31 shuffles for 32 coefficients), then lane a-1 adds coefficient a
		// (L[0], the potential term, is not carried)
		{
			float v[32];
#pragma unroll
			for (int a = 0; a < 32; ++a) v[a] = a + 1 < E::NC ? Lacc[a + 1] : 0.0f;
			transpose_reduce32(v, lane);
			if (lane + 1u < (unsigned) E::NC) atomicAdd(L + (size_t) target * STRIDE + lane + 1u, v[0]);
#pragma unroll
			for (int a = 33; a < E::NC; ++a) {
				float x = Lacc[a];
#pragma unroll
				for (int d = 16; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
				if (lane == (unsigned) (a & 31)) atomicAdd(L + (size_t) target * STRIDE + a, x);
			}
		}
	}
}

// L2L: every non-empty node adds the shifted local expansion of its parent; levels ascending.
template <int P>
__global__ void __launch_bounds__(128) k_l2l(const Ctrl* __restrict__ c, int l, const float4* __restrict__ geom,
                                             const uint2* __restrict__ info, const uint32_t* __restrict__ nparent,
                                             const uint32_t* __restrict__ nbegin, int rank, float* __restrict__ L) {
	using E = Expansion<P>;
	constexpr int STRIDE = coef_stride(P);
	const uint32_t lo = c->level_off[l], hi = c->level_off[l + 1];
	const uint32_t own_first = c->part[rank], own_end = c->part[rank + 1];
	for (uint32_t node = lo + blockIdx.x * blockDim.x + threadIdx.x; node < hi; node += gridDim.x * blockDim.x) {
		const uint32_t cnt = info[node].y;
		if (cnt == 0u) continue;
		const uint32_t nb0 = nbegin[node];
		if (nb0 >= own_end || nb0 + cnt <= own_first) continue;  // holds none of this rank's particles: its local expansion is never read
		const uint32_t par = nparent[node];
		const float4 g = geom[node], gp = geom[par];
		float lp[E::NC], lc[E::NC];
		const float4* Lp4 = reinterpret_cast<const float4*>(L + (size_t) par * STRIDE);
		float4* Lc4 = reinterpret_cast<float4*>(L + (size_t) node * STRIDE);
#pragma unroll
		for (int a = 0; a < (E::NC + 3) / 4; ++a) {
			const float4 v = Lp4[a], u = Lc4[a];
			lp[4 * a] = v.x; lc[4 * a] = u.x;
			if (4 * a + 1 < E::NC) { lp[4 * a + 1] = v.y; lc[4 * a + 1] = u.y; }
			if (4 * a + 2 < E::NC) { lp[4 * a + 2] = v.z; lc[4 * a + 2] = u.z; }
			if (4 * a + 3 < E::NC) { lp[4 * a + 3] = v.w; lc[4 * a + 3] = u.w; }
		}
		E::template l2l<1>(lc, lp, g.x - gp.x, g.y - gp.y, g.z - gp.z);
#pragma unroll
		for (int a = 0; a < (E::NC + 3) / 4; ++a)
			Lc4[a] = make_float4(lc[4 * a], 4 * a + 1 < E::NC ? lc[4 * a + 1] : 0.f, 4 * a + 2 < E::NC ? lc[4 * a + 2] : 0.f,
			                     4 * a + 3 < E::NC ? lc[4 * a + 3] : 0.f);
	}
}

template <int P>
void m2l_t(Sim& s) {
	const float eps2 = s.cfg.softening * s.cfg.softening;
	cudaFuncSetAttribute(k_m2l<P, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(M2LShared<P, 8>));
	cudaFuncSetAttribute(k_m2l<P, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(M2LShared<P, 1>));
	const uint32_t imp_base = s.let ? s.max_nodes : 0xffffffffu;
	k_m2l<P, 8><<<kNumSM * (P <= 4 ? 2 : 1), 256, sizeof(M2LShared<P, 8>), s.stream>>>(s.ctrl, s.pools.items[0], s.pools.items_cap, s.geom, s.M, s.L,
	                                                                  s.pools.m2l_id, s.pools.m2l_mask, s.pools.m2l_mask_lo, eps2, imp_base, s.Mimp);
	k_m2l<P, 1><<<kNumSM * (P <= 4 ? 16 : 8), 32, sizeof(M2LShared<P, 1>), s.stream>>>(s.ctrl, s.pools.items[1], s.pools.items_cap, s.geom, s.M, s.L,
	                                                                  s.pools.m2l_id, s.pools.m2l_mask, s.pools.m2l_mask_lo, eps2, imp_base, s.Mimp);
}
template <int P>
void l2l_t(Sim& s) {
	for (int l = 1; l <= (s.depth_bound < (int) s.cfg.max_depth ? s.depth_bound : (int) s.cfg.max_depth); ++l)
		k_l2l<P><<<kNumSM * 4, 128, 0, s.stream>>>(s.ctrl, l, s.geom, s.info, s.nparent, s.nbegin, s.rank, s.L);
}


}  // namespace nbody
