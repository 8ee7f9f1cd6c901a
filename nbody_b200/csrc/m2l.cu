// Stage 4: far field. M2L over the grouped interaction lists, then the L2L downsweep.
//
// Replaces the reference's order-0 far field (src/field.cl:153-211: the source
// monopole evaluated once at the target node centre and copied into a 32-byte slot
// per (leaf, interaction), reduced by src/force.cl:52-81 after CPU prefix sums at
// src/open_cl_simulation.cpp:419-484) with order-P Cartesian expansions.
//
// M2L kernel: one CTA per work item = 8 sibling targets (one warp each) sharing one
// candidate list. Candidate geometry + multipoles are staged in shared memory once
// per CTA in chunks of 128 (coalesced float4 loads of 16-byte-aligned records; the
// record stride of 36/20/12 floats makes the per-lane LDS.128 reads conflict-free),
// so a multipole is read from L2 once per 8 targets. Each lane owns one candidate at
// a time and keeps its own partial local expansion in registers; one shuffle
// reduction per target at the end, then RED.ADD into L. FP32-FMA bound: per
// (target, source) pair the derivative tensor (~100 flop at P=4) plus 175 FMAs.
#include "common.cuh"

namespace nbody {

constexpr int kM2LChunk = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
	const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Multipole record in shared memory, read with LDS.128 (records are 16-byte aligned and the
// record stride of 36/20/12 floats keeps a quarter warp on distinct banks).
struct SmemCoefs {
	const float4* p;
	__device__ __forceinline__ float operator[](int i) const {
		const float4 v = p[i >> 2];
		return (i & 3) == 0 ? v.x : (i & 3) == 1 ? v.y : (i & 3) == 2 ? v.z : v.w;
	}
};

template <int P>
__device__ __forceinline__ void m2l_one(float (&Lacc)[Expansion<P>::NC], const float4& tg, const float4& sg, const float* sM, float eps2) {
	using E = Expansion<P>;
	float D[E::NC];
	E::derivatives(tg.x - sg.x, tg.y - sg.y, tg.z - sg.z, eps2, D);
	const SmemCoefs M{reinterpret_cast<const float4*>(sM)};
	E::template m2l<1>(Lacc, M, D);
}

// One CTA per work item, items handed out by an atomic ticket. Candidate chunks are double
// buffered with cp.async: while the warps evaluate chunk c, chunk c+1 (geometry + multipole
// records, fetched by the 1-2 threads that own each slot) is in flight; one barrier per chunk.
template <int P, int NT>
__global__ void __launch_bounds__(NT == 8 ? 256 : 128, NT == 8 ? 2 : 4)
k_m2l(Ctrl* __restrict__ c, const Group* __restrict__ items, uint32_t items_cap, const float4* __restrict__ geom,
      const float* __restrict__ M, float* __restrict__ L, const uint32_t* __restrict__ m2l_id, const uint8_t* __restrict__ m2l_mask,
      float eps2) {
	using E = Expansion<P>;
	constexpr int STRIDE = coef_stride(P);
	constexpr int S4 = STRIDE / 4;
	constexpr int NTHREADS = NT == 8 ? 256 : 128;
	constexpr int TPS = NTHREADS / kM2LChunk;      // threads per candidate slot (2 or 1)
	constexpr int PER = (S4 + TPS - 1) / TPS;      // float4 copies per thread
	__shared__ __align__(16) float sM[2][kM2LChunk * STRIDE];
	__shared__ float4 sgeom[2][kM2LChunk];
	__shared__ uint8_t smask[2][kM2LChunk];
	__shared__ uint32_t s_item;
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	const unsigned my_slot = threadIdx.x / TPS, part = threadIdx.x % TPS;
	const uint32_t n_items = min(c->items_count[NT == 8 ? 0 : 1], items_cap);
	const float4* M4 = reinterpret_cast<const float4*>(M);
	for (;;) {
		__syncthreads();  // everyone is done with the previous item (and with s_item)
		if (threadIdx.x == 0) s_item = atomicAdd(&c->work_ticket[NT == 8 ? 0 : 1], 1u);
		__syncthreads();
		const uint32_t it = s_item;
		if (it >= n_items) break;
		const Group G = items[it];
		const uint32_t target = NT == 8 ? G.first + w : G.first;
		const float4 tg = geom[target];
		float Lacc[E::NC];
#pragma unroll
		for (int a = 0; a < E::NC; ++a) Lacc[a] = 0.0f;
		bool any = false;
		const uint32_t nchunks = (G.list_cnt + kM2LChunk - 1) / kM2LChunk;
		auto issue = [&](uint32_t chunk, int buf) {
			const uint32_t e = chunk * kM2LChunk + my_slot;
			if (e < G.list_cnt) {
				const uint32_t id = m2l_id[G.list_off + e];
				float4* dst = reinterpret_cast<float4*>(sM[buf]) + my_slot * S4;
				const float4* src = M4 + (size_t) id * S4;
#pragma unroll
				for (int j = 0; j < PER; ++j)
					if (part * PER + j < S4) cp_async16(dst + part * PER + j, src + part * PER + j);
				if (part == 0) cp_async16(&sgeom[buf][my_slot], geom + id);
				if (part == TPS - 1) smask[buf][my_slot] = m2l_mask[G.list_off + e];
			}
			cp_async_commit();
		};
		issue(0, 0);
		for (uint32_t ch = 0; ch < nchunks; ++ch) {
			const int cur = ch & 1;
			cp_async_wait_all();
			__syncthreads();  // chunk ch has landed for everyone; everyone has finished chunk ch-1
			if (ch + 1 < nchunks) issue(ch + 1, cur ^ 1);
			const uint32_t ns = min((uint32_t) kM2LChunk, G.list_cnt - ch * kM2LChunk);
			if (NT == 8) {
				for (uint32_t s = lane; s < ns; s += 32)
					if (smask[cur][s] >> w & 1u) { m2l_one<P>(Lacc, tg, sgeom[cur][s], sM[cur] + s * STRIDE, eps2); any = true; }
			} else {
				for (uint32_t s = threadIdx.x; s < ns; s += NTHREADS) { m2l_one<P>(Lacc, tg, sgeom[cur][s], sM[cur] + s * STRIDE, eps2); any = true; }
			}
		}
		if (!__any_sync(0xffffffffu, any)) continue;
		// warp reduction, then lane a adds coefficient a (L[0], the potential term, is not carried)
#pragma unroll
		for (int a = 1; a < E::NC; ++a) {
			float v = Lacc[a];
#pragma unroll
			for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
			if (lane == (unsigned) (a & 31)) atomicAdd(L + (size_t) target * STRIDE + a, v);
		}
	}
}

// L2L: every non-empty node adds the shifted local expansion of its parent; levels ascending.
template <int P>
__global__ void __launch_bounds__(128) k_l2l(const Ctrl* __restrict__ c, int l, const float4* __restrict__ geom,
                                             const uint2* __restrict__ info, const uint32_t* __restrict__ nparent, float* __restrict__ L) {
	using E = Expansion<P>;
	constexpr int STRIDE = coef_stride(P);
	const uint32_t lo = c->level_off[l], hi = c->level_off[l + 1];
	for (uint32_t node = lo + blockIdx.x * blockDim.x + threadIdx.x; node < hi; node += gridDim.x * blockDim.x) {
		if (info[node].y == 0u) continue;
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
static void m2l_t(Sim& s) {
	const float eps2 = s.cfg.softening * s.cfg.softening;
	k_m2l<P, 8><<<kNumSM * 16, 256, 0, s.stream>>>(s.ctrl, s.pools.items[0], s.pools.items_cap, s.geom, s.M, s.L, s.pools.m2l_id,
	                                              s.pools.m2l_mask, eps2);
	k_m2l<P, 1><<<kNumSM * 16, 128, 0, s.stream>>>(s.ctrl, s.pools.items[1], s.pools.items_cap, s.geom, s.M, s.L, s.pools.m2l_id,
	                                              s.pools.m2l_mask, eps2);
}
template <int P>
static void l2l_t(Sim& s) {
	for (int l = 1; l <= (int) s.cfg.max_depth; ++l)
		k_l2l<P><<<kNumSM * 4, 128, 0, s.stream>>>(s.ctrl, l, s.geom, s.info, s.nparent, s.L);
}

void launch_m2l(Sim& s) {
	switch (s.cfg.order) {
		case 2: m2l_t<2>(s); break;
		case 3: m2l_t<3>(s); break;
		default: m2l_t<4>(s); break;
	}
}
void launch_l2l(Sim& s) {
	switch (s.cfg.order) {
		case 2: l2l_t<2>(s); break;
		case 3: l2l_t<3>(s); break;
		default: l2l_t<4>(s); break;
	}
}

}  // namespace nbody
