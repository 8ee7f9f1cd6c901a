// Stage 2: multipole upsweep. P2M on every childless node with an 8-lane group per
// node (coalesced float4 particle loads, shuffle reduction), then M2M level by level
// from the deepest level to the root with real shift operators.
//
// Replaces src/moment.cl:6-67 (one work-item per node striding 48-byte AoS leaves)
// and src/moment.cl:72-160 + the host compaction loop src/open_cl_simulation.cpp:151-171
// (dead code, D6: no translation of the expansion centre).
#include "common.cuh"

namespace nbody {

template <int P>
__global__ void __launch_bounds__(256) k_p2m(const Ctrl* __restrict__ c, const float4* __restrict__ posq, const float4* __restrict__ geom,
                                             const uint2* __restrict__ info, const uint32_t* __restrict__ nbegin, float* __restrict__ M,
                                             float* __restrict__ L, int stride) {
	using E = Expansion<P>;
	const uint32_t n_nodes = c->n_nodes;
	const unsigned sub = threadIdx.x & 7u;
	const uint32_t groups = (gridDim.x * blockDim.x) >> 3;
	for (uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; node < ((n_nodes + 3u) & ~3u); node += groups) {
		// (the loop bound is rounded up to a multiple of 4 so that whole warps stay converged for the shuffles)
		const bool live = node < n_nodes;
		uint2 nf = make_uint2(1u, 0u);
		if (live) nf = info[node];
		float m[E::NC];
#pragma unroll
		for (int a = 0; a < E::NC; ++a) m[a] = 0.0f;
		const bool leaf = live && nf.x == 0u;
		if (leaf && nf.y) {
			const float4 g = geom[node];
			const uint32_t b = nbegin[node];
			for (uint32_t q = sub; q < nf.y; q += 8u) {
				const float4 p = posq[b + q];
				E::p2m(m, p.x - g.x, p.y - g.y, p.z - g.z, p.w);
			}
		}
#pragma unroll
		for (int a = 0; a < E::NC; ++a) {
			m[a] += __shfl_xor_sync(0xffffffffu, m[a], 1);
			m[a] += __shfl_xor_sync(0xffffffffu, m[a], 2);
			m[a] += __shfl_xor_sync(0xffffffffu, m[a], 4);
		}
		if (live) {
			// zero the local expansion of every node (M2L accumulates with atomics), write M of leaves
			float4* Lr = reinterpret_cast<float4*>(L + (size_t) node * stride);
			for (int a = sub; a < stride / 4; a += 8) Lr[a] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (leaf && sub == 0) {
				float4* Mr = reinterpret_cast<float4*>(M + (size_t) node * stride);
#pragma unroll
				for (int a = 0; a < (E::NC + 3) / 4; ++a)
					Mr[a] = make_float4(m[4 * a], 4 * a + 1 < E::NC ? m[4 * a + 1] : 0.f, 4 * a + 2 < E::NC ? m[4 * a + 2] : 0.f,
					                    4 * a + 3 < E::NC ? m[4 * a + 3] : 0.f);
			}
		}
	}
}

template <int P>
__global__ void __launch_bounds__(128) k_m2m(const Ctrl* __restrict__ c, int l, const float4* __restrict__ geom,
                                             const uint2* __restrict__ info, float* __restrict__ M, int stride) {
	using E = Expansion<P>;
	const uint32_t lo = c->level_off[l], hi = c->level_off[l + 1];
	for (uint32_t node = lo + blockIdx.x * blockDim.x + threadIdx.x; node < hi; node += gridDim.x * blockDim.x) {
		const uint2 nf = info[node];
		if (nf.x == 0u) continue;
		const float4 g = geom[node];
		float mp[E::NC];
#pragma unroll
		for (int a = 0; a < E::NC; ++a) mp[a] = 0.0f;
#pragma unroll 1
		for (uint32_t k = 0; k < 8; ++k) {
			const uint32_t cid = nf.x + k;
			if (info[cid].y == 0u) continue;
			const float4 gc = geom[cid];
			float mc[E::NC];
			const float4* Mr = reinterpret_cast<const float4*>(M + (size_t) cid * stride);
#pragma unroll
			for (int a = 0; a < (E::NC + 3) / 4; ++a) {
				const float4 v = Mr[a];
				mc[4 * a] = v.x;
				if (4 * a + 1 < E::NC) mc[4 * a + 1] = v.y;
				if (4 * a + 2 < E::NC) mc[4 * a + 2] = v.z;
				if (4 * a + 3 < E::NC) mc[4 * a + 3] = v.w;
			}
			E::m2m(mp, mc, gc.x - g.x, gc.y - g.y, gc.z - g.z);
		}
		float4* Mw = reinterpret_cast<float4*>(M + (size_t) node * stride);
#pragma unroll
		for (int a = 0; a < (E::NC + 3) / 4; ++a)
			Mw[a] = make_float4(mp[4 * a], 4 * a + 1 < E::NC ? mp[4 * a + 1] : 0.f, 4 * a + 2 < E::NC ? mp[4 * a + 2] : 0.f,
			                    4 * a + 3 < E::NC ? mp[4 * a + 3] : 0.f);
	}
}

template <int P>
static void upsweep_t(Sim& s) {
	k_p2m<P><<<kNumSM * 8, 256, 0, s.stream>>>(s.ctrl, s.posq[1], s.geom, s.info, s.nbegin, s.M, s.L, s.nc_stride);
	for (int l = (s.depth_bound < (int) s.cfg.max_depth ? s.depth_bound : (int) s.cfg.max_depth) - 1; l >= 0; --l)
		k_m2m<P><<<kNumSM * 4, 128, 0, s.stream>>>(s.ctrl, l, s.geom, s.info, s.M, s.nc_stride);
}

void launch_upsweep(Sim& s) {
	switch (s.cfg.order) {
		case 2: upsweep_t<2>(s); break;
		case 3: upsweep_t<3>(s); break;
		case 5: upsweep_t<5>(s); break;
		default: upsweep_t<4>(s); break;
	}
}

}  // namespace nbody
