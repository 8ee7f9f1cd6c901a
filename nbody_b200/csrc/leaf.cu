// Stage 5: near field. One warp per target leaf: the leaf's P2P source list is
// expanded into a per-warp shared-memory tile of source particles (x,y,z,q) and the
// warp evaluates T targets x S source slices (T*S = 32, T = leaf population rounded
// up to a power of two) so that small leaves still fill the warp. The epilogue adds
// the far field (L2P) and applies the integrator, so accelerations never make a
// round trip through a per-interaction buffer.
//
// Replaces src/field.cl:49-148 (8x8 work-group per leaf interaction writing one
// 16-byte slot per (leaf, partner leaf, interaction)), src/force.cl:21-81 (slot
// reductions), the CPU prefix sums at src/open_cl_simulation.cpp:371-417 and the
// serial host integration loop at :572-616.
// Field of source j on target i: q_j (x_j - x_i) / (|x_j - x_i|^2 + eps^2)^(3/2)
// (src/field.cl:17-32 with FORCE_CONSTANT folded into force_constant).
// 20 flop per evaluation by the SURVEY 8d convention: 3 FADD, 3 FFMA, MUFU.RSQ (2), 3 FMUL, 3 FFMA.
#include "common.cuh"

namespace nbody {

constexpr int kLeafWarps = 4;
constexpr int kLeafTile = 256;  // source particles per tile; two tiles (8 KB) per warp

__device__ __forceinline__ void leaf_cp_async16(void* smem, const void* gmem) {
	const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void leaf_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void leaf_cp_async_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void leaf_cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

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
	const Ctrl* c;
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
	uint32_t batch_entries;       // source leaves staged per tile (batch_entries * quota <= kLeafTile)
	unsigned long long* stat_inter;
	unsigned long long* stat_leaves;
};

// Every lane walks its slice of the tile (sources sl, sl+S, ...) for one target, or two when TWO.
template <bool SOFT, bool TWO>
__device__ __forceinline__ void tile_compute(const float4* buf, uint32_t fill, unsigned sl, unsigned S, const float4& ta, const float4& tb,
                                             float eps2, float (&acc)[6]) {
	__syncwarp();
	uint32_t j = sl;
	for (; j + 3 * S < fill; j += 4 * S) {
		const float4 s0 = buf[j], s1 = buf[j + S], s2 = buf[j + 2 * S], s3 = buf[j + 3 * S];
		p2p_interact<SOFT>(s0, ta.x, ta.y, ta.z, eps2, acc[0], acc[1], acc[2]);
		if (TWO) p2p_interact<SOFT>(s0, tb.x, tb.y, tb.z, eps2, acc[3], acc[4], acc[5]);
		p2p_interact<SOFT>(s1, ta.x, ta.y, ta.z, eps2, acc[0], acc[1], acc[2]);
		if (TWO) p2p_interact<SOFT>(s1, tb.x, tb.y, tb.z, eps2, acc[3], acc[4], acc[5]);
		p2p_interact<SOFT>(s2, ta.x, ta.y, ta.z, eps2, acc[0], acc[1], acc[2]);
		if (TWO) p2p_interact<SOFT>(s2, tb.x, tb.y, tb.z, eps2, acc[3], acc[4], acc[5]);
		p2p_interact<SOFT>(s3, ta.x, ta.y, ta.z, eps2, acc[0], acc[1], acc[2]);
		if (TWO) p2p_interact<SOFT>(s3, tb.x, tb.y, tb.z, eps2, acc[3], acc[4], acc[5]);
	}
	for (; j < fill; j += S) {
		const float4 s0 = buf[j];
		p2p_interact<SOFT>(s0, ta.x, ta.y, ta.z, eps2, acc[0], acc[1], acc[2]);
		if (TWO) p2p_interact<SOFT>(s0, tb.x, tb.y, tb.z, eps2, acc[3], acc[4], acc[5]);
	}
	__syncwarp();
}

// One warp per target leaf. Its source list is a chain of segments of {first particle, count} entries;
// batches of `batch_entries` source leaves are expanded into a shared-memory tile with cp.async while the
// previous tile is being evaluated (two tiles per warp), and the entries of the batch after that are already
// in registers — so neither the list walk nor the particle fetch sits on the critical path.
template <int P, bool SOFT>
__global__ void __launch_bounds__(kLeafWarps * 32) k_leaf(const LeafArgs a) {
	using E = Expansion<P>;
	constexpr int STRIDE = coef_stride(P);
	__shared__ float4 sbuf[kLeafWarps][2][kLeafTile];
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	if (a.c->status) return;  // a pool overflowed: the host grows it and re-runs the step; leave the state untouched
	const uint32_t n_nodes = a.c->n_nodes;
	const uint32_t own_first = a.c->part[a.rank], own_end = a.c->part[a.rank + 1];
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t EB = a.batch_entries, quota = kLeafTile / EB;
	unsigned long long inter = 0, leaves = 0;
	for (uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; node < n_nodes; node += warps) {
		const uint2 nf = a.info[node];
		if (nf.x != 0u || nf.y == 0u) continue;  // internal or empty
		const uint32_t b = a.nbegin[node];
		if (b < own_first || b >= own_end) continue;  // another rank's leaf
		++leaves;
		const uint32_t nt = nf.y;
		const float4 g = a.geom[node];
		for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
			const uint32_t ntc = min(32u, nt - t0);
			// Warp layout: T target lanes x S source slices, one or two targets per lane, whichever wastes fewer
			// lanes: T = ntc (or ceil(ntc/2) with two targets per lane), S = floor(32 / T); lanes >= T*S stay idle.
			const unsigned T1 = ntc, S1 = 32u / T1, T2 = (ntc + 1u) / 2u, S2 = 32u / T2;
			// estimated issue slots per useful interaction: 17 / (ntc*S1/32) with one target per lane, 15 / (ntc*S2/64) with
			// two (one LDS.128 and one loop step feed two interactions): two targets per lane when 544*S2 > 960*S1
			const bool two = 544u * S2 > 960u * S1;
			const unsigned T = two ? T2 : T1, S = two ? S2 : S1;
			const unsigned sl_raw = lane / T, t = lane - sl_raw * T;
			const bool lane_on = lane < T * S;
			const unsigned sl = lane_on ? sl_raw : (unsigned) kLeafTile;  // idle lanes: an empty slice
			const bool has_a = lane_on && t < ntc, has_b = two && lane_on && t + T < ntc;
			float4 tp = make_float4(0.f, 0.f, 0.f, 0.f), tq = tp;
			if (has_a) tp = a.posq[b + t0 + t];
			if (has_b) tq = a.posq[b + t0 + t + T];
			float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
			unsigned long long nsrc = 0;
			// ---- cursor over the segment chain ----
			uint32_t si = a.p2p_head[node], e0 = 0;
			Segment sg; sg.off = 0; sg.cnt = 0; sg.next = 0xffffffffu;
			auto fetch = [&](uint2& ent) -> bool {  // next batch of <= EB entries, lane l holds entry l; warp-uniform result
				while (e0 >= sg.cnt) {
					if (si == 0xffffffffu) return false;
					sg = a.seg[si]; si = sg.next; e0 = 0;
				}
				ent = make_uint2(0u, 0u);
				if (lane < EB && e0 + lane < sg.cnt) ent = a.p2p[sg.off + e0 + lane];
				e0 += EB;
				return true;
			};
			auto stage = [&](const uint2& ent, float4* tile, uint32_t& fill, unsigned& big) {  // issue the tile fill of one batch
				const bool bigl = ent.y > quota;  // an over-full source leaf (only at max depth): streamed separately
				const uint32_t v = bigl ? 0u : ent.y;
				uint32_t inc = v;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
					if (lane >= (unsigned) d) inc += u;
				}
				fill = __shfl_sync(0xffffffffu, inc, 31);
				big = __ballot_sync(0xffffffffu, bigl);
				const uint32_t dst = inc - v, maxc = __reduce_max_sync(0xffffffffu, v);
				for (uint32_t k = 0; k < maxc; ++k)
					if (k < v) leaf_cp_async16(tile + dst + k, a.posq + ent.x + k);
				leaf_cp_async_commit();
			};
			uint2 ent_cur = make_uint2(0u, 0u), ent_nxt = make_uint2(0u, 0u);
			uint32_t fill_cur = 0;
			unsigned big_cur = 0;
			int cur = 0;
			bool has_cur = fetch(ent_cur);
			if (has_cur) stage(ent_cur, sbuf[w][0], fill_cur, big_cur);
			bool has_nxt = has_cur && fetch(ent_nxt);
			while (has_cur) {
				uint32_t fill_nxt = 0;
				unsigned big_nxt = 0;
				if (has_nxt) stage(ent_nxt, sbuf[w][cur ^ 1], fill_nxt, big_nxt);
				else leaf_cp_async_commit();  // empty group keeps the wait_group arithmetic uniform
				uint2 ent_nn = make_uint2(0u, 0u);
				const bool has_nn = has_nxt && fetch(ent_nn);  // entries of the batch after next: in flight during the math
				leaf_cp_async_wait1();
				if (two) tile_compute<SOFT, true>(sbuf[w][cur], fill_cur, sl, S, tp, tq, a.eps2, acc);
				else tile_compute<SOFT, false>(sbuf[w][cur], fill_cur, sl, S, tp, tq, a.eps2, acc);
				nsrc += fill_cur;
				while (big_cur) {  // stream an over-full source leaf through the tile that has just been consumed
					const int first = __ffs(big_cur) - 1;
					const uint32_t fb = __shfl_sync(0xffffffffu, ent_cur.x, first), fc = __shfl_sync(0xffffffffu, ent_cur.y, first);
					for (uint32_t q0 = 0; q0 < fc; q0 += kLeafTile) {
						const uint32_t m = min((uint32_t) kLeafTile, fc - q0);
						for (uint32_t q = lane; q < m; q += 32) sbuf[w][cur][q] = a.posq[fb + q0 + q];
						if (two) tile_compute<SOFT, true>(sbuf[w][cur], m, sl, S, tp, tq, a.eps2, acc);
						else tile_compute<SOFT, false>(sbuf[w][cur], m, sl, S, tp, tq, a.eps2, acc);
					}
					nsrc += fc;
					big_cur &= big_cur - 1;
				}
				ent_cur = ent_nxt; fill_cur = fill_nxt; big_cur = big_nxt; has_cur = has_nxt;
				ent_nxt = ent_nn; has_nxt = has_nn;
				cur ^= 1;
			}
			leaf_cp_async_wait0();
			// sum over the S slices into the lanes of slice 0 (lane t collects lanes t + s*T)
#pragma unroll
			for (int q = 0; q < 6; ++q) {
				if (q >= 3 && !two) break;
				float v = acc[q];
				if ((T & (T - 1u)) == 0u) {  // T*S == 32: butterfly
					for (unsigned d = T; d < 32u; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
				} else {
					const float mine = lane_on ? v : 0.f;
					float sum = mine;
					for (unsigned sidx = 1; sidx < S; ++sidx) sum += __shfl_sync(0xffffffffu, mine, (lane + sidx * T) & 31u);
					v = sum;
				}
				acc[q] = v;
			}
			if (lane == 0) inter += nsrc * ntc;
			if (lane_on && sl_raw == 0) {
				// far field: L2P of this leaf's local expansion, then the integrator — for each target of the lane
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
#pragma unroll 1
				for (int r = 0; r < 2; ++r) {
					if (r == 0 ? !has_a : !has_b) continue;
					const float4 tt = r == 0 ? tp : tq;
					const float px = r == 0 ? acc[0] : acc[3], py = r == 0 ? acc[1] : acc[4], pz = r == 0 ? acc[2] : acc[5];
					float fx, fy, fz;
					E::l2p(l, tt.x - g.x, tt.y - g.y, tt.z - g.z, fx, fy, fz);
					const uint32_t i = b + t0 + t + (r ? T : 0u);
					const float4 vm = a.velm_in[i];
					const float sc = a.G * tt.w / vm.w;  // a = G q/m * field (src/force.cl:4-10, src/open_cl_simulation.cpp:602-604)
					const float axx = sc * (px + fx), ayy = sc * (py + fy), azz = sc * (pz + fz);
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
	if (s.cfg.softening > 0.0f) k_leaf<P, true><<<kNumSM * 7, kLeafWarps * 32, 0, s.stream>>>(a);
	else k_leaf<P, false><<<kNumSM * 7, kLeafWarps * 32, 0, s.stream>>>(a);
}

void launch_leaf(Sim& s) {
	LeafArgs a{};
	a.c = s.ctrl; a.posq = s.posq[1]; a.velm_in = s.velm[1]; a.posq_out = s.posq[0]; a.velm_out = s.velm[0]; a.acc = s.acc;
	a.geom = s.geom; a.info = s.info; a.nbegin = s.nbegin; a.p2p_head = s.p2p_head; a.seg = s.pools.seg; a.p2p = s.pools.p2p; a.L = s.L;
	a.eps2 = s.cfg.softening * s.cfg.softening; a.G = s.cfg.force_constant; a.dt = s.cfg.time_step;
	a.integrator = (int) s.cfg.integrator; a.no_integrate = (s.cfg.flags & NBODY_FLAG_NO_INTEGRATE) ? 1 : 0;
	a.rank = s.rank;
	{  // batch_entries * leaf_capacity <= kLeafTile, at most one entry per lane
		uint32_t eb = kLeafTile / (s.cfg.leaf_capacity ? s.cfg.leaf_capacity : 1u);
		a.batch_entries = eb < 1u ? 1u : (eb > 32u ? 32u : eb);
	}
	a.stat_inter = &s.ctrl->stat_p2p_inter; a.stat_leaves = &s.ctrl->stat_leaves;
	switch (s.cfg.order) {
		case 2: leaf_t<2>(s, a); break;
		case 3: leaf_t<3>(s, a); break;
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

__global__ void k_direct_finish(uint64_t n, const float4* __restrict__ field, const float4* __restrict__ posq, const float4* __restrict__ velm,
                                float4* __restrict__ posq_out, float4* __restrict__ velm_out, float4* __restrict__ acc, float G, float dt,
                                int integrator, int no_integrate) {
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
	    s.n, field, s.posq[1], s.velm[1], s.posq[0], s.velm[0], s.acc, s.cfg.force_constant, s.cfg.time_step, (int) s.cfg.integrator,
	    (s.cfg.flags & NBODY_FLAG_NO_INTEGRATE) ? 1 : 0);
}

}  // namespace nbody
