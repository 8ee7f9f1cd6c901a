// Partitioned mode (NBODY_FLAG_PARTITIONED): every rank holds ONLY the particles whose Morton key lies in its key range, builds the
// octree of those particles, and imports what it needs of the other ranks — a locally essential tree. This is SURVEY 8e proper
// (the reference is single-device, src/open_cl_simulation.cpp:627-632; its only answer to "N does not fit" is the batch loop at
// :81-98, broken by D9); comm.cu keeps the replicated scheme of round 1 for comparison.
//
// The construction that makes it small: the octree is a fixed hierarchy of cells, and whether a cell splits depends on its particle
// count alone. A cell whose key range lies inside one rank's key range is built by that rank exactly as a single GPU would build
// it. A cell that contains a splitter key strictly inside its range ("straddling": at most one per splitter and level, <= 21 (W-1)
// in total) holds particles of several ranks: the ranks all-gather their local counts of those cells and split them by the GLOBAL
// count (tree.cu: node_splits). Every rank's tree is then the global octree restricted to the cells that hold its own particles,
// and the dual traversal of a rank's own targets against (own tree + every other rank's tree), seeded with all the roots, produces
// exactly the global interaction lists for those targets: the MAC and the splitting rule (src/interaction.cl:64-82) read only cell
// geometry and has_children, which agree on every rank; a straddling source cell simply appears once per rank that holds part of
// it, with that part's multipole and particles (M2L and P2P are linear in the sources).
//
// One step, rank r (X = exchange with the other ranks; everything else is local):
//   1   keys of the own particles, radix sort, gather into the sorted arrays [1]; positions of the splitters in the sorted keys
//   X1  all-gather of the cut rows (W+1 words): everybody learns how many particles go from s to d
//   2a  pull: the particles that now belong to r are read straight out of the owners' sorted arrays over NVLink (peer pointers from
//       cudaIpc handles; no packing, no NCCL), run by run into the state arrays [0]; keys again, stable merge of the W sorted runs
//       (sort.cu: merge path), local counts of the straddling cells
//   X2  all-gather of the straddling counts (also the barrier after which the sorted arrays [1] may be overwritten)
//   2b  gather into [1]; octree with forced splits; P2M / M2M
//   X3  all-gather of the node counts, then the trees as the traversal reads them: geometry, child/count records and
//       first-particle indices of every rank (28 bytes per node) by three ncclAllGather into equal slots behind the own tree
//       (ids [max_nodes, ...); grouped ncclBroadcast when a rank's arrays are smaller than the largest tree); child pointers are
//       rebased, first-particle indices become "imported leaf" tags. The multipoles do not travel here.
//   3   traversal (seeds: all roots; it marks every imported leaf / node its lists name), then the halo: the marked leaves get
//       slots behind the own particles (prefix sum) and their particles are fetched from the owners' sorted arrays with NVLink
//       loads (8f rank 4), the list entries are pointed at the slots; the multipoles the M2L lists name are fetched from the
//       owners' compact exports the same way; M2L, L2L, P2P + L2P + integrator as on one GPU
//   X4  all-gather of {device time of stage 3, status}; X5 all-gather of the next step's splitter candidates (balance.h rule on the
//       device: the owner of each wanted boundary position looks up the key there)
// Host synchronisations per step: after X1 (run lengths), after X3's counts, at the end. NCCL carries control words and the trees;
// particles move by peer loads only.
//
// "Virtual ranks": the same code runs W ranks inside one process on one GPU (nbody_cuda_create_group / nbody_cuda_group_step): peer
// pointers are the other objects' arrays, the all-gathers are device-to-device copies, all members share one stream and the phases
// above are executed member by member between the exchange points. That is how the parity tests exercise 2, 4 and 8 ranks on the
// one-GPU test box.
#include <algorithm>
#include <cstring>
#include <vector>

#include "balance.h"
#include "common.cuh"
#include "merge_path.h"
#include "nccl_api.h"

namespace nbody {

int create_for_let(const nbody_cuda_config* cfg, uint64_t n, uint64_t cap, uint64_t halo, Sim** out);  // api.cu
void destroy_for_comm(Sim* s);                                                                         // api.cu
int grow_pools_after_overflow(Sim& s, uint32_t status);                                                // api.cu
int realloc_nodes(Sim& s, uint32_t max_nodes, uint32_t src_nodes);                                     // api.cu
float next_time_step(const nbody_cuda_config& cfg, float acc_max);                                     // checkpoint.cu

namespace {

constexpr int kSample = 512;                 // keys per rank for the initial splitters
constexpr uint64_t kKeyEnd = 1ull << 63;      // one past the largest Morton key
constexpr uint64_t kNoKey = ~0ull;

// What a rank publishes in the small all-gathers (one slot per rank, always sent whole).
struct XSlot {
	uint32_t cut[kMaxRanks + 1];               // X1: position of splitter d in this rank's sorted keys
	uint32_t n_nodes, n_levels, status;        // X3 / X4
	uint32_t node_cap;                         // X3: capacity of the own tree's arrays
	uint32_t acc_max2_bits;                    // X4 (variable time step)
	uint32_t straddle[kMaxRanks][kNumLevels];  // X2: local count of the depth-d cell that straddles splitter b
	unsigned long long own_ns;                 // X4: device time of the stages whose cost follows the partition
	uint64_t cand[kMaxRanks + 1];              // X5: key at the wanted boundary position k, if this rank owns that position
	// the rank's compact multipole export (orders 0..P-1 per own node): where the others fetch the multipoles their M2L lists name.
	// Written by the host when the buffer is (re)allocated; `gen` tells the readers to map it again.
	cudaIpcMemHandle_t mexp_handle;
	unsigned long long mexp_ptr;               // virtual ranks: the pointer itself
	uint32_t mexp_gen, _pad;
	uint64_t sample[kSample];                  // creation: evenly spaced sample of the sorted keys
};

struct PeerTable {
	const float4* posq[kMaxRanks];
	const float4* velm[kMaxRanks];
	const uint32_t* orig[kMaxRanks];
};
struct PullArgs {
	PeerTable peer;
	uint32_t src_off[kMaxRanks];
	uint32_t bound[kMaxRanks + 1];
	int world;
};
struct ImportTable { uint32_t off[kMaxRanks + 1]; int world; };  // imported node i belongs to rank s: off[s] <= i < off[s+1]
struct MexpTable { const float* p[kMaxRanks]; };
struct CountTable { uint32_t n[kMaxRanks]; };

}  // namespace

struct Let {
	int rank = 0, world = 1;
	bool virt = false;
	Sim* member[kMaxRanks] = {};     // virtual ranks: every rank's object; one process per GPU: only member[rank]
	ncclComm_t comm = nullptr;
	XSlot* xbuf = nullptr;           // device, world slots
	XSlot* xhost = nullptr;          // pinned mirror
	LetCtrl* lc_host = nullptr;      // pinned
	PeerTable peer{};                // every rank's sorted arrays [1] (own entry: own arrays)
	void* mapped[3 * kMaxRanks] = {};
	uint32_t* imp_rbegin = nullptr;  // per imported node: first particle in its owner's sorted array
	uint32_t* imp_hoff = nullptr;    // per imported node (+1): halo particles before it; before the scan: its count if a P2P list names it
	uint32_t* imp_mflag = nullptr;   // per imported node: named by an M2L list of this step
	uint32_t imp_alloc = 0;
	float* mexp = nullptr;           // own nodes' multipoles, orders 0..P-1, compact: what the other ranks fetch from
	uint32_t mexp_cap = 0, mexp_gen = 0;
	MexpTable peer_mexp{};           // the other ranks' exports (mapped through cudaIpc, or the members' pointers)
	void* mexp_mapped[kMaxRanks] = {};
	uint32_t mexp_seen[kMaxRanks] = {};
	uint32_t counts[kMaxRanks] = {}; // particles per rank in the current step
	uint32_t nodes[kMaxRanks] = {};
	ImportTable imp{};
	uint32_t imp_total = 0, imp_slot = 0;
	bool gather_ok = true;           // every rank's own arrays hold a whole slot: the trees travel by plain all-gather
	uint64_t global_first = 0;
	cudaEvent_t ev[12] = {};
	uint64_t pulled = 0;             // particles that arrived from other ranks in the last step
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_key(const uint64_t* __restrict__ keys, uint32_t n, uint64_t k) {
	uint32_t a = 0, z = n;
	while (a < z) { const uint32_t m = a + ((z - a) >> 1); if (keys[m] < k) a = m + 1; else z = m; }
	return a;
}

__global__ void k_let_cuts(const uint64_t* __restrict__ keys, uint32_t n, const LetCtrl* __restrict__ lc, XSlot* slot) {
	const int d = threadIdx.x;
	if (d > lc->world) return;
	slot->cut[d] = d == lc->world ? n : lower_bound_key(keys, n, lc->split[d]);
}

// The particles of the W source runs, concatenated in rank order: run s = elements [src_off[s], ...) of rank s's sorted arrays.
__global__ void k_let_pull(const PullArgs a, float4* __restrict__ posq, float4* __restrict__ velm, uint32_t* __restrict__ orig) {
	const uint32_t n = a.bound[a.world];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		int s = 0;
		while (s + 1 < a.world && i >= a.bound[s + 1]) ++s;
		const uint32_t j = a.src_off[s] + (i - a.bound[s]);
		posq[i] = a.peer.posq[s][j];
		velm[i] = a.peer.velm[s][j];
		orig[i] = a.peer.orig[s][j];
	}
}

// Local particle counts of the cells that straddle a splitter: block b-1 = splitter b, thread = depth.
__global__ void k_let_straddle(const uint64_t* __restrict__ keys, uint32_t n, const LetCtrl* __restrict__ lc, XSlot* slot) {
	const int b = blockIdx.x + 1, d = threadIdx.x;
	if (d >= kNumLevels) return;
	const uint64_t K = lc->split[b];
	const int sh = 3 * (kMaxDepth - d);
	const uint64_t lo = (K >> sh) << sh;
	uint32_t cnt = 0;
	if (K != lo && K < kKeyEnd) {  // K lies strictly inside the cell
		const uint64_t hi = lo + (1ull << sh);
		cnt = lower_bound_key(keys, n, hi) - lower_bound_key(keys, n, lo);
	}
	slot->straddle[b][d] = cnt;
}

__global__ void k_let_force(const XSlot* __restrict__ xb, LetCtrl* lc) {
	const int W = lc->world;
	for (int t = threadIdx.x; t < kMaxRanks * kNumLevels; t += blockDim.x) {
		const int b = t / kNumLevels, d = t - b * kNumLevels;
		uint32_t sum = 0;
		if (b >= 1 && b < W)
			for (int s = 0; s < W; ++s) sum += xb[s].straddle[b][d];
		lc->force[b][d] = sum;
	}
}

__global__ void k_let_publish(const Ctrl* __restrict__ c, uint32_t node_cap, XSlot* slot) {
	slot->n_nodes = c->n_nodes; slot->n_levels = c->n_levels; slot->status = c->status; slot->node_cap = node_cap;
}

// Imported trees: child pointers are relative to the owner's array -> rebase; the first-particle index is kept aside (it addresses
// the OWNER's sorted array) and replaced by a tag the traversal copies into the P2P entries.
__global__ void k_let_fixup(const ImportTable t, uint32_t imp_base, uint32_t total, uint2* __restrict__ info, uint32_t* __restrict__ nbegin,
                            uint32_t* __restrict__ rbegin, uint32_t* __restrict__ hoff) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= total; i += gridDim.x * blockDim.x) {
		hoff[i] = 0u;
		if (i == total) break;
		int s = 0;
		while (s + 1 < t.world && i >= t.off[s + 1]) ++s;
		uint2 nf = info[imp_base + i];
		if (nf.x) { nf.x += imp_base + t.off[s]; info[imp_base + i] = nf; }
		rbegin[i] = nbegin[imp_base + i];
		nbegin[imp_base + i] = kImported | i;
	}
}

// hoff has been scanned: leaf i owns halo slots [hoff[i], hoff[i+1]). One warp per 32 imported nodes; the particles of each used
// leaf are copied by the whole warp from its owner's sorted array (peer memory: NVLink loads).
__global__ void k_let_fetch(Ctrl* c, LetCtrl* lc, const PeerTable peer, const ImportTable t, uint32_t total, const uint32_t* __restrict__ hoff,
                            const uint32_t* __restrict__ rbegin, float4* __restrict__ posq, uint32_t halo_base, uint32_t halo_cap) {
	const unsigned lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t need = hoff[total];
	if (blockIdx.x == 0 && threadIdx.x == 0) { lc->halo_count = need; if (need > halo_cap) atomicOr(&c->status, kOvfHalo); }
	if (need > halo_cap || (c->status & ~kOvfHalo)) return;
	for (uint32_t base = warp * 32u; base < total; base += nwarps * 32u) {
		const uint32_t i = base + lane;
		uint32_t off = 0, cnt = 0, rb = 0;
		if (i < total) { off = hoff[i]; cnt = hoff[i + 1] - off; rb = rbegin[i]; }
		unsigned m = __ballot_sync(0xffffffffu, cnt != 0u);
		while (m) {
			const int k = __ffs(m) - 1;
			m &= m - 1u;
			const uint32_t o = __shfl_sync(0xffffffffu, off, k), cn = __shfl_sync(0xffffffffu, cnt, k), r0 = __shfl_sync(0xffffffffu, rb, k);
			const uint32_t ii = base + (uint32_t) k;
			int s = 0;
			while (s + 1 < t.world && ii >= t.off[s + 1]) ++s;
			const float4* src = peer.posq[s] + r0;
			for (uint32_t q = lane; q < cn; q += 32u) posq[halo_base + o + q] = src[q];
		}
	}
}

__global__ void k_let_translate(const Ctrl* __restrict__ c, uint64_t p2p_cap, uint2* __restrict__ p2p, const uint32_t* __restrict__ hoff,
                                uint32_t halo_base) {
	if (c->status) return;
	const uint64_t n = c->p2p_cursor < p2p_cap ? c->p2p_cursor : p2p_cap;
	for (uint64_t e = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; e < n; e += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t x = p2p[e].x;
		if (x & kImported) p2p[e].x = halo_base + hoff[x & ~kImported];
	}
}

// Own multipoles -> compact export: the first MS floats (orders 0..P-1) of every record.
__global__ void k_let_pack_m(const Ctrl* __restrict__ c, const float4* __restrict__ M4, int s4, float4* __restrict__ E4, int ms4, uint32_t cap) {
	const uint32_t n = c->n_nodes < cap ? c->n_nodes : cap;
	const uint64_t total = (uint64_t) n * ms4;
	for (uint64_t t = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; t < total; t += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t node = (uint32_t) (t / ms4), j = (uint32_t) (t - (uint64_t) node * ms4);
		E4[t] = M4[(size_t) node * s4 + j];
	}
}

// The multipoles the M2L lists name, from their owners' exports (peer memory: NVLink loads), into the local array the M2L kernel reads.
__global__ void k_let_fetch_m(const Ctrl* __restrict__ c, const MexpTable peer, const ImportTable t, uint32_t total, const uint32_t* __restrict__ mflag,
                              float4* __restrict__ Mimp4, int ms4) {
	if (c->status) return;
	const uint64_t n = (uint64_t) total * ms4;
	for (uint64_t q = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; q < n; q += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t i = (uint32_t) (q / ms4), j = (uint32_t) (q - (uint64_t) i * ms4);
		if (!mflag[i]) continue;
		int s = 0;
		while (s + 1 < t.world && i >= t.off[s + 1]) ++s;
		Mimp4[q] = reinterpret_cast<const float4*>(peer.p[s])[(size_t) (i - t.off[s]) * ms4 + j];
	}
}

__global__ void k_stamp(unsigned long long* dst) {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	*dst = t;
}

__global__ void k_let_finish(const Ctrl* __restrict__ c, const LetCtrl* __restrict__ lc, XSlot* slot) {
	slot->status = c->status;
	slot->own_ns = lc->t_end - lc->t_begin;
	slot->acc_max2_bits = c->acc_max2_bits;
	slot->n_nodes = c->n_nodes; slot->n_levels = c->n_levels;
}

__device__ bool let_targets(const XSlot* __restrict__ xb, int W, const CountTable& ct, uint32_t* part, uint32_t* tgt) {
	float work[kMaxRanks];
	part[0] = 0;
	for (int s = 0; s < W; ++s) {
		if (xb[s].status) return false;  // somebody repeats part of the step: the partition is decided after the last attempt
		part[s + 1] = part[s] + ct.n[s];
		work[s] = (float) xb[s].own_ns * 1e-6f;
	}
	rebalance_boundaries(W, part, work, 0.5f, tgt);
	return true;
}

// Per-step load rebalancing (balance.h): the wanted boundary positions of the next step in the global tree order; the rank that
// holds position tgt[k] publishes the key there.
__global__ void k_let_rebalance(const XSlot* __restrict__ xb, XSlot* mine, const LetCtrl* __restrict__ lc, const CountTable ct,
                                const uint64_t* __restrict__ keys) {
	if (threadIdx.x || blockIdx.x) return;
	const int W = lc->world, r = lc->rank;
	for (int k = 0; k <= W; ++k) mine->cand[k] = kNoKey;
	uint32_t part[kMaxRanks + 1], tgt[kMaxRanks + 1];
	if (!let_targets(xb, W, ct, part, tgt)) return;
	for (int k = 1; k < W; ++k)
		if (tgt[k] >= part[r] && tgt[k] < part[r + 1]) mine->cand[k] = keys[tgt[k] - part[r]];
}

__global__ void k_let_adopt(const XSlot* __restrict__ xb, LetCtrl* lc, const CountTable ct) {
	if (threadIdx.x || blockIdx.x) return;
	const int W = lc->world;
	uint32_t part[kMaxRanks + 1], tgt[kMaxRanks + 1];
	if (!let_targets(xb, W, ct, part, tgt)) return;
	for (int k = 1; k < W; ++k) {
		uint64_t K = kNoKey;
		for (int s = 0; s < W; ++s) if (xb[s].cand[k] != kNoKey) K = xb[s].cand[k];
		if (K == kNoKey) K = tgt[k] >= part[W] ? kKeyEnd : lc->split[k];
		if (K < lc->split[k - 1]) K = lc->split[k - 1];
		lc->split[k] = K;
	}
}

// A repeat of stage 3 after a list pool has grown: the tree and the imported trees stay, the lists start from scratch.
__global__ void k_reset_lists(Ctrl* c, uint32_t* __restrict__ p2p_head) {
	const uint32_t n = c->n_nodes;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p2p_head[i] = 0xffffffffu;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		c->status = 0;
		c->gq_count[0] = c->gq_count[1] = 0; c->items_count[0] = c->items_count[1] = 0;
		c->seg_cursor = 0; c->near_cursor[0] = c->near_cursor[1] = 0; c->p2p_cursor = 0; c->m2l_cursor = 0;
		c->stat_m2l_inter = c->stat_m2l_low = c->stat_p2p_entries = c->stat_p2p_inter = c->stat_near = c->stat_leaves = 0;
		for (int k = 0; k < 4; ++k) c->work_ticket[k] = 0;
		c->acc_max2_bits = 0;
		c->n_leaf_items = 0;
	}
}

__global__ void k_let_sample(const uint64_t* __restrict__ keys, uint32_t n, XSlot* slot) {
	const uint32_t j = threadIdx.x;
	if (j < (uint32_t) kSample) slot->sample[j] = n ? keys[(uint64_t) j * n / kSample] : kNoKey;
	if (j == 0) slot->cut[0] = n;
}

inline int grid_of(uint64_t n, int block) {
	const uint64_t want = (n + block - 1) / block, cap = (uint64_t) kNumSM * 16;
	return (int) (want < 1 ? 1 : (want > cap ? cap : want));
}

// ---------------------------------------------------------------------------------------------------------------------------
// exchanges
// ---------------------------------------------------------------------------------------------------------------------------
// All-gather of the exchange slots of the `nm` local members (virtual ranks: all of them, device-to-device; otherwise NCCL).
int xchg(Sim** m, int nm, bool to_host) {
	Let& L0 = *m[0]->let;
	const int W = L0.world;
	if (L0.virt) {
		for (int d = 0; d < nm; ++d)
			for (int s = 0; s < nm; ++s)
				if (s != d) NB_CUDA_CHECK(cudaMemcpyAsync(m[d]->let->xbuf + s, m[s]->let->xbuf + s, sizeof(XSlot), cudaMemcpyDeviceToDevice, m[0]->stream));
	} else {
		Let& L = *m[0]->let;
		NB_NCCL_CHECK(g_nccl.AllGather(L.xbuf + L.rank, L.xbuf, sizeof(XSlot), ncclChar, L.comm, m[0]->stream));
	}
	if (to_host)
		for (int d = 0; d < nm; ++d)
			NB_CUDA_CHECK(cudaMemcpyAsync(m[d]->let->xhost, m[d]->let->xbuf, sizeof(XSlot) * W, cudaMemcpyDeviceToHost, m[d]->stream));
	return NBODY_OK;
}

int sync_all(Sim** m, int nm) {
	for (int i = 0; i < nm; ++i) {
		NB_CUDA_CHECK(cudaStreamSynchronize(m[i]->stream));
		if (m[0]->let->virt) break;  // one shared stream
	}
	NB_CUDA_CHECK(cudaGetLastError());
	return NBODY_OK;
}

// Room for `need` imported nodes behind the own tree; the own part of the source-side arrays is kept.
int ensure_import_room(Sim& s, uint32_t need, uint32_t own_nodes) {
	Let& L = *s.let;
	if (s.src_nodes - s.max_nodes < need) {
		const uint64_t want = (uint64_t) s.max_nodes + need + need / 4 + 1024;
		if (want > 0x7fffffffull) { set_error("partitioned mode: more than 2^31 nodes on one rank"); return NBODY_ERR_CAPACITY; }
		float4* geom = nullptr; uint2* info = nullptr; uint32_t* nbegin = nullptr;
		if (cudaMalloc((void**) &geom, want * sizeof(float4)) != cudaSuccess || cudaMalloc((void**) &info, want * sizeof(uint2)) != cudaSuccess ||
		    cudaMalloc((void**) &nbegin, want * 4) != cudaSuccess) {
			cudaFree(geom); cudaFree(info); cudaFree(nbegin);
			set_error("partitioned mode: no memory for the imported trees");
			return NBODY_ERR_CUDA;
		}
		NB_CUDA_CHECK(cudaMemcpyAsync(geom, s.geom, own_nodes * sizeof(float4), cudaMemcpyDeviceToDevice, s.stream));
		NB_CUDA_CHECK(cudaMemcpyAsync(info, s.info, own_nodes * sizeof(uint2), cudaMemcpyDeviceToDevice, s.stream));
		NB_CUDA_CHECK(cudaMemcpyAsync(nbegin, s.nbegin, own_nodes * 4, cudaMemcpyDeviceToDevice, s.stream));
		NB_CUDA_CHECK(cudaStreamSynchronize(s.stream));
		cudaFree(s.geom); cudaFree(s.info); cudaFree(s.nbegin);
		s.device_bytes += (want - s.src_nodes) * (sizeof(float4) + sizeof(uint2) + 4);
		s.geom = geom; s.info = info; s.nbegin = nbegin;
		s.src_nodes = (uint32_t) want;
	}
	if (L.imp_alloc < need + 1) {
		const uint32_t want = need + need / 4 + 1024;
		const size_t per = 12 + (size_t) coef_stride((int) s.cfg.order - 1) * 4;  // rbegin, hoff, mflag, Mimp record
		if (L.imp_rbegin) { cudaFree(L.imp_rbegin); cudaFree(L.imp_hoff); cudaFree(L.imp_mflag); cudaFree(s.Mimp); s.device_bytes -= (uint64_t) L.imp_alloc * per; }
		L.imp_rbegin = L.imp_hoff = L.imp_mflag = nullptr; s.Mimp = nullptr;
		if (cudaMalloc((void**) &L.imp_rbegin, (size_t) want * 4) != cudaSuccess || cudaMalloc((void**) &L.imp_hoff, (size_t) want * 4) != cudaSuccess ||
		    cudaMalloc((void**) &L.imp_mflag, (size_t) want * 4) != cudaSuccess || cudaMalloc((void**) &s.Mimp, (size_t) want * (per - 12)) != cudaSuccess) {
			set_error("partitioned mode: no memory for the import tables");
			return NBODY_ERR_CUDA;
		}
		L.imp_alloc = want;
		s.device_bytes += (uint64_t) want * per;
	}
	return NBODY_OK;
}

// X3, second half: every rank's tree as the traversal reads it (geometry, child/count record, first particle: 28 bytes per node) lands
// behind the own tree. The multipoles do not travel here: after the traversal each rank fetches exactly the ones its M2L lists name.
int exchange_trees(Sim** m, int nm) {
	NvtxRange r("partitioned: all-gather of the trees");
	Let& L0 = *m[0]->let;
	const int W = L0.world;
	if (L0.virt) {
		for (int d = 0; d < nm; ++d) {
			Sim& D = *m[d];
			const Let& L = *D.let;
			for (int s = 0; s < nm; ++s) {
				if (s == d || L.nodes[s] == 0) continue;
				const Sim& S = *m[s];
				const size_t at = (size_t) D.max_nodes + L.imp.off[s], cnt = L.nodes[s];
				NB_CUDA_CHECK(cudaMemcpyAsync(D.geom + at, S.geom, cnt * sizeof(float4), cudaMemcpyDeviceToDevice, D.stream));
				NB_CUDA_CHECK(cudaMemcpyAsync(D.info + at, S.info, cnt * sizeof(uint2), cudaMemcpyDeviceToDevice, D.stream));
				NB_CUDA_CHECK(cudaMemcpyAsync(D.nbegin + at, S.nbegin, cnt * 4, cudaMemcpyDeviceToDevice, D.stream));
			}
		}
		return NBODY_OK;
	}
	Sim& D = *m[0];
	const Let& L = *D.let;
	if (L.gather_ok) {
		const size_t slot = L.imp_slot, at = D.max_nodes;
		NB_NCCL_CHECK(g_nccl.GroupStart());
		NB_NCCL_CHECK(g_nccl.AllGather(D.geom, D.geom + at, slot * sizeof(float4), ncclChar, L.comm, D.stream));
		NB_NCCL_CHECK(g_nccl.AllGather(D.info, D.info + at, slot * sizeof(uint2), ncclChar, L.comm, D.stream));
		NB_NCCL_CHECK(g_nccl.AllGather(D.nbegin, D.nbegin + at, slot * 4, ncclChar, L.comm, D.stream));
		NB_NCCL_CHECK(g_nccl.GroupEnd());
		return NBODY_OK;
	}
	NB_NCCL_CHECK(g_nccl.GroupStart());  // fallback: one rank's tree is larger than another rank's arrays
	for (int s = 0; s < W; ++s) {
		const size_t cnt = L.nodes[s];
		if (cnt == 0) continue;
		const size_t at = s == L.rank ? 0 : (size_t) D.max_nodes + L.imp.off[s];
		NB_NCCL_CHECK(g_nccl.Broadcast(D.geom + at, D.geom + at, cnt * sizeof(float4), ncclChar, s, L.comm, D.stream));
		NB_NCCL_CHECK(g_nccl.Broadcast(D.info + at, D.info + at, cnt * sizeof(uint2), ncclChar, s, L.comm, D.stream));
		NB_NCCL_CHECK(g_nccl.Broadcast(D.nbegin + at, D.nbegin + at, cnt * 4, ncclChar, s, L.comm, D.stream));
	}
	NB_NCCL_CHECK(g_nccl.GroupEnd());
	return NBODY_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// phases (each enqueues on the member's stream; the driver below places the exchanges between them)
// ---------------------------------------------------------------------------------------------------------------------------
int phase1(Sim& s) {
	NvtxRange r("partitioned: keys + sort + cuts");
	Let& L = *s.let;
	NB_CUDA_CHECK(cudaEventRecord(L.ev[0], s.stream));
	launch_keys(s, s.posq[0], s.n);
	launch_own_sort_range(s, 0, s.n);
	launch_gather(s, s.n);
	k_let_cuts<<<1, 32, 0, s.stream>>>(s.keys[0], (uint32_t) s.n, s.let_ctrl, L.xbuf + L.rank);
	return NBODY_OK;
}

int phase2a(Sim& s) {
	NvtxRange nvtx("partitioned: pull migrants + merge");
	Let& L = *s.let;
	const int W = L.world, r = L.rank;
	PullArgs pa{};
	pa.peer = L.peer;
	pa.world = W;
	uint64_t nn = 0;
	for (int q = 0; q < W; ++q) {
		const uint32_t off = L.xhost[q].cut[r], cnt = L.xhost[q].cut[r + 1] - off;
		pa.src_off[q] = off;
		pa.bound[q] = (uint32_t) nn;
		nn += cnt;
	}
	pa.bound[W] = (uint32_t) nn;
	uint64_t first = 0;
	for (int d = 0; d < W; ++d) {
		uint64_t c = 0;
		for (int q = 0; q < W; ++q) c += L.xhost[q].cut[d + 1] - L.xhost[q].cut[d];
		L.counts[d] = (uint32_t) c;
		if (d < r) first += c;
	}
	if (nn > s.cap) {
		set_error("partitioned mode: " + std::to_string(nn) + " particles migrate to rank " + std::to_string(r) + ", room for " +
		          std::to_string(s.cap) + " (raise pool_scale)");
		return NBODY_ERR_CAPACITY;
	}
	L.pulled = nn - (L.xhost[r].cut[r + 1] - L.xhost[r].cut[r]);
	L.global_first = first;
	s.n = nn; s.own_first = first; s.own_count = nn;
	if (nn) k_let_pull<<<grid_of(nn, 256), 256, 0, s.stream>>>(pa, s.posq[0], s.velm[0], s.orig[0]);
	launch_keys(s, s.posq[0], nn);
	launch_merge_runs(s, pa.bound, W);
	if (W > 1) k_let_straddle<<<W - 1, 32, 0, s.stream>>>(s.keys[0], (uint32_t) nn, s.let_ctrl, L.xbuf + r);
	return NBODY_OK;
}

// The compact multipole export holds max_nodes records; (re)allocated with the node arrays, its handle published through the slot.
int ensure_export(Sim& s) {
	Let& L = *s.let;
	if (L.mexp && L.mexp_cap == s.max_nodes) return NBODY_OK;
	const size_t rec = (size_t) coef_stride((int) s.cfg.order - 1) * 4;
	if (L.mexp) { cudaFree(L.mexp); s.device_bytes -= (uint64_t) L.mexp_cap * rec; L.mexp = nullptr; }
	if (cudaMalloc((void**) &L.mexp, (size_t) s.max_nodes * rec) != cudaSuccess) { set_error("partitioned mode: no memory for the multipole export"); return NBODY_ERR_CUDA; }
	L.mexp_cap = s.max_nodes;
	s.device_bytes += (uint64_t) L.mexp_cap * rec;
	XSlot& h = L.xhost[L.rank];  // staging: only these fields of the own slot are host-written
	std::memset(&h.mexp_handle, 0, sizeof(h.mexp_handle));
	if (!L.virt && cudaIpcGetMemHandle(&h.mexp_handle, L.mexp) != cudaSuccess) {
		set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError()));
		return NBODY_ERR_CUDA;
	}
	h.mexp_ptr = (unsigned long long) (uintptr_t) L.mexp;
	h.mexp_gen = ++L.mexp_gen;
	XSlot* d = L.xbuf + L.rank;
	NB_CUDA_CHECK(cudaMemcpyAsync(&d->mexp_handle, &h.mexp_handle, sizeof(h.mexp_handle), cudaMemcpyHostToDevice, s.stream));
	NB_CUDA_CHECK(cudaMemcpyAsync(&d->mexp_ptr, &h.mexp_ptr, sizeof(h.mexp_ptr), cudaMemcpyHostToDevice, s.stream));
	NB_CUDA_CHECK(cudaMemcpyAsync(&d->mexp_gen, &h.mexp_gen, sizeof(h.mexp_gen), cudaMemcpyHostToDevice, s.stream));
	NB_CUDA_CHECK(cudaStreamSynchronize(s.stream));  // the staging words live in the pinned mirror the next read-back overwrites
	return NBODY_OK;
}

int build_tree(Sim& s) {
	NvtxRange r("partitioned: octree + upsweep + export");
	Let& L = *s.let;
	NB_CUDA_CHECK(cudaEventRecord(L.ev[1], s.stream));
	launch_tree_build(s);
	NB_CUDA_CHECK(cudaEventRecord(L.ev[2], s.stream));
	launch_upsweep(s);
	int rc = ensure_export(s);
	if (rc) return rc;
	const int ms4 = coef_stride((int) s.cfg.order - 1) / 4;
	k_let_pack_m<<<kNumSM * 4, 256, 0, s.stream>>>(s.ctrl, reinterpret_cast<const float4*>(s.M), s.nc_stride / 4, reinterpret_cast<float4*>(L.mexp), ms4, L.mexp_cap);
	k_let_publish<<<1, 1, 0, s.stream>>>(s.ctrl, s.max_nodes, L.xbuf + L.rank);
	NB_CUDA_CHECK(cudaEventRecord(L.ev[3], s.stream));
	return NBODY_OK;
}

int phase2b(Sim& s) {
	Let& L = *s.let;
	k_let_force<<<1, 512, 0, s.stream>>>(L.xbuf, s.let_ctrl);
	launch_gather(s, s.n);
	if (s.n) NB_CUDA_CHECK(cudaMemcpyAsync(s.orig[0], s.orig[1], s.n * 4, cudaMemcpyDeviceToDevice, s.stream));
	return build_tree(s);
}

// After X3's counts are on the host: where the other ranks' trees go.
int plan_import(Sim& s) {
	Let& L = *s.let;
	const int W = L.world;
	// every rank's tree gets a slot of `maxc` nodes behind the own tree (the own slot stays unused): equal slots make the exchange
	// three plain all-gathers instead of 3 W grouped broadcasts
	uint32_t deepest = 1, maxc = 0;
	L.imp.world = W;
	L.gather_ok = true;
	for (int q = 0; q < W; ++q) {
		L.nodes[q] = L.xhost[q].n_nodes;
		deepest = std::max(deepest, L.xhost[q].n_levels);
		maxc = std::max(maxc, L.nodes[q]);
	}
	for (int q = 0; q < W; ++q) if (L.xhost[q].node_cap < maxc) L.gather_ok = false;  // (a sender reads maxc records of its own arrays)
	const uint64_t total = (uint64_t) W * maxc;
	for (int q = 0; q <= W; ++q) L.imp.off[q] = (uint32_t) ((uint64_t) q * maxc);
	L.imp_slot = maxc;
	if (total > 0x7ffffff0ull) { set_error("partitioned mode: more than 2^31 imported nodes"); return NBODY_ERR_CAPACITY; }
	L.imp_total = (uint32_t) total;
	s.trav_bound = (int) deepest - 1 > 0 ? (int) deepest - 1 : 1;
	for (int q = 0; q < W; ++q) {  // the other ranks' multipole exports: map again when a rank has reallocated its buffer
		const XSlot& x = L.xhost[q];
		if (q == L.rank) { L.peer_mexp.p[q] = L.mexp; continue; }
		if (x.mexp_gen == L.mexp_seen[q] && L.peer_mexp.p[q]) continue;
		if (L.virt) L.peer_mexp.p[q] = reinterpret_cast<const float*>((uintptr_t) x.mexp_ptr);
		else {
			if (L.mexp_mapped[q]) { cudaIpcCloseMemHandle(L.mexp_mapped[q]); L.mexp_mapped[q] = nullptr; }
			void* p = nullptr;
			const cudaError_t e = cudaIpcOpenMemHandle(&p, x.mexp_handle, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) { set_error(std::string("cudaIpcOpenMemHandle (multipoles of rank ") + std::to_string(q) + "): " + cudaGetErrorString(e)); return NBODY_ERR_CUDA; }
			L.mexp_mapped[q] = p;
			L.peer_mexp.p[q] = static_cast<const float*>(p);
		}
		L.mexp_seen[q] = x.mexp_gen;
	}
	return ensure_import_room(s, L.imp_total, L.nodes[L.rank]);
}

int stage3(Sim& s, bool retry) {
	NvtxRange r("partitioned: traversal + halo + M2L + L2L + P2P");
	Let& L = *s.let;
	const int W = L.world;
	cudaStream_t st = s.stream;
	if (retry) {
		k_reset_lists<<<kNumSM * 4, 256, 0, st>>>(s.ctrl, s.p2p_head);
		launch_upsweep(s);  // P2M also clears the local expansions the failed attempt accumulated into
		NB_CUDA_CHECK(cudaMemsetAsync(L.imp_hoff, 0, ((size_t) L.imp_total + 1) * 4, st));
	} else {
		k_let_fixup<<<grid_of((uint64_t) L.imp_total + 1, 256), 256, 0, st>>>(L.imp, s.max_nodes, L.imp_total, s.info, s.nbegin, L.imp_rbegin, L.imp_hoff);
	}
	if (L.imp_total) NB_CUDA_CHECK(cudaMemsetAsync(L.imp_mflag, 0, (size_t) L.imp_total * 4, st));
	s.imp_hoff = L.imp_hoff; s.imp_mflag = L.imp_mflag;
	s.seeds.n = 1; s.seeds.id[0] = 0;
	for (int q = 0; q < W; ++q)
		if (q != L.rank && L.nodes[q] && L.counts[q]) s.seeds.id[s.seeds.n++] = s.max_nodes + L.imp.off[q];
	NB_CUDA_CHECK(cudaEventRecord(L.ev[4], st));
	k_stamp<<<1, 1, 0, st>>>(&s.let_ctrl->t_begin);
	launch_traversal(s);
	NB_CUDA_CHECK(cudaEventRecord(L.ev[5], st));
	// halo: which imported leaves do the P2P lists name -> slots behind the own particles -> fetch over NVLink -> point the entries at them
	const uint32_t halo_base = (uint32_t) s.cap, halo_cap = (uint32_t) (s.src_cap - s.cap);
	launch_exclusive_scan(s, L.imp_hoff, L.imp_total + 1);  // (the traversal stored every named imported leaf's count there)
	k_let_fetch<<<kNumSM * 8, 256, 0, st>>>(s.ctrl, s.let_ctrl, L.peer, L.imp, L.imp_total, L.imp_hoff, L.imp_rbegin, s.posq[1], halo_base, halo_cap);
	k_let_translate<<<kNumSM * 8, 256, 0, st>>>(s.ctrl, s.pools.p2p_cap, s.pools.p2p, L.imp_hoff, halo_base);
	// ... and the multipoles: which imported nodes do the M2L lists name -> fetch orders 0..P-1 of exactly those from their owners
	if (L.imp_total) {
		const int ms4 = coef_stride((int) s.cfg.order - 1) / 4;
		k_let_fetch_m<<<kNumSM * 8, 256, 0, st>>>(s.ctrl, L.peer_mexp, L.imp, L.imp_total, L.imp_mflag, reinterpret_cast<float4*>(s.Mimp), ms4);
	}
	NB_CUDA_CHECK(cudaEventRecord(L.ev[6], st));
	launch_m2l(s);
	NB_CUDA_CHECK(cudaEventRecord(L.ev[7], st));
	launch_l2l(s);
	NB_CUDA_CHECK(cudaEventRecord(L.ev[8], st));
	launch_leaf(s);
	if (s.cfg.time_step_eta > 0.0f) launch_acc_max(s);
	k_stamp<<<1, 1, 0, st>>>(&s.let_ctrl->t_end);
	NB_CUDA_CHECK(cudaEventRecord(L.ev[9], st));
	k_let_finish<<<1, 1, 0, st>>>(s.ctrl, s.let_ctrl, L.xbuf + L.rank);
	return NBODY_OK;
}

int end_of_step(Sim& s, int part) {  // the two halves around X5
	Let& L = *s.let;
	CountTable ct{};
	for (int q = 0; q < L.world; ++q) ct.n[q] = L.counts[q];
	if (s.cfg.flags & NBODY_FLAG_STATIC_PARTITION) return NBODY_OK;
	if (part == 0) k_let_rebalance<<<1, 32, 0, s.stream>>>(L.xbuf, L.xbuf + L.rank, s.let_ctrl, ct, s.keys[0]);
	else k_let_adopt<<<1, 32, 0, s.stream>>>(L.xbuf, s.let_ctrl, ct);
	return NBODY_OK;
}

void fill_stats(Sim& s) {
	Let& L = *s.let;
	const Ctrl& c = *s.ctrl_host;
	nbody_cuda_stats& t = s.stats;
	t.n_particles = s.n; t.n_nodes = c.n_nodes; t.n_levels = c.n_levels; t.n_leaves = c.stat_leaves;
	t.m2l_entries = c.m2l_cursor; t.m2l_interactions = c.stat_m2l_inter; t.m2l_interactions_low = c.stat_m2l_low; t.p2p_entries = c.stat_p2p_entries;
	t.p2p_interactions = c.stat_p2p_inter >= s.n ? c.stat_p2p_inter - s.n : 0;  // drop the i == j terms
	t.near_entries = c.stat_near; t.device_bytes = s.device_bytes;
	auto ms = [&](int a, int b) { float v = 0; cudaEventElapsedTime(&v, L.ev[a], L.ev[b]); return v; };
	t.ms_sort = ms(0, 1); t.ms_tree = ms(1, 2); t.ms_upsweep = ms(2, 3); t.ms_traverse = ms(4, 5); t.ms_m2l = ms(6, 7); t.ms_l2l = ms(7, 8);
	t.ms_leaf = ms(8, 9); t.ms_import = ms(3, 4); t.ms_halo = ms(5, 6); t.ms_balance = ms(9, 10);
	t.ms_comm = t.ms_import + t.ms_halo + t.ms_balance; t.ms_total = ms(0, 10);
	float mx = 0.0f, sum = 0.0f;
	for (int q = 0; q < L.world; ++q) { const float w = (float) L.xhost[q].own_ns * 1e-6f; mx = std::max(mx, w); sum += w; }
	t.work_imbalance = sum > 0.0f ? mx * L.world / sum - 1.0f : 0.0f;
	t.halo_particles = L.lc_host->halo_count; t.imported_nodes = L.imp_total; t.migrated_particles = L.pulled;
}

// One step of the `nm` local members (all ranks of a virtual group, or this process's one rank).
int step_members(Sim** m, int nm) {
	Let& L0 = *m[0]->let;
	const int W = L0.world;
	int rc;
#define EACH(call) for (int i = 0; i < nm; ++i) { Sim& s = *m[i]; (void) s; if ((rc = (call))) return rc; }
	EACH((s.stats.retries = 0, phase1(s)));
	if ((rc = xchg(m, nm, true)) || (rc = sync_all(m, nm))) return rc;                      // X1
	EACH(phase2a(s));
	if ((rc = xchg(m, nm, false))) return rc;                                              // X2
	EACH(phase2b(s));
	for (int attempt = 0;; ++attempt) {                                                    // X3: node counts (and the tree build's overflow bits)
		if ((rc = xchg(m, nm, true)) || (rc = sync_all(m, nm))) return rc;
		bool any = false;
		for (int q = 0; q < W; ++q) any |= L0.xhost[q].status != 0;
		if (!any) break;
		if (attempt >= 24) { set_error("partitioned step: the node pool still overflows after 24 growth attempts"); return NBODY_ERR_CAPACITY; }
		for (int i = 0; i < nm; ++i) {
			Sim& s = *m[i];
			const uint32_t status = s.let->xhost[s.let->rank].status;
			if (!status) continue;
			if ((rc = grow_pools_after_overflow(s, status & (kOvfNodes | kOvfDepth)))) return rc;
			++s.stats.retries;
			if ((rc = build_tree(s))) return rc;
		}
	}
	EACH(plan_import(s));
	if ((rc = exchange_trees(m, nm))) return rc;
	EACH(stage3(s, false));
	for (int attempt = 0;; ++attempt) {
		if ((rc = xchg(m, nm, false))) return rc;                                            // X4
		EACH(end_of_step(s, 0));
		if ((rc = xchg(m, nm, true))) return rc;                                             // X5
		EACH(end_of_step(s, 1));
		for (int i = 0; i < nm; ++i) {
			Sim& s = *m[i];
			NB_CUDA_CHECK(cudaMemcpyAsync(s.ctrl_host, s.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s.stream));
			NB_CUDA_CHECK(cudaMemcpyAsync(s.let->lc_host, s.let_ctrl, sizeof(LetCtrl), cudaMemcpyDeviceToHost, s.stream));
			NB_CUDA_CHECK(cudaEventRecord(s.let->ev[10], s.stream));
		}
		if ((rc = sync_all(m, nm))) return rc;
		bool any = false;
		for (int q = 0; q < W; ++q) any |= L0.xhost[q].status != 0;
		if (!any) break;
		if (attempt >= 24) { set_error("partitioned step: pools still overflow after 24 growth attempts"); return NBODY_ERR_CAPACITY; }
		for (int i = 0; i < nm; ++i) {
			Sim& s = *m[i];
			const uint32_t status = s.let->xhost[s.let->rank].status;
			if (!status) continue;
			if (status & kOvfHalo) {
				set_error("partitioned mode: " + std::to_string(s.let->lc_host->halo_count) + " halo particles, room for " +
				          std::to_string(s.src_cap - s.cap) + " (raise pool_scale)");
				return NBODY_ERR_CAPACITY;
			}
			if ((rc = grow_pools_after_overflow(s, status & ~(kOvfNodes | kOvfDepth)))) return rc;
			++s.stats.retries;
			if ((rc = stage3(s, true))) return rc;
		}
	}
#undef EACH
	// bookkeeping of a completed step (api.cu does the same for the single-GPU path)
	float a2max = 0.0f;
	bool a2nan = false;
	for (int q = 0; q < W; ++q) {
		float a2;
		std::memcpy(&a2, &L0.xhost[q].acc_max2_bits, sizeof(float));
		if (a2 != a2) a2nan = true; else a2max = std::max(a2max, a2);
	}
	for (int i = 0; i < nm; ++i) {
		Sim& s = *m[i];
		s.time += s.dt;
		++s.steps_done;
		s.dt_last = s.dt;
		if (s.cfg.time_step_eta > 0.0f) {
			s.acc_max = a2nan ? std::sqrt(-1.0f) : std::sqrt(a2max);
			s.dt = next_time_step(s.cfg, s.acc_max);
		}
		s.depth_bound = std::min<int>((int) s.cfg.max_depth, (int) s.ctrl_host->n_levels);
		s.lists_valid = false;
		fill_stats(s);
	}
	return NBODY_OK;
}

void release(Sim* s) {
	if (!s) return;
	destroy_for_comm(s);  // calls let_destroy
}

int attach(Sim& s, int rank, int world, bool virt) {
	Let* L = new Let;
	s.let = L;
	L->rank = rank; L->world = world; L->virt = virt;
	s.rank = 0;  // the kernels' "own slice" is the whole local particle array
	if (cudaMalloc((void**) &L->xbuf, sizeof(XSlot) * world) != cudaSuccess || cudaMallocHost((void**) &L->xhost, sizeof(XSlot) * world) != cudaSuccess ||
	    cudaMalloc((void**) &s.let_ctrl, sizeof(LetCtrl)) != cudaSuccess || cudaMallocHost((void**) &L->lc_host, sizeof(LetCtrl)) != cudaSuccess) {
		set_error("partitioned mode: allocation of the exchange buffers failed");
		return NBODY_ERR_CUDA;
	}
	cudaMemset(L->xbuf, 0, sizeof(XSlot) * world);
	std::memset(L->xhost, 0, sizeof(XSlot) * world);
	for (auto& e : L->ev) if (cudaEventCreate(&e) != cudaSuccess) { set_error("event creation failed"); return NBODY_ERR_CUDA; }
	return NBODY_OK;
}

// Initial splitters from the all-gathered key samples (in xhost): weighted quantiles, equal particle counts per rank.
void initial_splitters(Let& L, uint64_t* split) {
	const int W = L.world;
	std::vector<std::pair<uint64_t, double>> smp;
	double total = 0.0;
	for (int q = 0; q < W; ++q) {
		const double n = (double) L.xhost[q].cut[0];
		if (n <= 0) continue;
		for (int j = 0; j < kSample; ++j) smp.push_back({L.xhost[q].sample[j], n / kSample});
		total += n;
	}
	std::stable_sort(smp.begin(), smp.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
	split[0] = 0; split[W] = kKeyEnd;
	size_t at = 0;
	double cum = 0.0;
	for (int k = 1; k < W; ++k) {
		const double goal = total * k / W;
		while (at < smp.size() && cum + smp[at].second <= goal) cum += smp[at++].second;
		split[k] = at < smp.size() ? smp[at].first : kKeyEnd;
		if (split[k] < split[k - 1]) split[k] = split[k - 1];
	}
}

// Common tail of creation: sorted key sample -> splitters on every member; peer tables.
int establish(Sim** m, int nm) {
	int rc;
	Let& L0 = *m[0]->let;
	const int W = L0.world;
	for (int i = 0; i < nm; ++i) {
		Sim& s = *m[i];
		Let& L = *s.let;
		launch_keys(s, s.posq[0], s.n);
		launch_own_sort_range(s, 0, s.n);
		k_let_sample<<<1, kSample, 0, s.stream>>>(s.keys[0], (uint32_t) s.n, L.xbuf + L.rank);
	}
	if ((rc = xchg(m, nm, true)) || (rc = sync_all(m, nm))) return rc;
	for (int i = 0; i < nm; ++i) {
		Sim& s = *m[i];
		Let& L = *s.let;
		LetCtrl h{};
		initial_splitters(L, h.split);
		h.world = W; h.rank = L.rank; h.imp_base = s.max_nodes;
		h.halo_base = (uint32_t) s.cap; h.halo_cap = (uint32_t) (s.src_cap - s.cap);
		uint64_t first = 0;
		for (int q = 0; q < W; ++q) { L.counts[q] = L.xhost[q].cut[0]; if (q < L.rank) first += L.counts[q]; }
		L.global_first = first; s.own_first = first; s.own_count = s.n;
		NB_CUDA_CHECK(cudaMemcpy(s.let_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice));
	}
	return NBODY_OK;
}

}  // namespace

void let_destroy(Sim& s) {
	if (!s.let) return;
	Let* L = s.let;
	for (void* p : L->mapped) if (p) cudaIpcCloseMemHandle(p);
	if (L->comm) g_nccl.CommDestroy(L->comm);
	if (L->xbuf) cudaFree(L->xbuf);
	if (L->xhost) cudaFreeHost(L->xhost);
	if (L->lc_host) cudaFreeHost(L->lc_host);
	if (s.let_ctrl) cudaFree(s.let_ctrl);
	if (L->imp_rbegin) cudaFree(L->imp_rbegin);
	if (L->imp_hoff) cudaFree(L->imp_hoff);
	if (L->imp_mflag) cudaFree(L->imp_mflag);
	if (s.Mimp) { cudaFree(s.Mimp); s.Mimp = nullptr; }
	for (void* p : L->mexp_mapped) if (p) cudaIpcCloseMemHandle(p);
	if (L->mexp) cudaFree(L->mexp);
	for (auto& e : L->ev) if (e) cudaEventDestroy(e);
	if (L->virt && L->rank != 0) s.stream = nullptr;  // the group's stream belongs to member 0
	delete L;
	s.let = nullptr; s.let_ctrl = nullptr;
}

int let_step(Sim& s) {
	if (s.let->virt) { set_error("a member of a virtual group steps with nbody_cuda_group_step"); return NBODY_ERR_STATE; }
	Sim* m[1] = {&s};
	return step_members(m, 1);
}

// Room per rank: the own particles may grow through migration and rebalancing, and the sources include the halo.
static void let_room(const nbody_cuda_config* cfg, uint64_t n_global, int world, uint64_t n_local, uint64_t* cap, uint64_t* halo) {
	const double sc = cfg->pool_scale > 0 ? cfg->pool_scale : 1.0;
	const uint64_t share = (n_global + world - 1) / world;
	// automatic slack: 50 % up to 2^24 particles per rank, shrinking above (a large rank's share moves by a smaller fraction per step)
	const double slack = cfg->partition_slack_pct ? 0.01 * cfg->partition_slack_pct : 0.5 * std::min(1.0, 16777216.0 / (double) std::max<uint64_t>(share, 1));
	*cap = std::max<uint64_t>(n_local, (uint64_t) (sc * (1.0 + slack) * (double) share)) + 4096;
	*halo = world > 1 ? (uint64_t) (sc * std::min(2.0, 4.0 * (slack + 0.05)) * (double) share) + 262144 : 16;
}

int let_create_distributed(const nbody_cuda_config* cfg, const nbody_particle* local_particles, uint64_t n_local, uint64_t n_global,
                           uint64_t global_offset, int rank, int world, const uint8_t* id, nbody_cuda_sim** out) {
	if (n_global > 0xfffffff0ull) { set_error("particle identities are 32-bit: at most 2^32-16 particles in total"); return NBODY_ERR_INVALID; }
	if (cfg->flags & NBODY_FLAG_DIRECT) { set_error("NBODY_FLAG_DIRECT (all-pairs validation path) is not available in partitioned mode"); return NBODY_ERR_INVALID; }
	uint64_t cap = 0, halo = 0;
	let_room(cfg, n_global, world, n_local, &cap, &halo);
	if (cap + halo >= 0x7ffffff0ull) { set_error("partitioned mode: more than 2^31 particles (own + halo) on one rank"); return NBODY_ERR_INVALID; }
	Sim* s = nullptr;
	int rc = create_for_let(cfg, n_local ? n_local : 1, cap, halo, &s);
	if (rc) return rc;
	s->n = n_local; s->n_global = n_global;
	auto fail = [&](int code) { release(s); return code; };
	if ((rc = attach(*s, rank, world, false))) return fail(rc);
	Let& L = *s->let;
	ncclUniqueId u;
	std::memcpy(&u, id, 128);
	ncclResult_t nr = g_nccl.CommInitRank(&L.comm, world, u, rank);
	if (nr != ncclSuccess) { set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(nr)); L.comm = nullptr; return fail(NBODY_ERR_COMM); }
	if (n_local) {
		if (cudaMemcpyAsync(s->aos_dev, local_particles, n_local * sizeof(nbody_particle), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) {
			set_error("upload failed"); return fail(NBODY_ERR_CUDA);
		}
		launch_import_ids(*s, s->aos_dev, n_local, (uint32_t) global_offset, false);
	}
	cudaMemsetAsync(s->acc, 0, s->cap * sizeof(float4), s->stream);
	// peer pointers: the other ranks' sorted arrays, mapped through cudaIpc handles that travel in one all-gather
	{
		struct Handles { cudaIpcMemHandle_t h[3]; };
		std::vector<Handles> all(world);
		Handles mine{};
		if (cudaIpcGetMemHandle(&mine.h[0], s->posq[1]) != cudaSuccess || cudaIpcGetMemHandle(&mine.h[1], s->velm[1]) != cudaSuccess ||
		    cudaIpcGetMemHandle(&mine.h[2], s->orig[1]) != cudaSuccess) {
			set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError()));
			return fail(NBODY_ERR_CUDA);
		}
		Handles* d = nullptr;
		if (cudaMalloc((void**) &d, sizeof(Handles) * world) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(NBODY_ERR_CUDA); }
		cudaMemcpyAsync(d + rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, s->stream);
		nr = g_nccl.AllGather(d + rank, d, sizeof(Handles), ncclChar, L.comm, s->stream);
		cudaMemcpyAsync(all.data(), d, sizeof(Handles) * world, cudaMemcpyDeviceToHost, s->stream);
		cudaStreamSynchronize(s->stream);
		cudaFree(d);
		if (nr != ncclSuccess) { set_error(std::string("ncclAllGather: ") + g_nccl.GetErrorString(nr)); return fail(NBODY_ERR_COMM); }
		for (int q = 0; q < world; ++q) {
			if (q == rank) { L.peer.posq[q] = s->posq[1]; L.peer.velm[q] = s->velm[1]; L.peer.orig[q] = s->orig[1]; continue; }
			for (int k = 0; k < 3; ++k) {
				void* p = nullptr;
				const cudaError_t e = cudaIpcOpenMemHandle(&p, all[q].h[k], cudaIpcMemLazyEnablePeerAccess);
				if (e != cudaSuccess) {
					set_error(std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(q) + "): " + cudaGetErrorString(e) +
					          " — partitioned mode needs CUDA peer access between the GPUs of the box");
					return fail(NBODY_ERR_CUDA);
				}
				L.mapped[3 * q + k] = p;
			}
			L.peer.posq[q] = (const float4*) L.mapped[3 * q]; L.peer.velm[q] = (const float4*) L.mapped[3 * q + 1];
			L.peer.orig[q] = (const uint32_t*) L.mapped[3 * q + 2];
		}
	}
	Sim* m[1] = {s};
	L.member[rank] = s;
	if ((rc = establish(m, 1))) return fail(rc);
	*out = reinterpret_cast<nbody_cuda_sim*>(s);
	return NBODY_OK;
}

}  // namespace nbody

using namespace nbody;

extern "C" {

int nbody_cuda_create_group(const nbody_cuda_config* cfg, const nbody_particle* particles, uint64_t n, int world, nbody_cuda_sim** sims_out) {
	if (!cfg || !particles || !sims_out) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (world < 1 || world > kMaxRanks) { set_error("bad world size (1..16 ranks)"); return NBODY_ERR_INVALID; }
	if (cfg->flags & NBODY_FLAG_DIRECT) { set_error("NBODY_FLAG_DIRECT (all-pairs validation path) is not available in partitioned mode"); return NBODY_ERR_INVALID; }
	for (int r = 0; r < world; ++r) sims_out[r] = nullptr;
	Sim* m[kMaxRanks] = {};
	auto fail = [&](int code) { for (int r = world - 1; r >= 0; --r) if (m[r]) release(m[r]); return code; };
	int rc;
	for (int r = 0; r < world; ++r) {
		const uint64_t lo = n * r / world, hi = n * (r + 1) / world, nl = hi - lo;
		uint64_t cap = 0, halo = 0;
		let_room(cfg, n, world, nl, &cap, &halo);
		if ((rc = create_for_let(cfg, nl ? nl : 1, cap, halo, &m[r]))) return fail(rc);
		Sim& s = *m[r];
		s.n = nl; s.n_global = n;
		if (r > 0) { cudaStreamDestroy(s.stream); s.stream = m[0]->stream; }  // one stream for the whole group: the phases are ordered by it
		if ((rc = attach(s, r, world, true))) return fail(rc);
		if (nl) {
			if (cudaMemcpyAsync(s.aos_dev, particles + lo, nl * sizeof(nbody_particle), cudaMemcpyHostToDevice, s.stream) != cudaSuccess) {
				set_error("upload failed"); return fail(NBODY_ERR_CUDA);
			}
			launch_import_ids(s, s.aos_dev, nl, (uint32_t) lo, false);
		}
		cudaMemsetAsync(s.acc, 0, s.cap * sizeof(float4), s.stream);
	}
	for (int r = 0; r < world; ++r)
		for (int q = 0; q < world; ++q) {
			Let& L = *m[r]->let;
			L.member[q] = m[q];
			L.peer.posq[q] = m[q]->posq[1]; L.peer.velm[q] = m[q]->velm[1]; L.peer.orig[q] = m[q]->orig[1];
		}
	if ((rc = establish(m, world))) return fail(rc);
	for (int r = 0; r < world; ++r) sims_out[r] = reinterpret_cast<nbody_cuda_sim*>(m[r]);
	return NBODY_OK;
}

int nbody_cuda_group_step(nbody_cuda_sim** sims, int world, float* time_out) {
	if (!sims || world < 1 || world > kMaxRanks) { set_error("bad argument"); return NBODY_ERR_INVALID; }
	Sim* m[kMaxRanks];
	for (int r = 0; r < world; ++r) {
		m[r] = reinterpret_cast<Sim*>(sims[r]);
		if (!m[r] || !m[r]->let || !m[r]->let->virt || m[r]->let->world != world || m[r]->let->rank != r) {
			set_error("group_step: not the members of one virtual group, in rank order");
			return NBODY_ERR_INVALID;
		}
	}
	NB_CUDA_CHECK(cudaSetDevice(m[0]->device));
	const int rc = step_members(m, world);
	if (rc == NBODY_OK && time_out) *time_out = m[0]->time;
	return rc;
}

void nbody_cuda_destroy_group(nbody_cuda_sim** sims, int world) {
	if (!sims) return;
	for (int r = world - 1; r >= 0; --r) { release(reinterpret_cast<Sim*>(sims[r])); sims[r] = nullptr; }  // member 0 (the stream's owner) last
}

}  // extern "C"
