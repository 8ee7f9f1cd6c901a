// Multi-GPU: one process per GPU, particles partitioned by Morton-key range.
//
// The reference is single-device (src/open_cl_simulation.cpp:627-632); this is new work
// (SURVEY 8e). Round-1 scheme = the survey's stated fallback: every rank keeps the full,
// identically ordered particle state and builds the same octree and multipoles (O(N),
// HBM-bound, a few per cent of a step), while the expensive stages — dual-tree
// traversal, M2L, P2P/L2P/integration — run only for the targets inside the rank's
// contiguous slice of the Morton-ordered particle array (slice boundaries snapped to
// leaf boundaries so a leaf never straddles two ranks). After the step the updated slices
// (position+charge, velocity+mass, acceleration) are exchanged with one grouped NCCL
// broadcast per rank over NVLink, which is the all-gather with unequal counts.
// Determinism of the radix sort and of the tree build makes the replicated structures
// bit-identical across ranks, so no tree data ever needs to travel.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "balance.h"
#include "common.cuh"
#include "nccl_api.h"

namespace nbody {

NcclApi g_nccl;

bool nccl_load() {
	if (g_nccl.ok) return true;
	void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) { set_error(std::string("cannot load libnccl.so.2: ") + dlerror()); return false; }
	auto sym = [&](const char* name) { return dlsym(h, name); };
	g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId)) sym("ncclGetUniqueId");
	g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank)) sym("ncclCommInitRank");
	g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy)) sym("ncclCommDestroy");
	g_nccl.CommSplit = (decltype(g_nccl.CommSplit)) sym("ncclCommSplit");
	g_nccl.AllGather = (decltype(g_nccl.AllGather)) sym("ncclAllGather");
	g_nccl.Broadcast = (decltype(g_nccl.Broadcast)) sym("ncclBroadcast");
	g_nccl.GroupStart = (decltype(g_nccl.GroupStart)) sym("ncclGroupStart");
	g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd)) sym("ncclGroupEnd");
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString)) sym("ncclGetErrorString");
	g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.Broadcast &&
	            g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
	if (!g_nccl.ok) set_error("libnccl.so.2 lacks a required symbol");
	return g_nccl.ok;
}

int create_for_comm(const nbody_cuda_config* cfg, uint64_t n, Sim** out);
int let_create_distributed(const nbody_cuda_config* cfg, const nbody_particle* local_particles, uint64_t n_local, uint64_t n_global,
                           uint64_t global_offset, int rank, int world, const uint8_t* id, nbody_cuda_sim** out);  // let.cu
int comm_wait_velocities(Sim& s);
void destroy_for_comm(Sim* s);

struct Comm {
	ncclComm_t comm = nullptr;
	ncclComm_t vel_comm = nullptr;  // a duplicate of `comm` for the velocity stream (only with NBODY_FLAG_DIST_SORT, when ncclCommSplit exists)
	int rank = 0, world = 1;
	uint32_t* part_host = nullptr;  // pinned, world + 1: boundaries of the last step
	float* work_host = nullptr;     // pinned, world: per-rank device time of the owned-slice stages of the last step
	float* work_dev = nullptr;      // device, world
	bool have_work = false;         // work_host describes the partition in part_host
	float* red_host = nullptr;      // pinned, world: scratch of comm_all_max
	float* red_dev = nullptr;       // device, world
	// velocities travel on a second stream: the other ranks need them only in the next step's leaf kernel
	cudaStream_t vel_stream = nullptr;
	cudaEvent_t ev_step = nullptr, ev_vel = nullptr, ev_pos = nullptr;
	bool vel_pending = false;       // a velocity exchange was issued that the compute stream has not been ordered after yet
	bool pos_pending = false;       // the same for a position exchange on the second stream (only with vel_comm)
};

struct PartitionTargets { uint32_t t[kMaxRanks + 1]; };


// Boundary r = the wanted position, moved down to the first particle of the leaf that contains it.
__global__ void k_partition(Ctrl* c, int world, uint32_t n, const PartitionTargets want, const uint2* __restrict__ info,
                            const uint32_t* __restrict__ nbegin) {
	const int r = threadIdx.x;
	if (r > world) return;
	uint32_t p = want.t[r] < n ? want.t[r] : n;
	if (r == 0) p = 0;
	if (r == world || p >= n) { c->part[r] = n; return; }
	uint32_t node = 0;
	for (;;) {
		const uint2 nf = info[node];
		if (nf.x == 0) break;
		uint32_t next = nf.x;
		for (uint32_t k = 0; k < 8; ++k) {
			const uint32_t cb = nbegin[nf.x + k], cc = info[nf.x + k].y;
			if (p >= cb && p < cb + cc) { next = nf.x + k; break; }
		}
		node = next;
	}
	c->part[r] = nbegin[node];
}

// Enqueues the partition kernel; the kernels of the step read the boundaries from the control block on the device.
// First step (or NBODY_FLAG_STATIC_PARTITION): equal particle counts. Afterwards: the boundaries follow the measured
// per-rank work of the previous step (balance.h).
int comm_partition(Sim& s) {
	Comm& cm = *s.comm;
	PartitionTargets want{};
	if (cm.have_work && !(s.cfg.flags & NBODY_FLAG_STATIC_PARTITION)) {
		rebalance_boundaries(cm.world, cm.part_host, cm.work_host, 0.5f, want.t);
	} else {
		for (int r = 0; r <= cm.world; ++r) want.t[r] = (uint32_t) ((uint64_t) s.n * r / cm.world);
	}
	k_partition<<<1, 32, 0, s.stream>>>(s.ctrl, cm.world, (uint32_t) s.n, want, s.info, s.nbegin);
	return NBODY_OK;
}

// After the step's single synchronisation: the host learns the slice boundaries (they arrive with the control block)
// and uses them for the exchange.
void comm_adopt_partition(Sim& s) {
	Comm& cm = *s.comm;
	for (int r = 0; r <= cm.world; ++r) cm.part_host[r] = s.ctrl_host->part[r];
	s.own_first = cm.part_host[cm.rank];
	s.own_count = cm.part_host[cm.rank + 1] - cm.part_host[cm.rank];
}

static int exchange(Sim& s, void* buf, size_t elem_bytes, cudaStream_t st = nullptr, ncclComm_t comm = nullptr) {
	Comm& cm = *s.comm;
	if (!st) st = s.stream;
	if (!comm) comm = cm.comm;
	NB_NCCL_CHECK(g_nccl.GroupStart());
	for (int r = 0; r < cm.world; ++r) {
		const size_t off = (size_t) cm.part_host[r] * elem_bytes, cnt = (size_t) (cm.part_host[r + 1] - cm.part_host[r]) * elem_bytes;
		if (cnt == 0) continue;
		char* p = static_cast<char*>(buf) + off;
		NB_NCCL_CHECK(g_nccl.Broadcast(p, p, cnt, ncclChar, r, comm, st));
	}
	NB_NCCL_CHECK(g_nccl.GroupEnd());
	return NBODY_OK;
}

int comm_exchange_aos(Sim& s) { return exchange(s, s.aos_dev, sizeof(nbody_particle)); }

// Distributed sort (NBODY_FLAG_DIST_SORT). The slices are those of the last completed step (or of creation): every rank knows
// all boundaries, and the state array every rank holds is in that order.
void comm_own_slice(const Sim& s, uint64_t* first, uint64_t* count) {
	const Comm& cm = *s.comm;
	*first = cm.part_host[cm.rank];
	*count = cm.part_host[cm.rank + 1] - cm.part_host[cm.rank];
}

// All-gather of the slice-wise sorted (key, index) runs in keys[0] / idx[0], 12 bytes per particle, on the compute stream;
// returns the run boundaries for the merge rounds.
int comm_sort_exchange(Sim& s, uint32_t* bound, int* nruns) {
	Comm& cm = *s.comm;
	int rc;
	// The previous step's velocity broadcasts may still be running on the second stream, and two collectives of ONE communicator
	// must never be in flight at once. With a communicator of their own (vel_comm) they simply keep running; without one the
	// compute stream is ordered behind them first (the rank's own slice sort, enqueued before this point, still overlaps them).
	if (!cm.vel_comm && (rc = comm_wait_velocities(s))) return rc;
	if ((rc = exchange(s, s.keys[0], sizeof(uint64_t)))) return rc;
	if ((rc = exchange(s, s.idx[0], sizeof(uint32_t)))) return rc;
	for (int r = 0; r <= cm.world; ++r) bound[r] = cm.part_host[r];
	*nruns = cm.world;
	return NBODY_OK;
}

// `own_ms`: device time this rank spent on the stages it runs for its own slice only (the input of the next
// step's rebalancing); all-gathered next to the slices, on the host after the step's final synchronisation.
int comm_step_exchange(Sim& s, float own_ms) {
	Comm& cm = *s.comm;
	int rc;
	const bool overlap = !(s.cfg.flags & NBODY_FLAG_NO_OVERLAP);
	// positions first: every rank's next step starts with the keys of ALL particles — unless the distributed sort is on, where a
	// rank computes keys for its own slice only and the other ranks' positions are first read by the gather AFTER the sort: with a
	// communicator of its own the second stream then carries the positions as well, behind this step's kernels and overlapped with
	// the next step's slice sort, key all-gather and merge rounds (comm_wait_positions orders the gather after them).
	const bool pos_on_second = overlap && cm.vel_comm != nullptr && (s.cfg.flags & NBODY_FLAG_DIST_SORT) && !(s.cfg.flags & NBODY_FLAG_CUB_SORT);
	if (!pos_on_second && (rc = exchange(s, s.posq[0], sizeof(float4)))) return rc;
	if (!overlap && (rc = exchange(s, s.velm[0], sizeof(float4)))) return rc;
	s.acc_partial = true;  // accelerations stay rank-local until somebody asks for them (comm_exchange_acc)
	cm.work_host[cm.rank] = own_ms;
	NB_CUDA_CHECK(cudaMemcpyAsync(cm.work_dev + cm.rank, cm.work_host + cm.rank, sizeof(float), cudaMemcpyHostToDevice, s.stream));
	NB_NCCL_CHECK(g_nccl.AllGather(cm.work_dev + cm.rank, cm.work_dev, 1, ncclFloat, cm.comm, s.stream));
	NB_CUDA_CHECK(cudaMemcpyAsync(cm.work_host, cm.work_dev, sizeof(float) * cm.world, cudaMemcpyDeviceToHost, s.stream));
	cm.have_work = true;  // valid once the caller has synchronised the stream
	if (overlap) {
		// velocities: nobody reads another rank's velocities before the next step's leaf kernel (or an export), so their
		// broadcasts run on a second stream behind this step's kernels and overlap the next step's sort, tree build,
		// upsweep and traversal; the consumer orders itself after ev_vel (comm_wait_velocities).
		NB_CUDA_CHECK(cudaEventRecord(cm.ev_step, s.stream));
		NB_CUDA_CHECK(cudaStreamWaitEvent(cm.vel_stream, cm.ev_step, 0));
		if (pos_on_second) {
			if ((rc = exchange(s, s.posq[0], sizeof(float4), cm.vel_stream, cm.vel_comm))) return rc;
			NB_CUDA_CHECK(cudaEventRecord(cm.ev_pos, cm.vel_stream));
			cm.pos_pending = true;
		}
		if ((rc = exchange(s, s.velm[0], sizeof(float4), cm.vel_stream, cm.vel_comm))) return rc;
		NB_CUDA_CHECK(cudaEventRecord(cm.ev_vel, cm.vel_stream));
		cm.vel_pending = true;
	}
	return NBODY_OK;
}

// Orders the compute stream after the velocity exchange in flight, if any. Called before the first reader or writer of the
// velocity plane after a step: the next step's velocity gather, an export, an import.
int comm_wait_velocities(Sim& s) {
	Comm& cm = *s.comm;
	if (!cm.vel_pending) return NBODY_OK;
	NB_CUDA_CHECK(cudaStreamWaitEvent(s.stream, cm.ev_vel, 0));
	cm.vel_pending = false;
	cm.pos_pending = false;  // the positions travel ahead of the velocities on the same stream
	return NBODY_OK;
}

// Orders the compute stream after a position exchange in flight on the second stream, if any: before the first reader of other
// ranks' positions in the next step (the gather after the distributed sort). Exports and imports go through comm_wait_velocities.
int comm_wait_positions(Sim& s) {
	Comm& cm = *s.comm;
	if (!cm.pos_pending) return NBODY_OK;
	NB_CUDA_CHECK(cudaStreamWaitEvent(s.stream, cm.ev_pos, 0));
	cm.pos_pending = false;
	return NBODY_OK;
}

// max over ranks / mean over ranks - 1 of the last step's owned-slice work (0 on a single GPU or before the first step)
float comm_work_imbalance(const Sim& s) {
	if (!s.comm || !s.comm->have_work) return 0.0f;
	float mx = 0.0f, sum = 0.0f;
	for (int r = 0; r < s.comm->world; ++r) { mx = s.comm->work_host[r] > mx ? s.comm->work_host[r] : mx; sum += s.comm->work_host[r]; }
	return sum > 0.0f ? mx * s.comm->world / sum - 1.0f : 0.0f;
}

// Maximum of one float over the ranks (variable time step: the largest acceleration of the step). An all-gather of
// `world` floats and a host-side maximum, so that every rank derives the next step from identical bits. NaN on any
// rank gives NaN everywhere.
int comm_all_max(Sim& s, float* value) {
	Comm& cm = *s.comm;
	// the velocity broadcasts of this step may still be in flight on the second stream, possibly on the SAME communicator:
	// two collectives of one communicator must not overlap, so the compute stream is ordered behind them first
	{ const int rcw = comm_wait_velocities(s); if (rcw) return rcw; }
	cm.red_host[cm.rank] = *value;
	NB_CUDA_CHECK(cudaMemcpyAsync(cm.red_dev + cm.rank, cm.red_host + cm.rank, sizeof(float), cudaMemcpyHostToDevice, s.stream));
	NB_NCCL_CHECK(g_nccl.AllGather(cm.red_dev + cm.rank, cm.red_dev, 1, ncclFloat, cm.comm, s.stream));
	NB_CUDA_CHECK(cudaMemcpyAsync(cm.red_host, cm.red_dev, sizeof(float) * cm.world, cudaMemcpyDeviceToHost, s.stream));
	NB_CUDA_CHECK(cudaStreamSynchronize(s.stream));
	float m = cm.red_host[0];
	for (int r = 1; r < cm.world; ++r) m = (cm.red_host[r] > m || cm.red_host[r] != cm.red_host[r]) ? cm.red_host[r] : m;
	*value = m;
	return NBODY_OK;
}

int comm_exchange_acc(Sim& s) {
	if (!s.acc_partial) return NBODY_OK;
	{ const int rcw = comm_wait_velocities(s); if (rcw) return rcw; }  // (see comm_all_max)
	int rc = exchange(s, s.acc, sizeof(float4));
	if (rc) return rc;
	s.acc_partial = false;
	return NBODY_OK;
}

void comm_destroy(Sim& s) {
	if (!s.comm) return;
	if (s.comm->vel_stream) { cudaStreamSynchronize(s.comm->vel_stream); cudaStreamDestroy(s.comm->vel_stream); }
	if (s.comm->ev_step) cudaEventDestroy(s.comm->ev_step);
	if (s.comm->ev_vel) cudaEventDestroy(s.comm->ev_vel);
	if (s.comm->ev_pos) cudaEventDestroy(s.comm->ev_pos);
	if (s.comm->vel_comm) g_nccl.CommDestroy(s.comm->vel_comm);
	if (s.comm->comm) g_nccl.CommDestroy(s.comm->comm);
	if (s.comm->part_host) cudaFreeHost(s.comm->part_host);
	if (s.comm->work_host) cudaFreeHost(s.comm->work_host);
	if (s.comm->work_dev) cudaFree(s.comm->work_dev);
	if (s.comm->red_host) cudaFreeHost(s.comm->red_host);
	if (s.comm->red_dev) cudaFree(s.comm->red_dev);
	delete s.comm;
	s.comm = nullptr;
}

}  // namespace nbody

using namespace nbody;

extern "C" {

int nbody_cuda_rebalance(int world, const uint32_t* boundaries, const float* work_ms, float damping, uint32_t* out) {
	if (!boundaries || !work_ms || !out || world < 1 || world > 16) { set_error("bad argument"); return NBODY_ERR_INVALID; }
	rebalance_boundaries(world, boundaries, work_ms, damping, out);
	return NBODY_OK;
}

int nbody_cuda_comm_unique_id(uint8_t* id) {
	if (!id) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	if (!nccl_load()) return NBODY_ERR_COMM;
	ncclUniqueId u;
	NB_NCCL_CHECK(g_nccl.GetUniqueId(&u));
	std::memcpy(id, &u, 128);
	return NBODY_OK;
}

int nbody_cuda_create_distributed(const nbody_cuda_config* cfg, const nbody_particle* local_particles, uint64_t n_local,
                                  uint64_t n_global, uint64_t global_offset, int rank, int world, const uint8_t* id,
                                  nbody_cuda_sim** out) {
	if (!out || !local_particles || !id) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	*out = nullptr;
	if (world < 1 || world > 16 || rank < 0 || rank >= world) { set_error("bad rank / world size (1..16 ranks)"); return NBODY_ERR_INVALID; }
	if (global_offset + n_local > n_global) { set_error("local slice exceeds the global particle count"); return NBODY_ERR_INVALID; }
	if (!nccl_load()) return NBODY_ERR_COMM;
	if (cfg && (cfg->flags & NBODY_FLAG_PARTITIONED))
		return let_create_distributed(cfg, local_particles, n_local, n_global, global_offset, rank, world, id, out);
	Sim* s = nullptr;
	int rc = create_for_comm(cfg, n_global, &s);
	if (rc) return rc;
	auto fail = [&](int code) { destroy_for_comm(s); return code; };
	s->comm = new Comm;
	Comm& cm = *s->comm;
	cm.rank = rank; cm.world = world;
	s->rank = rank;
	if (cudaMallocHost((void**) &cm.part_host, sizeof(uint32_t) * (world + 1)) != cudaSuccess ||
	    cudaMallocHost((void**) &cm.work_host, sizeof(float) * world) != cudaSuccess ||
	    cudaMalloc((void**) &cm.work_dev, sizeof(float) * world) != cudaSuccess ||
	    cudaMallocHost((void**) &cm.red_host, sizeof(float) * world) != cudaSuccess ||
	    cudaMalloc((void**) &cm.red_dev, sizeof(float) * world) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&cm.vel_stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreateWithFlags(&cm.ev_step, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&cm.ev_vel, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&cm.ev_pos, cudaEventDisableTiming) != cudaSuccess) { set_error("allocation of the partition tables failed"); return fail(NBODY_ERR_CUDA); }
	ncclUniqueId u;
	std::memcpy(&u, id, 128);
	ncclResult_t nr = g_nccl.CommInitRank(&cm.comm, world, u, rank);
	if (nr != ncclSuccess) { set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(nr)); cm.comm = nullptr; return fail(NBODY_ERR_COMM); }
	if ((cfg->flags & NBODY_FLAG_DIST_SORT) && g_nccl.CommSplit && !std::getenv("NBODY_NO_COMM_SPLIT")) {  // (the variable: A/B runs)
		// collective: every rank passes the same flags, so either all ranks split or none does
		if (g_nccl.CommSplit(cm.comm, 0, rank, &cm.vel_comm, nullptr) != ncclSuccess) cm.vel_comm = nullptr;
	}
	// assemble the global AoS array: every rank contributes its slice (all-gather with unequal counts)
	std::vector<unsigned long long> h(2 * world, 0ull);
	unsigned long long* d = nullptr;
	if (cudaMalloc((void**) &d, sizeof(unsigned long long) * 2 * world) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(NBODY_ERR_CUDA); }
	const unsigned long long mine[2] = {global_offset, n_local};
	cudaMemcpyAsync(d + 2 * rank, mine, sizeof(mine), cudaMemcpyHostToDevice, s->stream);
	nr = g_nccl.AllGather(d + 2 * rank, d, 2, ncclUint64, cm.comm, s->stream);
	cudaMemcpyAsync(h.data(), d, sizeof(unsigned long long) * 2 * world, cudaMemcpyDeviceToHost, s->stream);
	cudaStreamSynchronize(s->stream);
	cudaFree(d);
	if (nr != ncclSuccess) { set_error(std::string("ncclAllGather: ") + g_nccl.GetErrorString(nr)); return fail(NBODY_ERR_COMM); }
	unsigned long long covered = 0;
	for (int r = 0; r < world; ++r) covered += h[2 * r + 1];
	if (covered != n_global) { set_error("the ranks' slices do not add up to n_global"); return fail(NBODY_ERR_INVALID); }
	for (int r = 0; r < world; ++r) cm.part_host[r] = (uint32_t) h[2 * r];
	cm.part_host[world] = (uint32_t) n_global;
	for (int r = 0; r + 1 < world; ++r)
		if (h[2 * r] + h[2 * r + 1] != h[2 * (r + 1)]) { set_error("rank slices must be contiguous and ordered by rank"); return fail(NBODY_ERR_INVALID); }
	s->own_first = global_offset; s->own_count = n_local;
	if (cudaMemcpyAsync(s->aos_dev + global_offset, local_particles, n_local * sizeof(nbody_particle), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) {
		set_error("upload failed"); return fail(NBODY_ERR_CUDA);
	}
	g_nccl.GroupStart();
	for (int r = 0; r < world; ++r) {
		if (h[2 * r + 1] == 0) continue;
		nbody_particle* p = s->aos_dev + h[2 * r];
		g_nccl.Broadcast(p, p, h[2 * r + 1] * sizeof(nbody_particle), ncclChar, r, cm.comm, s->stream);
	}
	nr = g_nccl.GroupEnd();
	if (nr != ncclSuccess) { set_error(std::string("ncclBroadcast: ") + g_nccl.GetErrorString(nr)); return fail(NBODY_ERR_COMM); }
	launch_import(*s, s->aos_dev, n_global);
	cudaMemsetAsync(s->acc, 0, n_global * sizeof(float4), s->stream);
	if (cudaStreamSynchronize(s->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("import failed"); return fail(NBODY_ERR_CUDA); }
	*out = reinterpret_cast<nbody_cuda_sim*>(s);
	return NBODY_OK;
}

}  // extern "C"
