// C ABI of the solver (include/nbody_cuda.h): device memory arena, step orchestration,
// readback, parity exports. No CPU fallback: every entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace nbody {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
size_t sort_temp_bytes(uint64_t n);
int comm_step_exchange(Sim& s, float own_ms);       // comm.cu
float comm_work_imbalance(const Sim& s);            // comm.cu
void comm_destroy(Sim& s);                          // comm.cu
int comm_partition(Sim& s);                         // comm.cu
int comm_exchange_aos(Sim& s);                      // comm.cu
int comm_exchange_acc(Sim& s);                      // comm.cu
int comm_all_max(Sim& s, float* value);             // comm.cu
int comm_wait_velocities(Sim& s);                   // comm.cu
float next_time_step(const nbody_cuda_config& cfg, float acc_max);  // checkpoint.cu
void comm_adopt_partition(Sim& s);                  // comm.cu
void let_destroy(Sim& s);                           // let.cu
int let_step(Sim& s);                               // let.cu

namespace {

template <typename T>
int dev_alloc(Sim& s, T*& p, size_t count) {
	p = nullptr;
	if (count == 0) count = 1;
	NB_CUDA_CHECK(cudaMalloc((void**) &p, count * sizeof(T)));
	s.device_bytes += count * sizeof(T);
	return NBODY_OK;
}
template <typename T>
void dev_free(Sim& s, T*& p, size_t count) {
	if (p) { cudaFree(p); s.device_bytes -= std::max<size_t>(count, 1) * sizeof(T); p = nullptr; }
}

uint64_t clamp32(double v) { return (uint64_t) std::min(v, 4294967295.0); }

// Source-side arrays (geometry, child / count record, first particle: what the traversal reads of a node as a SOURCE) have room for
// `src_nodes` >= max_nodes entries: in partitioned mode the other ranks' trees are imported behind the own tree. Expansions and
// everything a node carries as a TARGET (list heads, parent, key) exist for the own tree only.
int alloc_nodes(Sim& s, uint32_t max_nodes, uint32_t src_nodes) {
	if (src_nodes < max_nodes) src_nodes = max_nodes;
	s.max_nodes = max_nodes;
	s.src_nodes = src_nodes;
	int rc;
	if ((rc = dev_alloc(s, s.geom, src_nodes))) return rc;
	if ((rc = dev_alloc(s, s.info, src_nodes))) return rc;
	if ((rc = dev_alloc(s, s.nbegin, src_nodes))) return rc;
	if ((rc = dev_alloc(s, s.M, (size_t) max_nodes * s.nc_stride))) return rc;
	if ((rc = dev_alloc(s, s.nparent, max_nodes))) return rc;
	if ((rc = dev_alloc(s, s.nkey, max_nodes))) return rc;
	if ((rc = dev_alloc(s, s.L, (size_t) max_nodes * s.nc_stride))) return rc;
	if ((rc = dev_alloc(s, s.near_ref, max_nodes))) return rc;
	if ((rc = dev_alloc(s, s.p2p_head, max_nodes))) return rc;
	if ((rc = dev_alloc(s, s.leaf_items, max_nodes))) return rc;
	return NBODY_OK;
}
void free_nodes(Sim& s) {
	const size_t m = s.max_nodes, sn = s.src_nodes;
	dev_free(s, s.geom, sn); dev_free(s, s.info, sn); dev_free(s, s.nbegin, sn); dev_free(s, s.M, m * s.nc_stride);
	dev_free(s, s.nparent, m); dev_free(s, s.nkey, m);
	dev_free(s, s.L, m * s.nc_stride); dev_free(s, s.near_ref, m); dev_free(s, s.p2p_head, m);
	dev_free(s, s.leaf_items, m);
}

struct PoolPlan { uint64_t near, p2p, m2l; uint32_t seg, gq, items; };

int alloc_pools(Sim& s, const PoolPlan& pl) {
	Pools& p = s.pools;
	int rc;
	p.near_cap = pl.near; p.p2p_cap = pl.p2p; p.m2l_cap = pl.m2l; p.seg_cap = pl.seg; p.gq_cap = pl.gq; p.items_cap = pl.items;
	for (int k = 0; k < 2; ++k) {
		if ((rc = dev_alloc(s, p.near[k], p.near_cap))) return rc;
		if ((rc = dev_alloc(s, p.gq[k], p.gq_cap))) return rc;
		if ((rc = dev_alloc(s, p.items[k], p.items_cap))) return rc;
	}
	if ((rc = dev_alloc(s, p.p2p, p.p2p_cap))) return rc;
	if ((rc = dev_alloc(s, p.m2l_id, p.m2l_cap))) return rc;
	if ((rc = dev_alloc(s, p.m2l_mask, p.m2l_cap))) return rc;
	if ((rc = dev_alloc(s, p.m2l_mask_lo, p.m2l_cap))) return rc;
	if ((rc = dev_alloc(s, p.seg, p.seg_cap))) return rc;
	return NBODY_OK;
}
void free_pools(Sim& s) {
	Pools& p = s.pools;
	for (int k = 0; k < 2; ++k) { dev_free(s, p.near[k], p.near_cap); dev_free(s, p.gq[k], p.gq_cap); dev_free(s, p.items[k], p.items_cap); }
	dev_free(s, p.p2p, p.p2p_cap); dev_free(s, p.m2l_id, p.m2l_cap); dev_free(s, p.m2l_mask, p.m2l_cap); dev_free(s, p.m2l_mask_lo, p.m2l_cap); dev_free(s, p.seg, p.seg_cap);
}
PoolPlan current_plan(const Sim& s) {
	const Pools& p = s.pools;
	return PoolPlan{p.near_cap, p.p2p_cap, p.m2l_cap, p.seg_cap, p.gq_cap, p.items_cap};
}


void free_all(Sim* s) {
	if (!s) return;
	cudaSetDevice(s->device);
	comm_destroy(*s);
	let_destroy(*s);
	for (int k = 0; k < 2; ++k) {
		dev_free(*s, s->posq[k], k == 1 ? s->src_cap : s->cap); dev_free(*s, s->velm[k], s->cap); dev_free(*s, s->orig[k], s->cap);
		dev_free(*s, s->keys[k], s->cap); dev_free(*s, s->idx[k], s->cap);
	}
	dev_free(*s, s->acc, s->cap);
	if (s->sort_tmp) { cudaFree(s->sort_tmp); s->sort_tmp = nullptr; }
	dev_free(*s, s->aos_dev, s->cap);
	dev_free(*s, s->scan_sums, (size_t) kScanBlocks + 1);
	free_nodes(*s);
	free_pools(*s);
	if (s->ctrl) cudaFree(s->ctrl);
	if (s->ctrl_host) cudaFreeHost(s->ctrl_host);
	for (auto& e : s->ev) if (e) cudaEventDestroy(e);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
}

int validate(const nbody_cuda_config* cfg, uint64_t n) {
	if (!cfg) { set_error("config is NULL"); return NBODY_ERR_INVALID; }
	if (cfg->abi_version != NBODY_CUDA_ABI_VERSION) { set_error("abi_version mismatch"); return NBODY_ERR_INVALID; }
	if (n == 0 || n > 0xfffffff0ull) { set_error("particle count must be in [1, 2^32-16)"); return NBODY_ERR_INVALID; }
	if (!(cfg->bounds[0] > 0 && cfg->bounds[1] > 0 && cfg->bounds[2] > 0)) { set_error("bounds must be positive"); return NBODY_ERR_INVALID; }
	if (cfg->order < 2 || cfg->order > 5) { set_error("order must be 2, 3, 4 or 5"); return NBODY_ERR_INVALID; }
	if (cfg->max_depth < 1 || cfg->max_depth > (uint32_t) kMaxDepth) { set_error("max_depth must be in [1,21]"); return NBODY_ERR_INVALID; }
	if (cfg->leaf_capacity < 1) { set_error("leaf_capacity must be >= 1"); return NBODY_ERR_INVALID; }
	if (!(cfg->softening >= 0) || !(cfg->mac_ratio > 0)) { set_error("softening must be >= 0 and mac_ratio > 0"); return NBODY_ERR_INVALID; }
	if (cfg->softening > 0 && !(cfg->softening * cfg->softening >= 1.17549435e-38f)) {
		// the softened kernels use rsqrt.approx.ftz: a denormal eps^2 would flush to zero and turn every particle's self term into 0 * inf
		set_error("softening^2 is denormal: use 0 (unsoftened kernels) or a softening >= 1.1e-19");
		return NBODY_ERR_INVALID;
	}
	if (cfg->integrator > 1) { set_error("unknown integrator"); return NBODY_ERR_INVALID; }
	if (!(cfg->low_order_tau >= 0)) { set_error("low_order_tau must be >= 0"); return NBODY_ERR_INVALID; }
	if (!(cfg->time_step_eta >= 0) || !(cfg->time_step_min >= 0) || !(cfg->time_step_max >= 0) ||
	    (cfg->time_step_max > 0 && cfg->time_step_min > cfg->time_step_max)) {
		set_error("time_step_eta/min/max must be >= 0 and min <= max");
		return NBODY_ERR_INVALID;
	}
	if (cfg->time_step_eta > 0 && !(cfg->time_step > 0)) { set_error("the variable time step needs time_step > 0"); return NBODY_ERR_INVALID; }
	return NBODY_OK;
}

// One attempt at a step: enqueue every stage, then one synchronisation to read the control block.
// `retry`: a repeat of the same step after a pool overflow. With the distributed sort the first attempt's sorted keys, permutation
// and gathered positions are kept (nothing has overwritten them: the leaf kernel refuses to run after an overflow), because the
// sort contains a collective and a rank retries alone — its peers are already waiting in the end-of-step exchange.
int run_pipeline(Sim& s, bool retry) {
	cudaStream_t st = s.stream;
	const bool direct = (s.cfg.flags & NBODY_FLAG_DIRECT) != 0;
	const bool keep_sort = retry && s.comm && (s.cfg.flags & NBODY_FLAG_DIST_SORT) && !(s.cfg.flags & NBODY_FLAG_CUB_SORT);
	NvtxRange step_range("nbody step");
	NB_CUDA_CHECK(cudaEventRecord(s.ev[0], st));
	int rc = NBODY_OK;
	{ NvtxRange r("keys + sort + gather"); rc = keep_sort ? NBODY_OK : launch_keys_sort_permute(s); }
	if (rc) return rc;
	NB_CUDA_CHECK(cudaEventRecord(s.ev[1], st));
	{ NvtxRange r("octree build"); launch_tree_build(s); }
	NB_CUDA_CHECK(cudaEventRecord(s.ev[2], st));
	if (s.comm && (rc = comm_partition(s))) return rc;
	if (!direct) {
		{ NvtxRange r("P2M + M2M"); launch_upsweep(s); }
		NB_CUDA_CHECK(cudaEventRecord(s.ev[3], st));
		{ NvtxRange r("dual-tree traversal"); launch_traversal(s); }
		NB_CUDA_CHECK(cudaEventRecord(s.ev[4], st));
		{ NvtxRange r("M2L"); launch_m2l(s); }
		NB_CUDA_CHECK(cudaEventRecord(s.ev[5], st));
		{ NvtxRange r("L2L"); launch_l2l(s); }
		if (s.comm) { if ((rc = comm_wait_velocities(s))) return rc; launch_gather_velocities(s); }
		NB_CUDA_CHECK(cudaEventRecord(s.ev[6], st));
		{ NvtxRange r("P2P + L2P + integrator"); launch_leaf(s); }
	} else {
		for (int k = 3; k <= 5; ++k) NB_CUDA_CHECK(cudaEventRecord(s.ev[k], st));
		if (s.comm) { if ((rc = comm_wait_velocities(s))) return rc; launch_gather_velocities(s); }
		NB_CUDA_CHECK(cudaEventRecord(s.ev[6], st));
		launch_direct(s);
	}
	if (s.cfg.time_step_eta > 0.0f) launch_acc_max(s);
	NB_CUDA_CHECK(cudaEventRecord(s.ev[7], st));
	NB_CUDA_CHECK(cudaMemcpyAsync(s.ctrl_host, s.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
	NB_CUDA_CHECK(cudaStreamSynchronize(st));  // the one synchronisation of a step
	NB_CUDA_CHECK(cudaGetLastError());
	if (s.comm && s.ctrl_host->status == 0) {
		// distributed: the slice boundaries were computed on the device; now that the host has them, exchange the slices
		comm_adopt_partition(s);
		float own_ms = 0.0f;  // traversal .. leaf kernel: the stages this rank runs for its own slice only
		NB_CUDA_CHECK(cudaEventElapsedTime(&own_ms, s.ev[3], s.ev[7]));
		if ((rc = comm_step_exchange(s, own_ms))) return rc;
		NB_CUDA_CHECK(cudaEventRecord(s.ev[8], st));
		NB_CUDA_CHECK(cudaStreamSynchronize(st));
	} else {
		NB_CUDA_CHECK(cudaEventRecord(s.ev[8], st));
		NB_CUDA_CHECK(cudaEventSynchronize(s.ev[8]));
	}
	return NBODY_OK;
}

int grow_after_overflow(Sim& s, uint32_t status) {
	const double g = 1.6;
	if (status & kOvfNodes) {
		const uint32_t want = (uint32_t) clamp32((double) s.max_nodes * g + 1024);
		const uint32_t extra = s.src_nodes - s.max_nodes;  // partitioned mode: keep the room for the imported trees
		free_nodes(s);
		int rc = alloc_nodes(s, want, (uint32_t) clamp32((double) want + extra));
		if (rc) return rc;
	}
	if (status & kOvfDepth) s.depth_bound = s.trav_bound = kMaxDepth;  // the tree outgrew last step's depth + 1: unbounded level loops
	if (status & ~(kOvfNodes | kOvfDepth)) {
		PoolPlan pl = current_plan(s);
		if (status & kOvfNear) pl.near = clamp32((double) pl.near * g);
		if (status & kOvfP2P) pl.p2p = clamp32((double) pl.p2p * g);
		if (status & kOvfM2L) pl.m2l = clamp32((double) pl.m2l * g);
		if (status & kOvfSeg) pl.seg = (uint32_t) clamp32((double) pl.seg * g);
		if (status & kOvfGroups) pl.gq = (uint32_t) clamp32((double) pl.gq * g);
		if (status & kOvfItems) pl.items = (uint32_t) clamp32((double) pl.items * g);
		const PoolPlan old = current_plan(s);
		if (pl.near == old.near && pl.p2p == old.p2p && pl.m2l == old.m2l && pl.seg == old.seg && pl.gq == old.gq && pl.items == old.items) {
			set_error("interaction-list pool hit the 2^32-entry limit");
			return NBODY_ERR_CAPACITY;
		}
		free_pools(s);
		int rc = alloc_pools(s, pl);
		if (rc) return rc;
	}
	return NBODY_OK;
}

// `cap` >= n: allocated length of the per-particle arrays; `halo`: extra room behind posq[1] (both only differ from n / 0 in
// partitioned mode, where a rank's particle count changes from step to step and the sources include other ranks' particles).
int create_common(const nbody_cuda_config* cfg, uint64_t n, Sim** out, uint64_t cap = 0, uint64_t halo = 0) {
	if (cap < n) cap = n;
	int rc = validate(cfg, n);
	if (rc) return rc;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		set_error("no CUDA device: this library has no CPU fallback");
		return NBODY_ERR_CUDA;
	}
	Sim* s = new Sim;
	s->cfg = *cfg;
	s->dt = cfg->time_step;
	if (s->cfg.pool_scale <= 0) s->cfg.pool_scale = 1.0f;
	if (cfg->device >= 0) s->device = cfg->device; else cudaGetDevice(&s->device);
	auto fail = [&](int code) { free_all(s); return code; };
	if (cudaSetDevice(s->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(NBODY_ERR_CUDA); }
	cudaDeviceProp prop{};
	cudaGetDeviceProperties(&prop, s->device);
	if (prop.major < 10) { set_error(std::string("device '") + prop.name + "' is not sm_100: this library is built for B200 only"); return fail(NBODY_ERR_CUDA); }
	s->n = n; s->cap = cap; s->src_cap = cap + halo; s->n_global = n;
	s->own_first = 0; s->own_count = n;
	s->nc_stride = coef_stride((int) s->cfg.order);
	if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(NBODY_ERR_CUDA); }
	for (auto& e : s->ev) if (cudaEventCreate(&e) != cudaSuccess) { set_error("event creation failed"); return fail(NBODY_ERR_CUDA); }
	for (int k = 0; k < 2; ++k) {
		if ((rc = dev_alloc(*s, s->posq[k], k == 1 ? s->src_cap : cap)) || (rc = dev_alloc(*s, s->velm[k], cap)) || (rc = dev_alloc(*s, s->orig[k], cap)) ||
		    (rc = dev_alloc(*s, s->keys[k], cap)) || (rc = dev_alloc(*s, s->idx[k], cap)))
			return fail(rc);
	}
	if ((rc = dev_alloc(*s, s->acc, cap)) || (rc = dev_alloc(*s, s->aos_dev, cap)) || (rc = dev_alloc(*s, s->scan_sums, (size_t) kScanBlocks + 1)))
		return fail(rc);
	s->sort_tmp_bytes = sort_temp_bytes(cap);
	if (cudaMalloc(&s->sort_tmp, std::max<size_t>(s->sort_tmp_bytes, 16)) != cudaSuccess) { set_error("sort scratch allocation failed"); return fail(NBODY_ERR_CUDA); }
	s->device_bytes += s->sort_tmp_bytes;
	if (cudaMalloc((void**) &s->ctrl, sizeof(Ctrl)) != cudaSuccess || cudaMallocHost((void**) &s->ctrl_host, sizeof(Ctrl)) != cudaSuccess) {
		set_error("control block allocation failed");
		return fail(NBODY_ERR_CUDA);
	}
	std::memset(s->ctrl_host, 0, sizeof(Ctrl));
	const double sc = s->cfg.pool_scale, dn = (double) n;  // (partitioned mode: n = the rank's initial share; the pools grow on demand)
	// node budget: ~0.35-0.6 nodes per particle at capacity 8 (measured with the oracle); fewer for larger leaves
	const double per_particle = std::min(1.25, 10.0 / (double) s->cfg.leaf_capacity);
	const uint32_t nodes0 = (uint32_t) clamp32(sc * (per_particle * dn + 65536));
	if ((rc = alloc_nodes(*s, nodes0, nodes0))) return fail(rc);
	PoolPlan pl;
	pl.near = clamp32(sc * (128.0 * dn * std::min(1.0, 16.0 / s->cfg.leaf_capacity) + 1048576));
	pl.p2p = clamp32(sc * (96.0 * dn * std::min(1.0, 16.0 / s->cfg.leaf_capacity) + 1048576));
	pl.m2l = clamp32(sc * (128.0 * dn * std::min(1.0, 16.0 / s->cfg.leaf_capacity) + 1048576));
	pl.seg = (uint32_t) clamp32(2.0 * s->max_nodes);
	pl.gq = s->max_nodes;
	pl.items = s->max_nodes;
	if ((rc = alloc_pools(*s, pl))) return fail(rc);
	*out = s;
	return NBODY_OK;
}

int upload(Sim& s, const nbody_particle* particles, uint64_t n) {
	NB_CUDA_CHECK(cudaMemcpyAsync(s.aos_dev, particles, n * sizeof(nbody_particle), cudaMemcpyHostToDevice, s.stream));
	launch_import(s, s.aos_dev, n);
	NB_CUDA_CHECK(cudaMemsetAsync(s.acc, 0, n * sizeof(float4), s.stream));
	NB_CUDA_CHECK(cudaStreamSynchronize(s.stream));
	NB_CUDA_CHECK(cudaGetLastError());
	s.lists_valid = false;
	return NBODY_OK;
}

// ---- host-side view of the last tree, for the parity exports --------------------
struct HostTree {
	uint32_t n_nodes = 0;
	std::vector<float4> geom;
	std::vector<uint2> info;
	std::vector<uint32_t> nbegin, nparent, dfs_of, lm_of, depth;
	std::vector<uint64_t> nkey;
};

int fetch_tree(Sim& s, HostTree& t) {
	if (s.steps_done == 0) { set_error("no step has been taken yet"); return NBODY_ERR_STATE; }
	const Ctrl& c = *s.ctrl_host;
	const uint32_t m = t.n_nodes = c.n_nodes;
	t.geom.resize(m); t.info.resize(m); t.nbegin.resize(m); t.nparent.resize(m); t.nkey.resize(m);
	NB_CUDA_CHECK(cudaMemcpy(t.geom.data(), s.geom, m * sizeof(float4), cudaMemcpyDeviceToHost));
	NB_CUDA_CHECK(cudaMemcpy(t.info.data(), s.info, m * sizeof(uint2), cudaMemcpyDeviceToHost));
	NB_CUDA_CHECK(cudaMemcpy(t.nbegin.data(), s.nbegin, m * 4, cudaMemcpyDeviceToHost));
	NB_CUDA_CHECK(cudaMemcpy(t.nparent.data(), s.nparent, m * 4, cudaMemcpyDeviceToHost));
	NB_CUDA_CHECK(cudaMemcpy(t.nkey.data(), s.nkey, m * 8, cudaMemcpyDeviceToHost));
	t.depth.resize(m);
	for (int l = 0; l < kNumLevels; ++l)
		for (uint32_t i = c.level_off[l]; i < c.level_off[l + 1] && i < m; ++i) t.depth[i] = (uint32_t) l;
	// level-major -> DFS pre-order (children in octant order)
	t.dfs_of.assign(m, 0); t.lm_of.assign(m, 0);
	std::vector<uint32_t> stack;
	stack.push_back(0);
	uint32_t next = 0;
	while (!stack.empty()) {
		const uint32_t id = stack.back(); stack.pop_back();
		t.dfs_of[id] = next; t.lm_of[next] = id; ++next;
		if (t.info[id].x) for (int k = 7; k >= 0; --k) stack.push_back(t.info[id].x + (uint32_t) k);
	}
	if (next != m) { set_error("tree export: node count mismatch"); return NBODY_ERR_STATE; }
	return NBODY_OK;
}

}  // namespace
}  // namespace nbody

using namespace nbody;

extern "C" {

void nbody_cuda_default_config(nbody_cuda_config* cfg) {
	if (!cfg) return;
	std::memset(cfg, 0, sizeof(*cfg));
	cfg->abi_version = NBODY_CUDA_ABI_VERSION;
	cfg->bounds[0] = cfg->bounds[1] = cfg->bounds[2] = 1.0f; cfg->bounds[3] = 0.0f;
	cfg->time_step = 0.001f;
	cfg->force_constant = 1.0f;
	cfg->softening = 0.01f;
	cfg->mac_ratio = 0.5f;
	cfg->leaf_capacity = 8;
	cfg->max_depth = 21;
	cfg->order = 4;
	cfg->integrator = NBODY_KICK_DRIFT;
	cfg->flags = 0;
	cfg->device = -1;
	cfg->pool_scale = 1.0f;
	cfg->low_order_tau = 0.13f;
}

void nbody_cuda_tuned_config(nbody_cuda_config* cfg) {
	if (!cfg) return;
	nbody_cuda_default_config(cfg);
	cfg->leaf_capacity = 48;
}

int nbody_cuda_create(const nbody_cuda_config* cfg, const nbody_particle* particles, uint64_t n, nbody_cuda_sim** out) {
	if (!out || !particles) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	*out = nullptr;
	Sim* s = nullptr;
	int rc = create_common(cfg, n, &s);
	if (rc) return rc;
	if ((rc = upload(*s, particles, n))) { free_all(s); return rc; }
	*out = reinterpret_cast<nbody_cuda_sim*>(s);
	return NBODY_OK;
}

void nbody_cuda_destroy(nbody_cuda_sim* sim) { free_all(reinterpret_cast<Sim*>(sim)); }

int nbody_cuda_set_particles(nbody_cuda_sim* sim, const nbody_particle* particles, uint64_t n) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !particles) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (n != s->n) { set_error("set_particles: particle count differs from the simulation's"); return NBODY_ERR_INVALID; }
	if (s->comm || s->let) { set_error("set_particles is not available on a distributed simulation"); return NBODY_ERR_STATE; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	return upload(*s, particles, n);
}

int nbody_cuda_step(nbody_cuda_sim* sim, float* time_out) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s) { set_error("NULL simulation"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	if (s->let) {  // partitioned mode: own particles + locally essential tree (let.cu)
		const int rc = let_step(*s);
		if (rc == NBODY_OK && time_out) *time_out = s->time;
		return rc;
	}
	s->stats.retries = 0;
	for (int attempt = 0;; ++attempt) {
		int rc = run_pipeline(*s, attempt > 0);
		if (rc) return rc;
		const uint32_t status = s->ctrl_host->status;
		if (status == 0) break;
		if (attempt >= 24) { set_error("step: pools still overflow after 24 growth attempts"); return NBODY_ERR_CAPACITY; }
		if ((rc = grow_after_overflow(*s, status))) return rc;
		++s->stats.retries;
	}
	std::swap(s->orig[0], s->orig[1]);
	s->time += s->dt;  // FP32 accumulation like src/open_cl_simulation.cpp:103
	++s->steps_done;
	// next step's level loops: one level deeper than this step's tree at most (kOvfDepth re-runs unbounded)
	s->depth_bound = s->trav_bound = std::min<int>((int) s->cfg.max_depth, (int) s->ctrl_host->n_levels);
	s->dt_last = s->dt;
	if (s->cfg.time_step_eta > 0.0f) {
		// variable time step: this step's largest acceleration sets the next step (rule in nbody_cuda_next_time_step)
		float a2;
		std::memcpy(&a2, &s->ctrl_host->acc_max2_bits, sizeof(float));
		if (s->comm) { int rc = comm_all_max(*s, &a2); if (rc) return rc; }
		s->acc_max = std::sqrt(a2);
		s->dt = next_time_step(s->cfg, s->acc_max);
	} else {
		s->dt = s->dt_last;  // fixed step: cfg.time_step, or what nbody_cuda_set_time_step installed
	}
	s->lists_valid = !(s->cfg.flags & NBODY_FLAG_DIRECT);
	// statistics
	const Ctrl& c = *s->ctrl_host;
	nbody_cuda_stats& t = s->stats;
	t.n_particles = s->n; t.n_nodes = c.n_nodes; t.n_levels = c.n_levels; t.n_leaves = c.stat_leaves;
	t.m2l_entries = c.m2l_cursor; t.m2l_interactions = c.stat_m2l_inter; t.m2l_interactions_low = c.stat_m2l_low; t.p2p_entries = c.stat_p2p_entries;
	t.p2p_interactions = c.stat_p2p_inter >= s->own_count ? c.stat_p2p_inter - s->own_count : 0;  // drop the i == j terms
	t.near_entries = c.stat_near; t.device_bytes = s->device_bytes;
	auto ms = [&](int a, int b) { float v = 0; cudaEventElapsedTime(&v, s->ev[a], s->ev[b]); return v; };
	t.ms_sort = ms(0, 1); t.ms_tree = ms(1, 2); t.ms_upsweep = ms(2, 3); t.ms_traverse = ms(3, 4); t.ms_m2l = ms(4, 5); t.ms_l2l = ms(5, 6);
	t.ms_leaf = ms(6, 7); t.ms_comm = ms(7, 8); t.ms_total = ms(0, 8);
	t.work_imbalance = comm_work_imbalance(*s);
	if (s->cfg.flags & NBODY_FLAG_DIRECT) t.p2p_interactions = s->n * (s->n - 1);
	if (time_out) *time_out = s->time;
	return NBODY_OK;
}

uint64_t nbody_cuda_num_particles(const nbody_cuda_sim* sim) { return sim ? reinterpret_cast<const Sim*>(sim)->n : 0; }

int nbody_cuda_get_particles(nbody_cuda_sim* sim, nbody_particle* out, uint64_t capacity) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !out) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (capacity < s->n) { set_error("get_particles: output buffer too small"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	if (s->comm) { int rc = comm_wait_velocities(*s); if (rc) return rc; }
	launch_export(*s, s->aos_dev, s->n);
	NB_CUDA_CHECK(cudaMemcpyAsync(out, s->aos_dev, s->n * sizeof(nbody_particle), cudaMemcpyDeviceToHost, s->stream));
	NB_CUDA_CHECK(cudaStreamSynchronize(s->stream));
	return NBODY_OK;
}

int nbody_cuda_get_permutation(nbody_cuda_sim* sim, uint32_t* orig_index, uint64_t capacity) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !orig_index || capacity < s->n) { set_error("get_permutation: bad argument"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	NB_CUDA_CHECK(cudaMemcpy(orig_index, s->orig[0], s->n * 4, cudaMemcpyDeviceToHost));
	return NBODY_OK;
}

int nbody_cuda_get_accelerations(nbody_cuda_sim* sim, float* xyz, uint64_t capacity) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !xyz || capacity < s->n) { set_error("get_accelerations: bad argument"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	if (s->comm) {  // collective in distributed mode: the slices are exchanged on first use
		int rc = comm_exchange_acc(*s);
		if (rc) return rc;
		NB_CUDA_CHECK(cudaStreamSynchronize(s->stream));
	}
	std::vector<float4> h(s->n);
	NB_CUDA_CHECK(cudaMemcpy(h.data(), s->acc, s->n * sizeof(float4), cudaMemcpyDeviceToHost));
	for (uint64_t i = 0; i < s->n; ++i) { xyz[3 * i] = h[i].x; xyz[3 * i + 1] = h[i].y; xyz[3 * i + 2] = h[i].z; }
	return NBODY_OK;
}

int nbody_cuda_get_keys(nbody_cuda_sim* sim, uint64_t* keys, uint64_t capacity) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !keys || capacity < s->n) { set_error("get_keys: bad argument"); return NBODY_ERR_INVALID; }
	if (s->steps_done == 0) { set_error("no step has been taken yet"); return NBODY_ERR_STATE; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	NB_CUDA_CHECK(cudaMemcpy(keys, s->keys[0], s->n * 8, cudaMemcpyDeviceToHost));
	return NBODY_OK;
}

int nbody_cuda_get_tree(nbody_cuda_sim* sim, uint32_t* n_nodes, uint32_t capacity, uint32_t* depth, uint64_t* prefix,
                        uint32_t* leaf_index, uint32_t* leaf_count, uint8_t* has_children, uint32_t* child_off9,
                        int32_t* parent_off, uint32_t* sibling, float* geom4) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !n_nodes) { set_error("get_tree: bad argument"); return NBODY_ERR_INVALID; }
	if (s->steps_done == 0) { set_error("no step has been taken yet"); return NBODY_ERR_STATE; }
	*n_nodes = s->ctrl_host->n_nodes;
	if (!depth && !prefix && !leaf_index && !leaf_count && !has_children && !child_off9 && !parent_off && !sibling && !geom4) return NBODY_OK;
	if (capacity < *n_nodes) { set_error("get_tree: capacity too small"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	HostTree t;
	int rc = fetch_tree(*s, t);
	if (rc) return rc;
	const uint32_t m = t.n_nodes;
	// subtree sizes, by sweeping DFS ids backwards
	std::vector<uint32_t> size(m, 1);
	for (uint32_t d = m; d-- > 0;) {
		const uint32_t id = t.lm_of[d];
		if (t.info[id].x) { uint32_t sz = 1; for (uint32_t k = 0; k < 8; ++k) sz += size[t.dfs_of[t.info[id].x + k]]; size[d] = sz; }
	}
	for (uint32_t d = 0; d < m; ++d) {
		const uint32_t id = t.lm_of[d];
		if (depth) depth[d] = t.depth[id];
		if (prefix) prefix[d] = t.nkey[id];
		if (leaf_index) leaf_index[d] = t.nbegin[id];
		if (leaf_count) leaf_count[d] = t.info[id].y;
		if (has_children) has_children[d] = t.info[id].x ? 1 : 0;
		if (child_off9) {
			for (uint32_t k = 0; k < 8; ++k) child_off9[9 * (size_t) d + k] = t.info[id].x ? t.dfs_of[t.info[id].x + k] - d : 0u;
			child_off9[9 * (size_t) d + 8] = size[d];
		}
		if (parent_off) parent_off[d] = id == 0 ? 0 : (int32_t) t.dfs_of[t.nparent[id]] - (int32_t) d;
		if (sibling) sibling[d] = id == 0 ? 0u : (uint32_t) ((t.nkey[id] >> (3 * (kMaxDepth - t.depth[id]))) & 7u);
		if (geom4) { geom4[4 * (size_t) d] = t.geom[id].x; geom4[4 * (size_t) d + 1] = t.geom[id].y; geom4[4 * (size_t) d + 2] = t.geom[id].z; geom4[4 * (size_t) d + 3] = t.geom[id].w; }
	}
	return NBODY_OK;
}

int nbody_cuda_get_lists(nbody_cuda_sim* sim, uint64_t* n_m2l, uint32_t* m2l_pairs, uint64_t* n_p2p, uint32_t* p2p_pairs) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !n_m2l || !n_p2p) { set_error("get_lists: bad argument"); return NBODY_ERR_INVALID; }
	if (!s->lists_valid) { set_error("get_lists: no interaction lists (take an FMM step first)"); return NBODY_ERR_STATE; }
	const Ctrl& c = *s->ctrl_host;
	*n_m2l = c.stat_m2l_inter;
	*n_p2p = c.stat_p2p_entries;
	if (!m2l_pairs && !p2p_pairs) return NBODY_OK;
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	HostTree t;
	int rc = fetch_tree(*s, t);
	if (rc) return rc;
	if (m2l_pairs) {
		std::vector<uint32_t> ids(c.m2l_cursor);
		std::vector<uint8_t> masks(c.m2l_cursor);
		NB_CUDA_CHECK(cudaMemcpy(ids.data(), s->pools.m2l_id, ids.size() * 4, cudaMemcpyDeviceToHost));
		NB_CUDA_CHECK(cudaMemcpy(masks.data(), s->pools.m2l_mask, masks.size(), cudaMemcpyDeviceToHost));
		uint64_t w = 0;
		for (int which = 0; which < 2; ++which) {
			std::vector<Group> items(c.items_count[which]);
			NB_CUDA_CHECK(cudaMemcpy(items.data(), s->pools.items[which], items.size() * sizeof(Group), cudaMemcpyDeviceToHost));
			for (const Group& g : items)
				for (uint32_t e = 0; e < g.list_cnt; ++e)
					for (uint32_t b = 0; b < g.nt; ++b)
						if (masks[g.list_off + e] >> b & 1u) {
							m2l_pairs[2 * w] = t.dfs_of[g.first + b];
							m2l_pairs[2 * w + 1] = t.dfs_of[ids[g.list_off + e]];
							++w;
						}
		}
		if (w != *n_m2l) { set_error("get_lists: M2L count mismatch"); return NBODY_ERR_STATE; }
	}
	if (p2p_pairs) {
		std::vector<uint2> src(c.p2p_cursor);
		std::vector<uint32_t> head(t.n_nodes);
		// entries carry {first particle, count}: the source leaf is the non-empty childless node that starts there
		std::vector<uint32_t> leaf_at(s->n, 0xffffffffu);
		for (uint32_t id = 0; id < t.n_nodes; ++id)
			if (t.info[id].x == 0 && t.info[id].y > 0) leaf_at[t.nbegin[id]] = id;
		std::vector<Segment> segs(c.seg_cursor);
		NB_CUDA_CHECK(cudaMemcpy(src.data(), s->pools.p2p, src.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
		NB_CUDA_CHECK(cudaMemcpy(segs.data(), s->pools.seg, segs.size() * sizeof(Segment), cudaMemcpyDeviceToHost));
		NB_CUDA_CHECK(cudaMemcpy(head.data(), s->p2p_head, head.size() * 4, cudaMemcpyDeviceToHost));
		uint64_t w = 0;
		for (uint32_t id = 0; id < t.n_nodes; ++id)
			for (uint32_t si = head[id]; si != 0xffffffffu; si = segs[si].next)
				for (uint32_t e = 0; e < segs[si].cnt; ++e) {
					p2p_pairs[2 * w] = t.dfs_of[id];
					p2p_pairs[2 * w + 1] = t.dfs_of[leaf_at[src[segs[si].off + e].x]];
					++w;
				}
		if (w != *n_p2p) { set_error("get_lists: P2P count mismatch"); return NBODY_ERR_STATE; }
	}
	return NBODY_OK;
}

int nbody_cuda_get_expansions(nbody_cuda_sim* sim, float* multipoles, float* locals, uint64_t capacity_floats) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s) { set_error("NULL simulation"); return NBODY_ERR_INVALID; }
	if (!s->lists_valid) { set_error("get_expansions: take an FMM step first"); return NBODY_ERR_STATE; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	HostTree t;
	int rc = fetch_tree(*s, t);
	if (rc) return rc;
	const int nc = ncoef((int) s->cfg.order), st = s->nc_stride;
	if (capacity_floats < (uint64_t) t.n_nodes * nc) { set_error("get_expansions: capacity too small"); return NBODY_ERR_INVALID; }
	std::vector<float> h((size_t) t.n_nodes * st);
	for (int which = 0; which < 2; ++which) {
		float* dst = which == 0 ? multipoles : locals;
		if (!dst) continue;
		NB_CUDA_CHECK(cudaMemcpy(h.data(), which == 0 ? s->M : s->L, h.size() * 4, cudaMemcpyDeviceToHost));
		for (uint32_t d = 0; d < t.n_nodes; ++d)
			for (int a = 0; a < nc; ++a) dst[(size_t) d * nc + a] = h[(size_t) t.lm_of[d] * st + a];
	}
	return NBODY_OK;
}

int nbody_cuda_get_stats(nbody_cuda_sim* sim, nbody_cuda_stats* stats) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !stats) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	s->stats.device_bytes = s->device_bytes;
	*stats = s->stats;
	return NBODY_OK;
}

int nbody_cuda_direct_field(int device, const float* src_posq, uint64_t n_src, const float* tgt_pos4, uint64_t n_tgt,
                            float softening, float* field_xyz, float* ms_out, uint32_t repeats) {
	if (!src_posq || !tgt_pos4 || !field_xyz || n_src == 0 || n_tgt == 0) { set_error("direct_field: bad argument"); return NBODY_ERR_INVALID; }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: this library has no CPU fallback"); return NBODY_ERR_CUDA; }
	if (device >= 0) NB_CUDA_CHECK(cudaSetDevice(device));
	float4 *src = nullptr, *tgt = nullptr, *out = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	int rc = NBODY_OK;
	auto cleanup = [&]() { cudaFree(src); cudaFree(tgt); cudaFree(out); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
#define DF_CHECK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_error(std::string(#x) + ": " + cudaGetErrorString(_e)); cleanup(); return NBODY_ERR_CUDA; } } while (0)
	DF_CHECK(cudaMalloc((void**) &src, n_src * 16));
	DF_CHECK(cudaMalloc((void**) &tgt, n_tgt * 16));
	DF_CHECK(cudaMalloc((void**) &out, n_tgt * 16));
	DF_CHECK(cudaMemcpy(src, src_posq, n_src * 16, cudaMemcpyHostToDevice));
	DF_CHECK(cudaMemcpy(tgt, tgt_pos4, n_tgt * 16, cudaMemcpyHostToDevice));
	DF_CHECK(cudaEventCreate(&e0));
	DF_CHECK(cudaEventCreate(&e1));
	if (repeats == 0) repeats = 1;
	direct_field_device(src, n_src, tgt, n_tgt, softening * softening, out, 0);  // warm-up
	DF_CHECK(cudaEventRecord(e0, 0));
	for (uint32_t r = 0; r < repeats; ++r) direct_field_device(src, n_src, tgt, n_tgt, softening * softening, out, 0);
	DF_CHECK(cudaEventRecord(e1, 0));
	DF_CHECK(cudaEventSynchronize(e1));
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	if (ms_out) *ms_out = ms / repeats;
	std::vector<float4> h(n_tgt);
	DF_CHECK(cudaMemcpy(h.data(), out, n_tgt * 16, cudaMemcpyDeviceToHost));
	for (uint64_t i = 0; i < n_tgt; ++i) { field_xyz[3 * i] = h[i].x; field_xyz[3 * i + 1] = h[i].y; field_xyz[3 * i + 2] = h[i].z; }
#undef DF_CHECK
	cleanup();
	return rc;
}

int nbody_cuda_get_owned_particles(nbody_cuda_sim* sim, nbody_particle* out, uint64_t capacity) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !out) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (capacity < s->own_count) { set_error("get_owned_particles: output buffer too small"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	if (s->comm) { int rc = comm_wait_velocities(*s); if (rc) return rc; }
	launch_export(*s, s->aos_dev, s->n);
	const uint64_t first = s->let ? 0 : s->own_first;  // partitioned mode: the arrays hold the owned particles only
	NB_CUDA_CHECK(cudaMemcpyAsync(out, s->aos_dev + first, s->own_count * sizeof(nbody_particle), cudaMemcpyDeviceToHost, s->stream));
	NB_CUDA_CHECK(cudaStreamSynchronize(s->stream));
	return NBODY_OK;
}

int nbody_cuda_set_owned_particles(nbody_cuda_sim* sim, const nbody_particle* particles, uint64_t n) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !particles) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (n != s->own_count) { set_error("set_owned_particles: count differs from the owned range"); return NBODY_ERR_INVALID; }
	NB_CUDA_CHECK(cudaSetDevice(s->device));
	if (s->let) {  // partitioned mode: nothing to exchange, the rank holds exactly its own particles; identities are kept
		NB_CUDA_CHECK(cudaMemcpyAsync(s->aos_dev, particles, n * sizeof(nbody_particle), cudaMemcpyHostToDevice, s->stream));
		launch_import_ids(*s, s->aos_dev, n, 0u, true);
		NB_CUDA_CHECK(cudaStreamSynchronize(s->stream));
		NB_CUDA_CHECK(cudaGetLastError());
		return NBODY_OK;
	}
	if (s->comm) { int rc = comm_wait_velocities(*s); if (rc) return rc; }  // the import below overwrites the velocity plane
	NB_CUDA_CHECK(cudaMemcpyAsync(s->aos_dev + s->own_first, particles, n * sizeof(nbody_particle), cudaMemcpyHostToDevice, s->stream));
	if (s->comm) { int rc = comm_exchange_aos(*s); if (rc) return rc; }
	// the permutation is kept: the caller hands back the particles it read with get_owned_particles
	{
		uint32_t* keep = s->orig[1];
		NB_CUDA_CHECK(cudaMemcpyAsync(keep, s->orig[0], s->n * 4, cudaMemcpyDeviceToDevice, s->stream));
		launch_import(*s, s->aos_dev, s->n);
		NB_CUDA_CHECK(cudaMemcpyAsync(s->orig[0], keep, s->n * 4, cudaMemcpyDeviceToDevice, s->stream));
	}
	NB_CUDA_CHECK(cudaStreamSynchronize(s->stream));
	NB_CUDA_CHECK(cudaGetLastError());
	return NBODY_OK;
}

int nbody_cuda_owned_range(nbody_cuda_sim* sim, uint64_t* first, uint64_t* count) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !first || !count) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	*first = s->own_first; *count = s->own_count;
	return NBODY_OK;
}

const char* nbody_cuda_last_error(void) { return g_error.c_str(); }

}  // extern "C"

// used by comm.cu
namespace nbody {
int create_for_comm(const nbody_cuda_config* cfg, uint64_t n, Sim** out) { return create_common(cfg, n, out); }
int create_for_let(const nbody_cuda_config* cfg, uint64_t n, uint64_t cap, uint64_t halo, Sim** out) { return create_common(cfg, n, out, cap, halo); }
int grow_pools_after_overflow(Sim& s, uint32_t status) { return grow_after_overflow(s, status); }
int realloc_nodes(Sim& s, uint32_t max_nodes, uint32_t src_nodes) { free_nodes(s); return alloc_nodes(s, max_nodes, src_nodes); }
void destroy_for_comm(Sim* s) { free_all(s); }
}  // namespace nbody
