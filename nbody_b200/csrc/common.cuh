// Shared declarations of the B200 FMM gravity solver (sm_100a only).
// HBM layout, control block and the launch functions of each stage.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: the ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include "../../include/nbody_cuda.h"
#include "expansion.cuh"

namespace nbody {

constexpr int kMaxDepth = 21;           // 21 bits per dimension, 63-bit keys
constexpr int kNumLevels = kMaxDepth + 1;
constexpr int kNumSM = 148;             // B200: grids are sized in multiples of this
constexpr int kScanBlocks = 592;        // 4 * 148 tiles per level scan (<= 1024)

// status bits raised by kernels when a pool is too small; the host grows the pool and re-runs the step
enum : uint32_t {
	kOvfNodes = 1u, kOvfNear = 2u, kOvfP2P = 4u, kOvfM2L = 8u, kOvfSeg = 16u, kOvfGroups = 32u, kOvfItems = 64u,
	kOvfDepth = 128u,  // the tree wanted to grow below the depth the level loops were bounded to (last step's depth + 1): re-run unbounded
	kOvfHalo = 256u    // partitioned mode: the halo particles do not fit behind the own particles
};
constexpr int kMaxRanks = 16;
constexpr uint32_t kImported = 0x80000000u;  // partitioned mode: a first-particle index with this bit names an imported leaf, not a particle

// One work item of the traversal / of the M2L kernel: `nt` sibling targets
// (8 = the children of one parent, 1 = a carried childless node) that share one
// candidate list.
struct Group {
	uint32_t first;      // first target node id (level-major)
	uint32_t nt;         // 8 or 1
	uint32_t list_off;   // traversal: offset of the shared near list (in the previous level's pool);  M2L: offset into the m2l pool
	uint32_t list_cnt;   // entries in that list
	uint32_t n_cand;     // traversal: number of candidate slots the near list expands to (8 per split entry, 1 per childless)
	uint32_t _pad[3];
};

struct Segment { uint32_t off, cnt, next; };
struct TraverseSeeds { uint32_t n = 1; uint32_t id[16] = {}; };  // source roots of the traversal: the own root (0), then the imported trees' roots  // one piece of a target leaf's P2P source list (chained)

// Device-resident control block: everything the host would otherwise have to read back between launches.
struct Ctrl {
	uint32_t level_off[kNumLevels + 2];  // nodes of level l are [level_off[l], level_off[l+1])
	uint32_t n_nodes;
	uint32_t status;
	uint32_t scan_ticket;
	uint32_t n_levels;
	uint32_t gq_count[2];                // traversal group queues, ping-pong by round
	uint32_t items_count[2];             // M2L work items: [0] = 8-target groups, [1] = single targets
	uint32_t seg_cursor;
	uint32_t _pad0;
	unsigned long long near_cursor[2];   // near-list pools, ping-pong by round
	unsigned long long p2p_cursor;
	unsigned long long m2l_cursor;
	unsigned long long stat_m2l_inter, stat_m2l_low, stat_p2p_entries, stat_p2p_inter, stat_near, stat_leaves;
	uint32_t work_ticket[4];             // dynamic work distribution of the persistent kernels
	uint32_t part[kMaxRanks + 1];        // distributed: rank r owns particles [part[r], part[r+1]) of the tree-ordered array
	uint32_t n_leaf_items;               // entries of the compacted leaf list (k_leaf_items; small trees only)
	uint32_t acc_max2_bits;              // variable time step: bits of max |a|^2 over this rank's slice (k_acc_max; non-negative floats order like uints)
};

struct Pools {
	uint32_t* near[2] = {nullptr, nullptr};
	uint64_t near_cap = 0;
	uint2* p2p = nullptr;            // P2P source lists: {first particle, particle count} of each source leaf
	uint64_t p2p_cap = 0;
	uint32_t* m2l_id = nullptr;
	uint8_t* m2l_mask = nullptr;     // bit t: target t of the group accepts this candidate
	uint8_t* m2l_mask_lo = nullptr;  // bit t: ... and evaluates it at order P-1 (subset of m2l_mask)
	uint64_t m2l_cap = 0;
	Segment* seg = nullptr;
	uint32_t seg_cap = 0;
	Group* gq[2] = {nullptr, nullptr};
	uint32_t gq_cap = 0;
	Group* items[2] = {nullptr, nullptr};
	uint32_t items_cap = 0;
};

struct Comm;  // NCCL state (comm.cu)
struct Let;   // partitioned mode: own particles + locally essential tree (let.cu)

// Partitioned mode, device side. Rank r owns the particles whose Morton key lies in [split[r], split[r+1]).
// A cell that contains a splitter strictly inside its key range holds particles of several ranks ("straddling" cell; at most one
// per splitter and level): whether it splits is decided by its GLOBAL particle count, so that every rank's tree is the global
// octree restricted to the cells that hold its own particles.
struct LetCtrl {
	uint64_t split[kMaxRanks + 1];               // split[0] = 0, split[world] = 2^63
	uint32_t force[kMaxRanks][kNumLevels];       // [b][d]: global particle count of the depth-d cell that straddles splitter b (0: none)
	int world, rank;
	uint32_t imp_base;                           // node id of the first imported node (= the own tree's capacity)
	uint32_t imp_off[kMaxRanks + 1];             // imported tree of rank s occupies node ids [imp_base + imp_off[s], imp_base + imp_off[s+1])
	uint32_t halo_base, halo_cap, halo_count;    // halo particles live at posq[1][halo_base ...)
	unsigned long long t_begin, t_end;           // %globaltimer around the stages whose cost follows the partition (rebalancing input)
};

struct Sim {
	nbody_cuda_config cfg;
	int device = 0;
	cudaStream_t stream = nullptr;
	uint64_t n = 0;           // particles this object holds (partitioned mode: the rank's own particles in the current step)
	uint64_t cap = 0;         // allocated length of the per-particle arrays (= n except in partitioned mode, where n varies)
	uint64_t src_cap = 0;     // allocated length of posq[1] (partitioned mode: cap + room for the halo particles)
	uint64_t n_global = 0;
	float time = 0.0f;
	uint64_t steps_done = 0;  // steps taken by this object (0 = no tree, lists or accelerations yet)
	uint64_t steps_base = 0;  // steps taken before the checkpoint this object was loaded from
	float dt = 0.0f;          // the step the next step() takes (cfg.time_step unless the variable time step or the caller changed it)
	float dt_last = 0.0f;     // the step the last step() took
	float acc_max = 0.0f;     // max |a| of the last step (time_step_eta > 0 only)
	int nc_stride = 0;  // floats per multipole/local record

	// particle state, SoA of two float4 planes; [0] = state (order of the last step), [1] = sorted scratch of the current step
	float4* posq[2] = {nullptr, nullptr};  // x y z charge
	float4* velm[2] = {nullptr, nullptr};  // vx vy vz mass
	uint32_t* orig[2] = {nullptr, nullptr};
	float4* acc = nullptr;                 // ax ay az (w unused), order of the last step
	uint64_t* keys[2] = {nullptr, nullptr};
	uint32_t* idx[2] = {nullptr, nullptr};
	void* sort_tmp = nullptr;
	size_t sort_tmp_bytes = 0;
	nbody_particle* aos_dev = nullptr;     // staging for AoS48 <-> SoA conversion
	nbody_particle* aos_host = nullptr;    // pinned

	// octree, level-major; the 8 children of a split node are contiguous
	uint32_t max_nodes = 0;      // capacity of the own tree
	uint32_t src_nodes = 0;      // allocated length of the source-side arrays geom / info / nbegin (>= max_nodes; partitioned mode:
	                             // the other ranks' trees are imported at ids [max_nodes, src_nodes))
	int depth_bound = kMaxDepth; // the level loops of a step run to this depth (last step's depth + 1; kOvfDepth re-runs unbounded)
	TraverseSeeds seeds;
	uint32_t* imp_hoff = nullptr;   // partitioned mode (let.cu): per imported leaf, its particle count once a P2P list names it ...
	uint32_t* imp_mflag = nullptr;  // ... and per imported node, whether an M2L list names it: both marked by the traversal as it writes the lists
	int trav_bound = kMaxDepth;  // rounds of the traversal (partitioned mode: the deepest tree of any rank; otherwise depth_bound)
	float4* geom = nullptr;      // centre xyz, dimensions.x
	uint2* info = nullptr;       // {first child (0 = childless), particle count}
	uint32_t* nbegin = nullptr;  // first particle
	uint32_t* nparent = nullptr;
	uint64_t* nkey = nullptr;    // key prefix
	float* M = nullptr;          // multipoles, nc_stride floats per node
	float* L = nullptr;          // locals (pure derivatives), nc_stride floats per node
	float* Mimp = nullptr;       // partitioned mode: multipole orders 0..P-1 of the imported nodes the M2L lists name (let.cu)
	uint2* near_ref = nullptr;   // per target node: {offset, count} of its near list in the current round's pool
	uint32_t* p2p_head = nullptr;  // per node: head of its P2P segment chain (0xffffffff = none)
	uint32_t* leaf_items = nullptr;  // compacted list of the slice's non-empty leaves (the leaf kernel's tickets on small trees)
	uint32_t* scan_sums = nullptr; // kScanBlocks + 1

	Pools pools;
	Ctrl* ctrl = nullptr;        // device
	Ctrl* ctrl_host = nullptr;   // pinned copy read after each step
	Comm* comm = nullptr;
	Let* let = nullptr;
	LetCtrl* let_ctrl = nullptr;   // device (nullptr unless partitioned)
	// distributed: this rank's slice of the tree-ordered particle array
	uint64_t own_first = 0, own_count = 0;
	int rank = 0;                // index into Ctrl::part (0 on a single GPU)

	nbody_cuda_stats stats{};
	cudaEvent_t ev[12] = {};
	uint64_t device_bytes = 0;
	bool lists_valid = false;
	bool acc_partial = false;  // distributed: `acc` holds only this rank's slice until comm_exchange_acc()
};

// ---- tracing hook: one NVTX range per stage of a step, around the stage's launches -------------------------
struct NvtxRange {
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
	NvtxRange(const NvtxRange&) = delete;
	NvtxRange& operator=(const NvtxRange&) = delete;
};

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string& msg);
#define NB_CUDA_CHECK(expr)                                                                         \
	do {                                                                                               \
		cudaError_t _e = (expr);                                                                         \
		if (_e != cudaSuccess) {                                                                         \
			::nbody::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
			return NBODY_ERR_CUDA;                                                                         \
		}                                                                                                \
	} while (0)

// ---- stage launchers (each enqueues on sim.stream; no host synchronisation) --
void launch_import(Sim& s, const nbody_particle* aos_dev, uint64_t n);                 // AoS48 -> SoA state
void launch_import_ids(Sim& s, const nbody_particle* aos_dev, uint64_t n, uint32_t first_id, bool keep_ids);
void launch_keys(Sim& s, const float4* pos, uint64_t n);                               // keys[0] / idx[0] of pos[0..n)
void launch_gather(Sim& s, uint64_t n);                                                // state [0] -> sorted [1] through idx[0]
void launch_exclusive_scan(Sim& s, uint32_t* data, uint32_t n);                        // in place, scratch = sort_tmp (sort.cu)
void launch_export(Sim& s, nbody_particle* aos_dev, uint64_t n);                       // SoA state -> AoS48
int launch_keys_sort_permute(Sim& s);                                                 // stage 1a: keys, radix sort, gather
void launch_own_sort_range(Sim& s, uint64_t first, uint64_t n);                        // the radix sort over a sub-range of keys[0] / idx[0]
void launch_merge_runs(Sim& s, const uint32_t* bound, int nruns);                      // distributed sort: pairwise merges of sorted runs
int comm_sort_exchange(Sim& s, uint32_t* bound, int* nruns);                           // distributed sort: all-gather of the sorted runs (comm.cu)
void comm_own_slice(const Sim& s, uint64_t* first, uint64_t* count);                   // distributed: this rank's slice of the state order
int comm_wait_positions(Sim& s);                                                      // distributed sort: order the stream after a position exchange on the second stream
void launch_gather_velocities(Sim& s);                                                 // distributed: the velocity half of stage 1a's gather, delayed until the leaf kernel needs it
void launch_tree_build(Sim& s);                                                        // stage 1b: linear octree, level-major
void launch_upsweep(Sim& s);                                                           // stage 2: P2M + M2M
void launch_traversal(Sim& s);                                                         // stage 3: dual-tree traversal -> lists
void launch_m2l(Sim& s);                                                               // stage 4a: M2L over grouped lists
void launch_l2l(Sim& s);                                                               // stage 4b: L2L downsweep
void launch_leaf(Sim& s);                                                              // stage 5: P2P + L2P + integrator
void launch_direct(Sim& s);                                                            // all-pairs P2P + integrator (validation)
void launch_acc_max(Sim& s);                                                           // variable time step: max |a|^2 of the owned slice -> Ctrl
int direct_field_device(const float4* src, uint64_t n_src, const float4* tgt, uint64_t n_tgt, float eps2, float4* out, cudaStream_t st);

// device helpers shared by several translation units
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Sum each of 32 values over the 32 lanes of the warp: every halving step hands the partner lane the half of
// the values it is responsible for, so 31 shuffles do the work of 160. On return v[0] of lane l holds the total of value l.
__device__ __forceinline__ void transpose_reduce32(float (&v)[32], unsigned lane) {
#pragma unroll
	for (int h = 16; h >= 1; h >>= 1) {  // partner = lane ^ h: lanes with that bit set keep the upper h values
		const bool up = (lane & (unsigned) h) != 0u;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
			v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
		}
	}
}

}  // namespace nbody
