/*
 * nbody_cuda.h — C ABI of the B200-native FMM gravity solver.
 *
 * This is the drop-in boundary for the reference's simulation interface
 * (duanebyer/nbody; citations relative to the reference tree):
 *   - nbody::Simulation<TScalar,TVector>::step() / ::particles()
 *         include/nbody/simulation.h:6-37
 *   - OpenClSimulation(bounds, particles, timeStep, log)
 *         include/nbody/open_cl_simulation.h:194-198, src/open_cl_simulation.cpp:15-51
 *   - NaiveSimulation(particles, forceConstant, timeStep)
 *         include/nbody/naive_simulation.h:25-33
 * The reference defines no FFI; a maintainer binds these entry points from the
 * C++ class in include/nbody/cuda_simulation.h.
 *
 * Plain C: pointers, sizes and POD structs only. No exception crosses this
 * boundary: every call returns NBODY_OK (0) or an error code, and
 * nbody_cuda_last_error() returns the message of the last failure on the
 * calling thread. There is no CPU fallback: without a CUDA device (sm_100)
 * nbody_cuda_create fails with NBODY_ERR_CUDA.
 */
#ifndef NBODY_CUDA_H_
#define NBODY_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBODY_CUDA_ABI_VERSION 1

/* Particle record at the boundary: layout of Simulation::Particle instantiated
 * with the reference's 16-byte vector (include/nbody/simulation.h:16-35,
 * include/nbody/device/types.h:53-74): 48 bytes, 16-byte aligned. */
typedef struct nbody_particle {
	float position[4];
	float velocity[4];
	float mass;
	float charge;
	float _pad[2];
} nbody_particle;

enum {
	NBODY_OK = 0,
	NBODY_ERR_INVALID = 1,   /* bad argument */
	NBODY_ERR_CUDA = 2,      /* CUDA runtime failure / no device */
	NBODY_ERR_CAPACITY = 3,  /* a device pool could not be grown */
	NBODY_ERR_STATE = 4,     /* call not valid in the current state */
	NBODY_ERR_COMM = 5       /* NCCL failure */
};

/* integrators */
enum {
	NBODY_KICK_DRIFT = 0,     /* v += a dt; x += v_new dt  (src/naive_simulation.cpp:28-42) */
	NBODY_EXPLICIT_EULER = 1  /* v += a dt; x += v_old dt  (src/open_cl_simulation.cpp:602-607) */
};

/* flags */
enum {
	NBODY_FLAG_KEEP_LISTS = 1u,   /* keep interaction lists after step() for nbody_cuda_get_lists */
	NBODY_FLAG_NO_INTEGRATE = 2u, /* step() computes accelerations only (state is re-ordered, not advanced) */
	NBODY_FLAG_DIRECT = 4u,       /* all-pairs direct sum instead of the FMM (validation / P2P microbenchmark) */
	NBODY_FLAG_CUB_SORT = 8u,     /* sort with CUB's DeviceRadixSort instead of the built-in radix sort (comparison only) */
	NBODY_FLAG_STATIC_PARTITION = 16u, /* multi-GPU: equal particle counts per rank every step instead of per-step load rebalancing */
	NBODY_FLAG_NO_OVERLAP = 32u,      /* multi-GPU: exchange the velocities on the compute stream inside step() instead of on a second
	                                     stream overlapped with the next step's sort / tree / traversal (comparison only) */
	NBODY_FLAG_PARTITIONED = 128u,    /* multi-GPU: every rank holds only the particles of its Morton-key range and imports a locally
	                                     essential tree (the other ranks' trees by NCCL all-gather, halo particles by NVLink peer loads)
	                                     instead of replicating the particle state and the tree on every rank. Memory per GPU ~ N/W + halo. */
	NBODY_FLAG_DIST_SORT = 64u        /* multi-GPU (replicated scheme): every rank sorts only the keys of its own slice, the sorted runs are
	                                     all-gathered and merged pairwise, instead of every rank sorting all keys. Same permutation
	                                     bit for bit. Ignored on a single GPU and with NBODY_FLAG_CUB_SORT. */
};

typedef struct nbody_cuda_config {
	uint32_t abi_version;   /* NBODY_CUDA_ABI_VERSION */
	float bounds[4];        /* root box [0,bounds) (src/open_cl_simulation.cpp:20,44; src/main.cpp:24) */
	float time_step;        /* src/main.cpp:69 */
	float force_constant;   /* naive convention: > 0 attracts like charges; default +1 */
	float softening;        /* PARTICLE_RADIUS, src/field.cl:3-5; default 0.01 */
	float mac_ratio;        /* NODE_APPROX_RATIO, src/interaction.cl:3-5; default 0.5 */
	uint32_t leaf_capacity; /* octree node capacity, src/open_cl_simulation.cpp:41-47; default 8 */
	uint32_t max_depth;     /* <= 21 */
	uint32_t order;         /* expansion order P in {2,3,4,5}; default 4 (5: 56 coefficients, the M2L kernel then runs one CTA per SM at 254 registers) */
	uint32_t integrator;    /* NBODY_KICK_DRIFT (default) | NBODY_EXPLICIT_EULER */
	uint32_t flags;
	int32_t device;         /* CUDA device ordinal; -1 = current */
	float pool_scale;       /* multiplies the initial sizes of the list pools; default 1 */
	float low_order_tau;    /* adaptive-order M2L: accepted pairs with 0.75 (dimA+dimB)^2 / d^2 < tau are evaluated at
	                           order P-1 (same lists, less work; 0 = always order P). default 0.13: RMS error 4.4e-4 on
	                           the Plummer model against 2.1e-4 at full order (tools/explore notes in DESIGN.md) */
	/* Variable time step (the reference's TODO:2-3 "variable timestep"; it ships fixed steps only, src/main.cpp:69).
	 * time_step_eta = 0 (default): every step uses time_step, like the reference. > 0: step n uses dt_n, with
	 * dt_0 = time_step and dt_{n+1} = clamp(eta * sqrt(len / max_i |a_i|), time_step_min, time_step_max), the maximum
	 * taken over the accelerations step n computed; len = softening, or bounds[0] * 2^-max_depth when softening is 0.
	 * time_step_max = 0 means time_step; time_step_min = 0 means no lower bound. */
	float time_step_eta;
	float time_step_min;
	float time_step_max;
	uint32_t partition_slack_pct; /* partitioned mode: room for own particles = (100 + this) % of N / ranks (migration and rebalancing change a
	                                 rank's count); 0 = automatic (50 % up to 16M particles per rank, less above) */
	uint32_t _reserved[2];
} nbody_cuda_config;

/* per-step statistics (tracing/metrics hook) */
typedef struct nbody_cuda_stats {
	uint64_t n_particles;
	uint64_t n_nodes, n_leaves, n_levels;
	uint64_t m2l_entries;      /* (source, target-mask) entries in the grouped M2L lists */
	uint64_t m2l_interactions; /* directed target<-source M2L evaluations */
	uint64_t m2l_interactions_low; /* ... of which evaluated at order P-1 (low_order_tau) */
	uint64_t p2p_entries;      /* directed target-leaf <- source-leaf pairs */
	uint64_t p2p_interactions; /* directed target-particle <- source-particle evaluations */
	uint64_t near_entries;     /* total near-list entries written by the traversal */
	uint64_t retries;          /* times a step was re-run after growing a pool */
	uint64_t device_bytes;     /* bytes currently allocated on the device */
	/* device time of the last step, milliseconds (CUDA events on the compute stream) */
	float ms_total, ms_sort, ms_tree, ms_upsweep, ms_traverse, ms_m2l, ms_l2l, ms_leaf, ms_comm;
	float work_imbalance;      /* multi-GPU: max over ranks / mean over ranks - 1 of the owned-slice device time of the last step */
	/* partitioned mode (NBODY_FLAG_PARTITIONED); 0 otherwise */
	uint64_t halo_particles;     /* other ranks' particles fetched for the P2P lists of the last step */
	uint64_t imported_nodes;     /* nodes of the other ranks' trees held as sources */
	uint64_t migrated_particles; /* particles that arrived from other ranks at the start of the last step */
	float ms_import;             /* the parts of ms_comm: all-gather of the ranks' trees (node counts, then the node records over NCCL) */
	float ms_halo;               /*   halo: mark the imported leaves the P2P lists name, prefix sum, NVLink fetch, list fix-up */
	float ms_balance;            /*   end of step: work times and next step's splitters */
	float _pad;
} nbody_cuda_stats;

typedef struct nbody_cuda_sim nbody_cuda_sim; /* opaque; single-threaded use */

/* Fill cfg with the reference's constants (bounds 1,1,1; dt 0.001; G +1; eps 0.01;
 * MAC 0.5; capacity 8; depth 21; order 4; kick-drift; low_order_tau 0.13). */
void nbody_cuda_default_config(nbody_cuda_config* cfg);
/* The same with the B200 tuning of the one constant the reference itself marks as provisional (octree node capacity 8,
 * "FIXME ... should be adjustable by the user", src/open_cl_simulation.cpp:41-47): leaf_capacity = 48. Same physics (MAC,
 * softening, order, integrator), same accuracy (the error is set by the upper tree levels, which are identical: 4.3e-4 RMS at
 * both, Plummer 2^24), 2.7x the throughput on a B200 (64 against 177 ms per step). nbody::CudaSimulation uses this when no
 * configuration is passed; tree-topology parity against the reference's contract is tested at capacities 3, 8, 32 and 48. */
void nbody_cuda_tuned_config(nbody_cuda_config* cfg);

/* Construct from host particles (copied; caller keeps ownership of `particles`).
 * Replaces the OpenClSimulation ctor, src/open_cl_simulation.cpp:15-51. */
int nbody_cuda_create(const nbody_cuda_config* cfg, const nbody_particle* particles, uint64_t n, nbody_cuda_sim** out);
void nbody_cuda_destroy(nbody_cuda_sim* sim);

/* Replace the whole particle state from host memory (same n). */
int nbody_cuda_set_particles(nbody_cuda_sim* sim, const nbody_particle* particles, uint64_t n);

/* Advance one time step; blocking; *time_out = new simulation time (FP32
 * accumulation, src/open_cl_simulation.cpp:103-105). Replaces Simulation::step(). */
int nbody_cuda_step(nbody_cuda_sim* sim, float* time_out);

uint64_t nbody_cuda_num_particles(const nbody_cuda_sim* sim);

/* Copy the particle state to host, in tree (Morton/DFS leaf) order like
 * OpenClSimulation::particles(), src/open_cl_simulation.cpp:53-68 (input order
 * before the first step). Replaces Simulation::particles(). */
int nbody_cuda_get_particles(nbody_cuda_sim* sim, nbody_particle* out, uint64_t capacity);

/* orig_index[i] = index in the constructor's array of the particle now at i. */
int nbody_cuda_get_permutation(nbody_cuda_sim* sim, uint32_t* orig_index, uint64_t capacity);

/* Accelerations of the last step, xyz per particle, same order as get_particles.
 * (Distributed: collective — every rank must call it; the slices are exchanged on first use.) */
int nbody_cuda_get_accelerations(nbody_cuda_sim* sim, float* xyz, uint64_t capacity);

/* ---- parity exports (tests only) ---------------------------------------- */
/* Sorted Morton keys of the last step. */
int nbody_cuda_get_keys(nbody_cuda_sim* sim, uint64_t* keys, uint64_t capacity);
/* Octree of the last step in the reference's DFS pre-order node_t contract
 * (include/nbody/device/types.h:124-141). Any pointer may be NULL.
 * Call with all NULL to obtain *n_nodes. child_off9: 9 per node. geom4: centre xyz + dimensions.x */
int nbody_cuda_get_tree(nbody_cuda_sim* sim, uint32_t* n_nodes, uint32_t capacity, uint32_t* depth, uint64_t* prefix,
                        uint32_t* leaf_index, uint32_t* leaf_count, uint8_t* has_children, uint32_t* child_off9,
                        int32_t* parent_off, uint32_t* sibling, float* geom4);
/* Directed interaction lists of the last step (needs NBODY_FLAG_KEEP_LISTS), as
 * (target, source) pairs of DFS node ids. Call with NULL to obtain the counts. */
int nbody_cuda_get_lists(nbody_cuda_sim* sim, uint64_t* n_m2l, uint32_t* m2l_pairs, uint64_t* n_p2p, uint32_t* p2p_pairs);
/* Multipole (M) and local (L) coefficients per DFS node, ncoef(order) floats each.
 * M_m = sum q (y-c)^m/m!;  L_n = d^n Phi / dx^n at the cell centre (far field only). */
int nbody_cuda_get_expansions(nbody_cuda_sim* sim, float* multipoles, float* locals, uint64_t capacity_floats);

int nbody_cuda_get_stats(nbody_cuda_sim* sim, nbody_cuda_stats* stats);

/* ---- variable time step -------------------------------- */
/* Override the time step the NEXT step() uses (caller-driven variable steps; with time_step_eta > 0 the library
 * replaces it again after that step). dt must be finite and > 0. Distributed: every rank must pass the same value. */
int nbody_cuda_set_time_step(nbody_cuda_sim* sim, float dt);
/* *next_dt = the step the next step() will take; *last_dt = the one the last step() took (0 before the first);
 * *acc_max = max_i |a_i| of the last step (0 unless time_step_eta > 0). Any pointer may be NULL. */
int nbody_cuda_get_time_step(nbody_cuda_sim* sim, float* next_dt, float* last_dt, float* acc_max);
/* The rule on its own (host arithmetic, no device needed): the step that follows a step whose largest acceleration
 * was acc_max under configuration cfg. Returns cfg->time_step when time_step_eta <= 0 or acc_max is not finite and > 0. */
float nbody_cuda_next_time_step(const nbody_cuda_config* cfg, float acc_max);

/* ---- checkpoint / restart ------------------------------- */
/* The reference's only output is particles.csv, which drops velocities, masses and charges (src/main.cpp:88-95), so a
 * run cannot be resumed from it. A checkpoint file holds the full state: configuration, simulation time (the FP32
 * accumulator step() returns), step count, the time step the next step() will use, the 48-byte particle records in the
 * order particles() returns them and the permutation nbody_cuda_get_permutation returns. Layout: nbody_checkpoint_header,
 * n * 48 bytes, n * 4 bytes; little-endian; `checksum` = sum over i of (w_i + 0x9E3779B97F4A7C15) * (2 i + 1) mod 2^64, w_i the
 * 64-bit words of the header (19 words, its checksum field taken as zero), then of the particle array, then of the permutation
 * (zero-padded to a whole word); version 1 files, whose sum covers the two arrays only, are still read. Resuming reproduces the
 * uninterrupted run bit for bit with NBODY_FLAG_DIRECT; the FMM path sums its interaction lists in an order that depends
 * on kernel timing, so there (as between any two runs of it) the states agree to FP32 round-off, not bitwise. */
#define NBODY_CHECKPOINT_MAGIC 0x31504b435944424eull /* the bytes "NBDYCKP1" */
#define NBODY_CHECKPOINT_VERSION 2u
typedef struct nbody_checkpoint_header {
	uint64_t magic;
	uint32_t version;
	uint32_t header_bytes;      /* sizeof(nbody_checkpoint_header) */
	uint64_t n_particles;
	uint64_t steps_done;
	float time;                 /* simulation time */
	float next_time_step;       /* the dt the next step() takes */
	float last_time_step;
	float last_acc_max;
	uint64_t checksum;
	nbody_cuda_config config;
} nbody_checkpoint_header;

/* Write the state of `sim` to `path` (blocking). Distributed, replicated scheme: every rank holds the whole state and may call
 * it (each with its own path); it is a collective only in that every rank must have finished the same step. Partitioned scheme
 * (NBODY_FLAG_PARTITIONED): NBODY_ERR_STATE — a rank holds only its own particles and the file format is single-rank; save per
 * rank with nbody_cuda_get_owned_particles + nbody_cuda_get_permutation + nbody_cuda_checkpoint_write. */
int nbody_cuda_checkpoint_save(nbody_cuda_sim* sim, const char* path);
/* Host-only (no device needed): read and validate the header of a checkpoint file. */
int nbody_cuda_checkpoint_info(const char* path, nbody_checkpoint_header* header);
/* Host-only: read the payload; verifies the checksum. Either pointer may be NULL. capacity in particles. */
int nbody_cuda_checkpoint_read(const char* path, nbody_particle* particles, uint32_t* orig_index, uint64_t capacity);
/* Host-only: write a checkpoint from host arrays (orig_index NULL = identity); fills magic, version, sizes, checksum. */
int nbody_cuda_checkpoint_write(const char* path, const nbody_checkpoint_header* header, const nbody_particle* particles,
                                const uint32_t* orig_index);
/* Construct a simulation that continues where the checkpoint stopped. cfg NULL = the configuration stored in the
 * file (on the current device); otherwise cfg replaces it (bounds, order, capacity ... may change between runs). */
int nbody_cuda_checkpoint_load(const char* path, const nbody_cuda_config* cfg, nbody_cuda_sim** out);
/* Simulation time and number of steps taken (what a checkpoint stores). */
int nbody_cuda_get_time(nbody_cuda_sim* sim, float* time, uint64_t* steps_done);

/* All-pairs softened field of `n_src` sources (x,y,z,q) on `n_tgt` targets (x,y,z,*)
 * with the tiled P2P kernel; host buffers in, field (not yet scaled by G q/m) out.
 * Used to validate large runs against direct summation on a target subsample and as
 * the P2P FP32 microbenchmark; *ms = kernel time. */
int nbody_cuda_direct_field(int device, const float* src_posq, uint64_t n_src, const float* tgt_pos4, uint64_t n_tgt,
                            float softening, float* field_xyz, float* ms, uint32_t repeats);

/* The device side of the distributed sort (NBODY_FLAG_DIST_SORT) on ONE GPU, without NCCL: `keys` (host, 63-bit) is cut
 * into `nruns` slices with boundaries bound[0..nruns] (bound[0] = 0, bound[nruns] = n, 1..16 runs); every slice is radix-sorted
 * as a rank would sort its own, then the runs are merged pairwise as every rank does after the all-gather. Returns the sorted
 * keys and, per output position, the input index: by contract the stable sort of all keys, whatever the boundaries.
 * A validation entry point (tests, bring-up of the multi-GPU path on a one-GPU box); no reference counterpart. */
int nbody_cuda_sort_runs(int device, const uint64_t* keys, uint64_t n, const uint32_t* bound, int nruns, uint64_t* keys_out,
                         uint32_t* index_out);

/* ---- multi-GPU (one process per GPU; Morton-range partition; NCCL + NVLink peer memory) ------- */
/* The reference is single-device (src/open_cl_simulation.cpp:627-632); this part of the ABI has no counterpart there. */
/* 128-byte NCCL unique id created on rank 0 and sent to the other ranks by the caller. */
int nbody_cuda_comm_unique_id(uint8_t id[128]);
/* Same as nbody_cuda_create, but this rank passes only ITS slice of the global particle set; `global_offset` is the index of its
 * first particle (the slices need not be sorted in any way). Collective: every rank calls it with the same cfg->flags.
 * With NBODY_FLAG_PARTITIONED the object then holds ONLY the particles of this rank's Morton-key range: nbody_cuda_num_particles,
 * get_particles, get_permutation (indices into the GLOBAL input array), get_keys, get_accelerations and get_stats answer for those
 * particles, in tree order (ranks in rank order = the global tree order), and their number changes from step to step.
 * Without the flag (round 1's replicated scheme) every rank holds the whole state and the calls answer for all n_global particles. */
int nbody_cuda_create_distributed(const nbody_cuda_config* cfg, const nbody_particle* local_particles, uint64_t n_local,
                                  uint64_t n_global, uint64_t global_offset, int rank, int world, const uint8_t id[128],
                                  nbody_cuda_sim** out);
/* Virtual ranks: the partitioned scheme (NBODY_FLAG_PARTITIONED) with `world` ranks inside ONE process on ONE GPU — the same
 * phases and kernels, peer pointers to the other members' arrays instead of cudaIpc mappings, device-to-device copies instead of
 * NCCL. Rank r starts with particles [n r / world, n (r+1) / world). A validation entry point (the parity tests run 2, 4 and 8
 * ranks on a one-GPU box); each member answers the per-rank calls (get_particles, get_permutation, get_keys, get_accelerations,
 * get_stats: its own particles, in tree order; concatenated in rank order they are the global tree order). */
int nbody_cuda_create_group(const nbody_cuda_config* cfg, const nbody_particle* particles, uint64_t n, int world, nbody_cuda_sim** sims_out);
int nbody_cuda_group_step(nbody_cuda_sim** sims, int world, float* time_out);
void nbody_cuda_destroy_group(nbody_cuda_sim** sims, int world);

/* Range [first, first+count) of the tree-ordered particle array owned by this rank after the last step
 * (before the first step: the slice passed to nbody_cuda_create_distributed). Single GPU: [0, n). */
int nbody_cuda_owned_range(nbody_cuda_sim* sim, uint64_t* first, uint64_t* count);
/* Host <-> device transfer of this rank's OWNED slice only (count particles, tree order): the distributed
 * counterparts of get_particles / set_particles. Partitioned scheme: purely local (no other rank is involved).
 * Replicated scheme: set_ is collective, every rank must call it. */
int nbody_cuda_get_owned_particles(nbody_cuda_sim* sim, nbody_particle* out, uint64_t capacity);
int nbody_cuda_set_owned_particles(nbody_cuda_sim* sim, const nbody_particle* particles, uint64_t n);
/* Per-step load rebalancing, the host-side rule on its own (no device needed): from last step's slice boundaries
 * (world + 1 entries, first 0, last n) and each rank's device time for its slice, the wanted boundaries of the next
 * step (world + 1 entries; before they are snapped to leaf boundaries). damping in (0,1]: fraction of the correction
 * applied per step (the library uses 0.5); <= 0 returns the input boundaries. step() applies this internally unless
 * NBODY_FLAG_STATIC_PARTITION is set. */
int nbody_cuda_rebalance(int world, const uint32_t* boundaries, const float* work_ms, float damping, uint32_t* out);

const char* nbody_cuda_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_CUDA_H_ */
