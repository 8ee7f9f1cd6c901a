// ============================================================================
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's OpenCL C device kernels, compiled for the HOST from the files where they lie
// under /root/reference/src (oracle/Makefile, target `clref`, output oracle/_ref/libclref.so):
//   interaction.cl, field.cl, verify.cl   — byte for byte as they are;
//   moment.cl, force.cl                   — a transient copy under oracle/_ref/gen/ (deleted after the link) in which the OpenCL vector
//                                           literal "(vector_t) (" reads "make_vector_t(" (a C++ compiler parses
//                                           the literal as a cast of a comma expression); nothing else changes.
// oracle/shim_cl/opencl_c_host.h supplies float4, dot, sqrt, min, the work-item functions and the
// atomics. The reference's HOST side (src/open_cl_simulation.cpp) cannot be compiled — it is written
// against glade::Orthtree, which is not in the tree (CMakeLists.txt:33) — so the orchestration
// around the kernels is restated below, each block citing the lines it follows, with the reference's
// memory batching (:179-186, :286-325) left out: everything is processed in one batch, which is what
// the reference does whenever its buffers fit.
//
// What this pins (tests/test_reference_kernels.py): given one octree (the oracle's — the octree itself
// stays unpinned, it is glade's), the reference's own kernels produce the interaction lists, the
// leaf moments, the near-field pair forces and the monopole far field the oracle restates.
//
// Reference defects the harness can switch off, because they sit in code restated HERE or because the
// kernel's intent is unambiguous:
//   D5 (:156-157, the compaction loop is bounded by a counter that was just zeroed, so the upsweep
//       never runs): `repair` runs the loop over the whole array, as its comment describes;
//   D7 (src/interaction.cl:116-122, the atomic_inc results go to a private copy): `repair` stores
//       the two indices the kernel computes into the interaction records;
//   far-field work-group size (src/field.cl:189-201 covers at most get_local_size(0) leaves of the
//       target node; the host launches CL_KERNEL_PREFERRED_WORK_GROUP_SIZE_MULTIPLE work items,
//       src/open_cl_simulation.cpp:951-957): the caller chooses the local size.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#define __OPENCL_VERSION__ 120  // include/nbody/device/types.h:4-6 selects its device-side typedefs with this

namespace refcl {
#define REFCL_INSIDE_NAMESPACE 1
#include "opencl_c_host.h"
thread_local WorkItem g_work_item;
#include "interaction.cl"   // /root/reference/src/interaction.cl (-I)
#include "field.cl"         // /root/reference/src/field.cl
#include "verify.cl"        // /root/reference/src/verify.cl
#include "gen/force.cl"     // oracle/_ref/gen/force.cl  (vector literals rewritten, see above)
#include "gen/moment.cl"    // oracle/_ref/gen/moment.cl
#undef kernel
#undef global
}  // namespace refcl

namespace {

using namespace refcl;

static_assert(sizeof(float4) == 16 && alignof(float4) == 16, "OpenCL float4");
static_assert(sizeof(leaf_t) == 48 && sizeof(node_t) == 160 && sizeof(interaction_t) == 20, "SURVEY 3.2 layout");
static_assert(offsetof(node_t, child_indices) == 36 && offsetof(node_t, has_children) == 88 && offsetof(node_t, value) == 96, "SURVEY 3.2 offsets");

struct State {
	std::vector<node_t> nodes;
	std::vector<leaf_t> leafs;
	std::vector<interaction_t> leaf_inter, node_inter;
	std::uint64_t rounds = 0;
};

// One NDRange launch, work items in ascending order; f is the kernel call.
template <class F>
void launch_1d(std::size_t groups, std::size_t local, F f) {
	g_work_item = WorkItem{};
	g_work_item.local_size[0] = local; g_work_item.local_size[1] = 1; g_work_item.local_size[2] = 1;
	for (std::size_t g = 0; g < groups; ++g)
		for (std::size_t l = 0; l < local; ++l) {
			g_work_item.group[0] = g; g_work_item.local[0] = l;
			f();
		}
}
// cl::NDRange(globalSize, localSize) / cl::NDRange(localSize, localSize): one group per item, 8 x 8 work items
template <class F>
void launch_8x8(std::size_t groups, F f) {
	g_work_item = WorkItem{};
	g_work_item.local_size[0] = 8; g_work_item.local_size[1] = 8; g_work_item.local_size[2] = 1;
	for (std::size_t g = 0; g < groups; ++g)
		for (std::size_t l1 = 0; l1 < 8; ++l1)
			for (std::size_t l0 = 0; l0 < 8; ++l0) {
				g_work_item.group[0] = g; g_work_item.local[0] = l0; g_work_item.local[1] = l1;
				f();
			}
}
// numItems / localSize + (numItems % localSize != 0) + (numItems == 0) groups of localSize (e.g. :786-792)
std::size_t groups_for(std::size_t items, std::size_t local) { return items / local + (items % local != 0) + (items == 0); }
constexpr std::size_t kLocal = 32;  // stands in for CL_KERNEL_PREFERRED_WORK_GROUP_SIZE_MULTIPLE where the kernel does not depend on it

}  // namespace

extern "C" {

// verify.cl: the sizes the reference checks against its host structs (src/open_cl_simulation.cpp:700-768)
void clref_type_sizes(std::uint32_t* sizes9) {
	launch_1d(1, 1, [&] { verify_device_type_sizes(sizes9); });
}

// The octree in the reference's node_t / leaf_t records (include/nbody/device/types.h:112-141) from the oracle's DFS arrays.
// position = cell index * dimensions per axis, dimensions = bounds * 2^-depth (glade's own arithmetic is unknown; for the
// power-of-two boxes the reference uses, src/main.cpp:24, every formula gives the same bits).
// particles12: sorted (tree-order) records {pos[4], vel[4], mass, charge, pad, pad}.
void* clref_create(std::uint32_t num_nodes, const std::uint32_t* depth, const std::uint64_t* prefix, const std::uint32_t* leaf_index,
                   const std::uint32_t* leaf_count, const std::uint8_t* has_children, const std::uint32_t* child_off9,
                   const std::int32_t* parent_off, const std::uint32_t* sibling, const float* bounds3, std::uint32_t num_leafs,
                   const float* particles12) {
	State* s = new State;
	s->nodes.resize(num_nodes);
	std::memset(s->nodes.data(), 0, sizeof(node_t) * num_nodes);
	for (std::uint32_t i = 0; i < num_nodes; ++i) {
		node_t& n = s->nodes[i];
		std::uint32_t ix = 0, iy = 0, iz = 0;
		for (std::uint32_t l = 0; l < depth[i]; ++l) {
			const std::uint32_t d = (std::uint32_t) (prefix[i] >> (3 * (20 - l))) & 7u;
			ix = ix << 1 | (d & 1u); iy = iy << 1 | (d >> 1 & 1u); iz = iz << 1 | (d >> 2 & 1u);
		}
		const float sc = std::ldexp(1.0f, -(int) depth[i]);
		n.dimensions = float4(bounds3[0] * sc, bounds3[1] * sc, bounds3[2] * sc, 0.0f);
		n.position = float4((float) ix * n.dimensions.x, (float) iy * n.dimensions.y, (float) iz * n.dimensions.z, 0.0f);
		n.depth = depth[i];
		for (int k = 0; k < 9; ++k) n.child_indices[k] = child_off9[9 * (std::size_t) i + k];
		n.parent_index = parent_off[i];
		n.sibling_index = sibling[i];
		n.leaf_count = leaf_count[i];
		n.leaf_index = leaf_index[i];
		n.has_children = has_children[i];
	}
	s->leafs.resize(num_leafs);
	std::memset(s->leafs.data(), 0, sizeof(leaf_t) * num_leafs);
	for (std::uint32_t i = 0; i < num_leafs; ++i) {
		const float* r = particles12 + 12 * (std::size_t) i;
		leaf_t& l = s->leafs[i];
		l.position = float4(r[0], r[1], r[2], 0.0f);
		l.value.velocity = float4(r[4], r[5], r[6], 0.0f);
		l.value.mass = r[8];
		l.value.moment.charge = r[9];
	}
	return s;
}
void clref_free(void* h) { delete (State*) h; }

// Cell centres as src/interaction.cl:65-67 and src/moment.cl:27 compute them, and dimensions.x (the MAC's extent, :71).
void clref_node_geometry(void* h, float* centre_dim4) {
	State& s = *(State*) h;
	for (std::size_t i = 0; i < s.nodes.size(); ++i) {
		const float4 c = s.nodes[i].position + s.nodes[i].dimensions / 2;
		centre_dim4[4 * i] = c.x; centre_dim4[4 * i + 1] = c.y; centre_dim4[4 * i + 2] = c.z; centre_dim4[4 * i + 3] = s.nodes[i].dimensions.x;
	}
}

// ---- moments: computeOctreeBuffers, src/open_cl_simulation.cpp:126-174 -------------------------------------------------
// returns the number of upsweep launches
std::uint32_t clref_compute_moments(void* h, int repair_d5) {
	State& s = *(State*) h;
	const index_t nn = (index_t) s.nodes.size(), nl = (index_t) s.leafs.size();
	// the kernel looks 7 entries past the one it examines (src/moment.cl:111-113): room for that behind both lists
	std::vector<index_t> processed(nn + 8, 0), fresh(nn + 8, 0);
	launch_1d(groups_for(nn, kLocal), kLocal, [&] { compute_moments_from_leafs(nl, s.leafs.data(), nn, s.nodes.data(), fresh.data()); });  // :145
	std::uint32_t launches = 0;
	std::size_t num_processed = nl;  // :149
	std::size_t live = nn;           // entries of `fresh` that hold data
	while (num_processed != 0) {     // :150
		// :152-163 — "remove zeros from the processed nodes and collapse the remaining entries to the front"
		num_processed = 0;
		const std::size_t bound = repair_d5 ? live : num_processed;  // as written the bound is the counter zeroed one line above (D5)
		for (std::size_t i = 0; i < bound; ++i)
			if (fresh[i] != 0) { fresh[num_processed] = fresh[i]; ++num_processed; }
		for (std::size_t i = num_processed; i < fresh.size(); ++i) fresh[i] = 0;  // (resize: nothing behind the list)
		std::copy(fresh.begin(), fresh.end(), processed.begin());                 // :165-167
		live = num_processed;
		const std::size_t items = num_processed / 8 + (num_processed % 8 != 0);    // :810-812, numNodesToScan = 8
		launch_1d(groups_for(items, kLocal), kLocal, [&] {
			compute_moments_from_nodes(nn, s.nodes.data(), (index_t) num_processed, processed.data(), fresh.data(), 8);
		});
		++launches;
		if (launches > 64) break;  // (a tree has at most 22 levels)
	}
	return launches;
}
void clref_get_moments(void* h, float* charge, float* dipole4, float* qcross4, float* qtrace4) {
	State& s = *(State*) h;
	for (std::size_t i = 0; i < s.nodes.size(); ++i) {
		const node_moment_t& m = s.nodes[i].value.moment;
		charge[i] = m.charge;
		std::memcpy(dipole4 + 4 * i, &m.dipole_moment, 16);
		std::memcpy(qcross4 + 4 * i, &m.quadrupole_cross_terms, 16);
		std::memcpy(qtrace4 + 4 * i, &m.quadrupole_trace_terms, 16);
	}
}

// ---- traversal: computeInteractionBuffers, src/open_cl_simulation.cpp:176-266 ------------------------------------------
void clref_traverse(void* h, std::uint64_t* n_node, std::uint64_t* n_leaf, std::uint64_t* rounds) {
	State& s = *(State*) h;
	s.leaf_inter.clear(); s.node_inter.clear(); s.rounds = 0;
	const index_t nn = (index_t) s.nodes.size();
	std::vector<interaction_t> pending(1), fresh;
	std::memset(pending.data(), 0, sizeof(interaction_t));  // the seed {0, 0}: include/nbody/open_cl_simulation.h:127-133
	while (!pending.empty()) {                             // step()'s do-while, :79-98
		++s.rounds;
		fresh.assign(64 * pending.size(), interaction_t{});  // newInteractions.zero(), :233
		launch_8x8(pending.size(), [&] { find_interactions(nn, s.nodes.data(), (index_t) pending.size(), pending.data(), fresh.data()); });
		std::vector<interaction_t> next;
		for (const interaction_t& x : fresh) {               // :242-266
			if (x.node_a_index == 0 && x.node_b_index == 0) {}
			else if (x.can_reduce) next.push_back(x);
			else if (!x.can_approx) s.leaf_inter.push_back(x);
			else s.node_inter.push_back(x);
		}
		pending.swap(next);
	}
	*n_node = s.node_inter.size(); *n_leaf = s.leaf_inter.size(); *rounds = s.rounds;
}
void clref_get_lists(void* h, std::uint32_t* node_pairs, std::uint32_t* leaf_pairs) {
	State& s = *(State*) h;
	for (std::size_t i = 0; i < s.node_inter.size(); ++i) { node_pairs[2 * i] = s.node_inter[i].node_a_index; node_pairs[2 * i + 1] = s.node_inter[i].node_b_index; }
	for (std::size_t i = 0; i < s.leaf_inter.size(); ++i) { leaf_pairs[2 * i] = s.leaf_inter[i].node_a_index; leaf_pairs[2 * i + 1] = s.leaf_inter[i].node_b_index; }
}

// ---- forces: rest of computeInteractionBuffers (:341-369) + computeForceBuffers (:486-570) ------------------------------
// leaf_force4 / node_force4: per leaf, the two force_t arrays the integration adds (:589-590).
// node_local_size: work-group size of compute_node_interaction_fields (the reference uses the device's preferred multiple).
void clref_forces(void* h, int repair_d7, std::uint32_t node_local_size, float* leaf_force4, float* node_force4) {
	State& s = *(State*) h;
	const index_t nn = (index_t) s.nodes.size(), nl = (index_t) s.leafs.size();
	std::vector<interaction_t> leaf_inter = s.leaf_inter, node_inter = s.node_inter;
	std::vector<index_t> num_leaf_inter(nn, 0), num_node_inter(nn, 0), max_leaf_count(nn, 0);  // :226-228
	auto indices = [&](std::vector<interaction_t>& list, std::vector<index_t>& counter) {         // :343-350
		launch_1d(groups_for(list.size(), kLocal), kLocal, [&] {
			compute_interaction_indices(nn, counter.data(), (index_t) list.size(), list.data());
		});
		if (repair_d7) {  // what the kernel computes and drops: the values its two atomic_inc calls return, in work-item order
			std::vector<index_t> replay(nn, 0);
			for (interaction_t& x : list) {
				x.node_a_interaction_index = replay[x.node_a_index]++;
				x.node_b_interaction_index = replay[x.node_b_index]++;
			}
		}
	};
	indices(leaf_inter, num_leaf_inter);
	indices(node_inter, num_node_inter);
	launch_1d(groups_for(leaf_inter.size(), kLocal), kLocal, [&] {                                // :353-357
		compute_node_max_interactions_leaf_count(nn, s.nodes.data(), max_leaf_count.data(), (index_t) leaf_inter.size(), leaf_inter.data());
	});
	// computeLeafFieldIndices, :371-417
	std::vector<index_t> leaf_field_idx(nl + 1, 0), node_field_idx(nl + 1, 0), parent_inter(nn, 0);
	for (index_t i = 0; i < nn; ++i) {
		const node_t& n = s.nodes[i];
		if (n.has_children) continue;
		const index_t fields = max_leaf_count[i] * num_leaf_inter[i];
		for (index_t l = n.leaf_index; l < n.leaf_index + n.leaf_count; ++l) leaf_field_idx[l + 1] = leaf_field_idx[l] + fields;
	}
	// computeNodeFieldIndices, :419-484
	for (index_t i = 0; i < nn; ++i) {
		const node_t& n = s.nodes[i];
		if (!n.has_children) continue;
		for (index_t k = 0; k < 8; ++k) parent_inter[i + n.child_indices[k]] += num_node_inter[i] + parent_inter[i];
	}
	for (index_t i = 0; i < nn; ++i) {
		const node_t& n = s.nodes[i];
		if (n.has_children) continue;
		const index_t fields = num_node_inter[i] + parent_inter[i];
		for (index_t l = n.leaf_index; l < n.leaf_index + n.leaf_count; ++l) node_field_idx[l + 1] = node_field_idx[l] + fields;
	}
	std::vector<leaf_field_t> leaf_fields(leaf_field_idx[nl]);
	std::vector<node_field_t> node_fields(node_field_idx[nl]);
	if (!leaf_fields.empty()) std::memset(leaf_fields.data(), 0, sizeof(leaf_field_t) * leaf_fields.size());  // :524-525
	if (!node_fields.empty()) std::memset(node_fields.data(), 0, sizeof(node_field_t) * node_fields.size());
	std::vector<force_t> leaf_forces(nl), node_forces(nl);
	std::memset(leaf_forces.data(), 0, sizeof(force_t) * nl);
	std::memset(node_forces.data(), 0, sizeof(force_t) * nl);
	launch_8x8(leaf_inter.size() + (leaf_inter.empty() ? 1 : 0), [&] {                             // :542-548, :922-930
		compute_leaf_interaction_fields(nl, s.leafs.data(), leaf_field_idx.data(), nn, s.nodes.data(), max_leaf_count.data(),
		                                (index_t) leaf_inter.size(), leaf_inter.data(), (index_t) leaf_fields.size(), leaf_fields.data());
	});
	launch_1d(2 * (node_inter.size() + (node_inter.empty() ? 1 : 0)), node_local_size, [&] {       // :549-554, :951-957
		compute_node_interaction_fields(nl + 1, node_field_idx.data(), nn, s.nodes.data(), parent_inter.data(), (index_t) node_inter.size(),
		                                node_inter.data(), (index_t) node_fields.size(), node_fields.data());
	});
	launch_1d(groups_for(nl, kLocal), kLocal, [&] {                                                // :557-561
		convert_leaf_fields_to_forces(nl, s.leafs.data(), leaf_field_idx.data(), leaf_forces.data(), (index_t) leaf_fields.size(), leaf_fields.data());
	});
	launch_1d(groups_for(nl, kLocal), kLocal, [&] {                                                // :562-566
		convert_node_fields_to_forces(nl, s.leafs.data(), node_field_idx.data(), node_forces.data(), (index_t) node_fields.size(), node_fields.data());
	});
	std::memcpy(leaf_force4, leaf_forces.data(), sizeof(force_t) * nl);
	std::memcpy(node_force4, node_forces.data(), sizeof(force_t) * nl);
}

// computeIntegrationBuffers, src/open_cl_simulation.cpp:594-609: v_new = v + F/m dt, x_new = x + v_OLD dt.
// Updates the harness's leaf records in place (the re-bucketing that follows in the reference is glade's `move`, absent).
void clref_integrate(void* h, const float* leaf_force4, const float* node_force4, float dt, float* particles12_out) {
	State& s = *(State*) h;
	for (std::size_t l = 0; l < s.leafs.size(); ++l) {
		leaf_t& leaf = s.leafs[l];
		float* pos = &leaf.position.x;
		float* vel = &leaf.value.velocity.x;
		for (unsigned i = 0; i < 3; ++i) {
			const float force = leaf_force4[4 * l + i] + node_force4[4 * l + i];
			const float old_velocity = vel[i];
			vel[i] += force / leaf.value.mass * dt;
			pos[i] += old_velocity * dt;
		}
		if (particles12_out) {
			float* r = particles12_out + 12 * l;
			std::memset(r, 0, 48);
			std::memcpy(r, pos, 12); std::memcpy(r + 4, vel, 12);
			r[8] = leaf.value.mass; r[9] = leaf.value.moment.charge;
		}
	}
}

// One force evaluation + integration with the kernels above, in one call (no Python between the stages): what an OpenCL CPU device
// would execute for OpenClSimulation::step() (src/open_cl_simulation.cpp:70-106) once the octree exists, on ONE host thread.
// Returns the number of field slots the reference's design allocates (leaf + node), for the record.
std::uint64_t clref_step(void* h, int repair, std::uint32_t node_local_size, float dt, float* particles12_out);

// find_interactions' verdict on arbitrary pairs of cells: pair k = two childless, non-empty nodes with lower corner pos_*4[4k..4k+2]
// and edge dim_*[k]; one launch over the n interactions (A_k, B_k); out[k] = can_approx (src/interaction.cl:76-82).
// Lets a test put the product's acceptance test next to the reference's kernel on geometry no octree would produce.
void clref_can_approx(std::uint64_t n, const float* pos_a4, const float* dim_a, const float* pos_b4, const float* dim_b, std::uint8_t* out) {
	std::vector<node_t> nodes(2 * n + 1);
	std::memset(nodes.data(), 0, sizeof(node_t) * nodes.size());
	std::vector<interaction_t> pending(n), fresh(64 * n, interaction_t{});
	for (std::uint64_t k = 0; k < n; ++k) {
		node_t& a = nodes[1 + 2 * k];
		node_t& b = nodes[2 + 2 * k];
		a.position = float4(pos_a4[4 * k], pos_a4[4 * k + 1], pos_a4[4 * k + 2], 0.0f); a.dimensions = float4(dim_a[k], dim_a[k], dim_a[k], 0.0f);
		b.position = float4(pos_b4[4 * k], pos_b4[4 * k + 1], pos_b4[4 * k + 2], 0.0f); b.dimensions = float4(dim_b[k], dim_b[k], dim_b[k], 0.0f);
		a.leaf_count = b.leaf_count = 1;
		pending[k] = interaction_t{(index_t) (1 + 2 * k), (index_t) (2 + 2 * k), 0, 0, 0, 0};
	}
	launch_8x8(n, [&] { find_interactions((index_t) nodes.size(), nodes.data(), (index_t) n, pending.data(), fresh.data()); });
	for (std::uint64_t k = 0; k < n; ++k) out[k] = fresh[64 * k].can_approx;
}

// src/field.cl:17-32 and src/force.cl:4-10 on one pair: the FORCE (charge x field) on a and on b.
void clref_pair_force(float qa, float qb, const float* pa3, const float* pb3, float* force_a3, float* force_b3) {
	leaf_moment_t ma{qa}, mb{qb};
	const float4 a(pa3[0], pa3[1], pa3[2], 0.0f), b(pb3[0], pb3[1], pb3[2], 0.0f);
	const leaf_field_pair_t f = leaf_moment_field(ma, mb, a, b);
	const force_t fa = leaf_field_to_force(ma, f.field_a, a), fb = leaf_field_to_force(mb, f.field_b, b);
	force_a3[0] = fa.force.x; force_a3[1] = fa.force.y; force_a3[2] = fa.force.z;
	force_b3[0] = fb.force.x; force_b3[1] = fb.force.y; force_b3[2] = fb.force.z;
}

std::uint64_t clref_step(void* h, int repair, std::uint32_t node_local_size, float dt, float* particles12_out) {
	State& s = *(State*) h;
	clref_compute_moments(h, repair);
	std::uint64_t a, b, r;
	clref_traverse(h, &a, &b, &r);
	std::vector<float> lf(4 * s.leafs.size()), nf(4 * s.leafs.size());
	clref_forces(h, repair, node_local_size, lf.data(), nf.data());
	clref_integrate(h, lf.data(), nf.data(), dt, particles12_out);
	return a + b;
}

}  // extern "C"
