// ============================================================================
// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's algorithm for the hot path
// (duanebyer/nbody; citations are relative to /root/reference). Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this library. The product (libnbody_cuda.so) never links it and
// has no CPU fallback.
//
// Parity status:
//  * naive direct sum  — PINNED: orc_naive_step_as_written() is checked bit
//    for bit against the unmodified reference source compiled into
//    oracle/_ref/libnaive_ref.so (tests/test_oracle.py) and against the
//    golden vectors generated from it (tests/golden/).
//  * octree topology   — "parity unpinned": the reference's octree is the
//    un-vendored, un-versioned glade::Orthtree (CMakeLists.txt:33). What is
//    restated here is the contract inferred from how the reference CONSUMES
//    node_t (include/nbody/device/types.h:124-141 and the sites listed at each
//    function below).
//  * traversal / MAC   — PINNED on a given octree: orc_traverse() restates
//    src/interaction.cl:22-99 and the host partition loop
//    src/open_cl_simulation.cpp:247-266 (FP32, no FMA contraction) and is
//    checked entry for entry, in order, against the reference's own
//    find_interactions kernel compiled for the host from the file where it lies
//    (oracle/_ref/libclref.so, oracle/ref_cl_harness.cpp) and against the golden
//    lists generated from it (tests/test_reference_kernels.py,
//    tests/golden/clref_golden.npz). One deliberate difference: a childless
//    root keeps its self pair (see orc_traverse).
//  * expansions        — orc_fmm_field() at order 1 IS the reference's far field
//    (monopole at the target cell centre, src/field.cl:35-47,187-210) and at
//    order 2 carries the reference's leaf moments (src/moment.cl:27-54): both
//    checked against the reference's kernels to FP32 round-off (same test file).
//    Orders 2..5 continue that scheme; the reference has nothing to compare.
//  * accelerations     — FP64 direct sum with the reference's softening
//    (src/field.cl:22-24, pair term checked against leaf_moment_field) and the
//    naive loop's pair structure (src/naive_simulation.cpp:7-46, with the x-only
//    delta defect D1 fixed).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off).
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------
// Morton keys (new definition, SURVEY 3.2 "Morton key definition"): 21 bits per
// dimension, digit = x | y<<1 | z<<2 (z most significant inside each 3-bit
// digit), level-1 digit in the top bits of a 63-bit key.
// ----------------------------------------------------------------------------
constexpr int kMaxDepth = 21;

inline std::uint64_t spread3(std::uint32_t v) {
	std::uint64_t x = v & 0x1fffffu;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

inline std::uint32_t quantise(float x, float scale) {
	// one FP32 multiply, clamp, truncate (device: __fmul_rn, fminf/fmaxf, __float2uint_rz)
	float v = x * scale;
	v = std::fmin(std::fmax(v, 0.0f), 2097151.0f);
	return (std::uint32_t) v;
}

struct Tree {
	std::uint32_t capacity = 8, max_depth = kMaxDepth;
	float bounds[3] = {1, 1, 1};
	// DFS pre-order node arrays (the reference's node_t fields,
	// include/nbody/device/types.h:124-141)
	std::vector<std::uint32_t> depth, leaf_index, leaf_count, sibling;
	std::vector<std::uint64_t> prefix;  // key prefix: the top 3*depth bits of the 63-bit key, low bits zero
	std::vector<std::uint8_t> has_children;
	std::vector<std::uint32_t> child_off;  // 9 per node (8 children + end of subtree), relative
	std::vector<std::int32_t> parent_off;  // relative, 0 for root
	std::vector<float> geom;               // 4 per node: centre x,y,z and dimensions.x
	// traversal output: unordered pairs in DFS ids
	std::vector<std::uint32_t> m2l, p2p;
	std::uint64_t rounds = 0, mac_tests = 0;
};

// glade::Orthtree bulk build as consumed at src/open_cl_simulation.cpp:41-47
// (capacity 8, all 8 children exist once a node splits, DFS pre-order storage,
// leaf ranges for every node). Split rule: more than `capacity` leaves and
// depth < max_depth (the reference's FIXME at :280-283 notes nodes can exceed
// capacity; here that only happens at max_depth).
std::uint32_t build_rec(Tree& t, const std::uint64_t* keys, std::uint32_t begin, std::uint32_t end,
                        std::uint32_t depth, std::uint64_t prefix, std::int32_t parent, std::uint32_t sib) {
	const std::uint32_t id = (std::uint32_t) t.depth.size();
	t.depth.push_back(depth);
	t.prefix.push_back(prefix);
	t.leaf_index.push_back(begin);
	t.leaf_count.push_back(end - begin);
	t.sibling.push_back(sib);
	t.has_children.push_back(0);
	t.parent_off.push_back(parent < 0 ? 0 : parent - (std::int32_t) id);
	for (int k = 0; k < 9; ++k) t.child_off.push_back(0);
	// geometry: position = ix * dim, centre = position + dim/2, per axis, FP32,
	// one rounding per op (src/interaction.cl:65-67 computes the same centre).
	std::uint32_t ix = 0, iy = 0, iz = 0;
	for (std::uint32_t l = 0; l < depth; ++l) {
		const std::uint32_t d = (std::uint32_t) (prefix >> (3 * (kMaxDepth - 1 - l))) & 7u;
		ix = ix << 1 | (d & 1u); iy = iy << 1 | (d >> 1 & 1u); iz = iz << 1 | (d >> 2 & 1u);
	}
	const float sc = std::ldexp(1.0f, -(int) depth);
	const float dx = t.bounds[0] * sc, dy = t.bounds[1] * sc, dz = t.bounds[2] * sc;
	t.geom.push_back((float) ix * dx + dx * 0.5f);
	t.geom.push_back((float) iy * dy + dy * 0.5f);
	t.geom.push_back((float) iz * dz + dz * 0.5f);
	t.geom.push_back(dx);
	if (end - begin > t.capacity && depth < t.max_depth) {
		t.has_children[id] = 1;
		const int shift = 3 * (kMaxDepth - 1 - (int) depth);
		std::uint32_t b = begin;
		for (std::uint32_t k = 0; k < 8; ++k) {
			std::uint32_t e = b;
			while (e < end && ((keys[e] >> shift) & 7u) == k) ++e;
			const std::uint32_t cid = build_rec(t, keys, b, e, depth + 1, prefix | (std::uint64_t) k << shift, (std::int32_t) id, k);
			t.child_off[9 * (std::size_t) id + k] = cid - id;
			b = e;
		}
	}
	t.child_off[9 * (std::size_t) id + 8] = (std::uint32_t) t.depth.size() - id;
	return id;
}

// ----------------------------------------------------------------------------
// MAC, src/interaction.cl:64-82. FP32, IEEE ops, evaluation order fixed here and
// mirrored with __fmul_rn/__fadd_rn/__fdiv_rn on the device.
// ----------------------------------------------------------------------------
inline bool mac_accept(const float* ga, const float* gb, float ratio_sq) {
	const float dx = gb[0] - ga[0], dy = gb[1] - ga[1], dz = gb[2] - ga[2];
	const float d2 = (dx * dx + dy * dy) + dz * dz;
	const float ext = ga[3] + gb[3];
	const float ext2 = (0.75f * ext) * ext;
	return ext2 / d2 < ratio_sq;
}

// Multi-index tables for Cartesian Taylor expansions of total order <= p.
struct Idx {
	int p, n;
	std::vector<int> ex, ey, ez, ord;
	std::vector<double> fact;  // i! j! k!
	std::vector<int> lut;      // (i,j,k) -> index, (p+1)^3
	explicit Idx(int p_) : p(p_) {
		lut.assign((p + 1) * (p + 1) * (p + 1), -1);
		n = 0;
		static const double f[] = {1, 1, 2, 6, 24, 120, 720, 5040, 40320, 362880, 3628800};
		for (int o = 0; o <= p; ++o)
			for (int i = o; i >= 0; --i)
				for (int j = o - i; j >= 0; --j) {
					const int k = o - i - j;
					ex.push_back(i); ey.push_back(j); ez.push_back(k); ord.push_back(o);
					fact.push_back(f[i] * f[j] * f[k]);
					lut[(i * (p + 1) + j) * (p + 1) + k] = n++;
				}
	}
	int at(int i, int j, int k) const { return lut[(i * (p + 1) + j) * (p + 1) + k]; }
};

// Taylor coefficients a_k = (1/k!) d^k/dx^k [ (|x|^2 + eps^2)^(-1/2) ] by the
// three-term recurrence  |k| R^2 a_k + (2|k|-1) sum_i x_i a_{k-e_i} + (|k|-1) sum_i a_{k-2e_i} = 0.
void kernel_taylor(const Idx& I, const double x[3], double eps2, double* a) {
	const double R2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + eps2;
	a[0] = 1.0 / std::sqrt(R2);
	for (int q = 1; q < I.n; ++q) {
		const int e[3] = {I.ex[q], I.ey[q], I.ez[q]};
		const int o = I.ord[q];
		double s1 = 0, s2 = 0;
		for (int d = 0; d < 3; ++d) {
			if (e[d] >= 1) { int f[3] = {e[0], e[1], e[2]}; f[d] -= 1; s1 += x[d] * a[I.at(f[0], f[1], f[2])]; }
			if (e[d] >= 2) { int f[3] = {e[0], e[1], e[2]}; f[d] -= 2; s2 += a[I.at(f[0], f[1], f[2])]; }
		}
		a[q] = -((2.0 * o - 1.0) * s1 + (o - 1.0) * s2) / (o * R2);
	}
}

inline double ipow(double b, int e) { double r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

void parallel_for(std::uint64_t n, int threads, const std::function<void(std::uint64_t, std::uint64_t)>& fn) {
	if (threads <= 1 || n < 2) { fn(0, n); return; }
	std::vector<std::thread> th;
	const std::uint64_t chunk = (n + threads - 1) / threads;
	for (int t = 0; t < threads; ++t) {
		const std::uint64_t b = std::min<std::uint64_t>(n, t * chunk), e = std::min<std::uint64_t>(n, b + chunk);
		if (b < e) th.emplace_back(fn, b, e);
	}
	for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

// ---- keys / sort -----------------------------------------------------------
// pos: n records of `stride` floats (x,y,z first). bounds: 3 floats.
void orc_morton_keys(std::uint64_t n, const float* pos, std::uint32_t stride, const float* bounds, std::uint64_t* keys) {
	const float sx = 2097152.0f / bounds[0], sy = 2097152.0f / bounds[1], sz = 2097152.0f / bounds[2];
	for (std::uint64_t i = 0; i < n; ++i) {
		const float* p = pos + (std::size_t) stride * i;
		keys[i] = spread3(quantise(p[0], sx)) | spread3(quantise(p[1], sy)) << 1 | spread3(quantise(p[2], sz)) << 2;
	}
}

// Stable ascending sort; perm[i] = index into the input of the i-th sorted key.
void orc_sort_keys(std::uint64_t n, const std::uint64_t* keys, std::uint64_t* sorted, std::uint32_t* perm) {
	std::vector<std::uint32_t> p(n);
	for (std::uint64_t i = 0; i < n; ++i) p[i] = (std::uint32_t) i;
	std::stable_sort(p.begin(), p.end(), [&](std::uint32_t a, std::uint32_t b) { return keys[a] < keys[b]; });
	for (std::uint64_t i = 0; i < n; ++i) { perm[i] = p[i]; sorted[i] = keys[p[i]]; }
}

// ---- tree ------------------------------------------------------------------
void* orc_tree_build(std::uint64_t n, const std::uint64_t* sorted_keys, const float* bounds,
                     std::uint32_t capacity, std::uint32_t max_depth) {
	Tree* t = new Tree;
	t->capacity = capacity;
	t->max_depth = std::min<std::uint32_t>(max_depth, kMaxDepth);
	for (int k = 0; k < 3; ++k) t->bounds[k] = bounds[k];
	build_rec(*t, sorted_keys, 0, (std::uint32_t) n, 0, 0, -1, 0);
	return t;
}
void orc_tree_free(void* h) { delete (Tree*) h; }
std::uint32_t orc_tree_num_nodes(void* h) { return (std::uint32_t) ((Tree*) h)->depth.size(); }
void orc_tree_get(void* h, std::uint32_t* depth, std::uint64_t* prefix, std::uint32_t* leaf_index,
                  std::uint32_t* leaf_count, std::uint8_t* has_children, std::uint32_t* child_off9,
                  std::int32_t* parent_off, std::uint32_t* sibling, float* geom4) {
	Tree& t = *(Tree*) h;
	const std::size_t m = t.depth.size();
	if (depth) std::memcpy(depth, t.depth.data(), 4 * m);
	if (prefix) std::memcpy(prefix, t.prefix.data(), 8 * m);
	if (leaf_index) std::memcpy(leaf_index, t.leaf_index.data(), 4 * m);
	if (leaf_count) std::memcpy(leaf_count, t.leaf_count.data(), 4 * m);
	if (has_children) std::memcpy(has_children, t.has_children.data(), m);
	if (child_off9) std::memcpy(child_off9, t.child_off.data(), 36 * m);
	if (parent_off) std::memcpy(parent_off, t.parent_off.data(), 4 * m);
	if (sibling) std::memcpy(sibling, t.sibling.data(), 4 * m);
	if (geom4) std::memcpy(geom4, t.geom.data(), 16 * m);
}

// ---- traversal -------------------------------------------------------------
// Rounds of find_interactions (src/interaction.cl:10-100) + the host partition
// (src/open_cl_simulation.cpp:247-266), seeded with {root, root}
// (include/nbody/open_cl_simulation.h:127-133). Output pairs are unordered and
// appear once; both directions are evaluated from each (src/field.cl:25-30).
void orc_traverse(void* h, float mac_ratio, std::uint64_t* n_m2l, std::uint64_t* n_p2p) {
	Tree& t = *(Tree*) h;
	t.m2l.clear(); t.p2p.clear(); t.rounds = 0; t.mac_tests = 0;
	const float ratio_sq = mac_ratio * mac_ratio;
	std::vector<std::uint32_t> cur, next;
	// The reference seeds the reducible list with {0,0} unconditionally; a
	// childless root then yields exactly one leaf interaction (root, root).
	cur.push_back(0); cur.push_back(0);
	while (!cur.empty()) {
		++t.rounds;
		next.clear();
		for (std::size_t q = 0; q < cur.size(); q += 2) {
			const std::uint32_t A = cur[q], B = cur[q + 1];
			// output slot order 64*i + 8*lid_b + lid_a  => lid_b outer, lid_a inner
			for (std::uint32_t lb = 0; lb < 8; ++lb)
				for (std::uint32_t la = 0; la < 8; ++la) {
					const std::uint32_t ca = t.has_children[A] ? A + t.child_off[9 * (std::size_t) A + la] : A;
					const std::uint32_t cb = t.has_children[B] ? B + t.child_off[9 * (std::size_t) B + lb] : B;
					if (t.leaf_count[ca] == 0 || t.leaf_count[cb] == 0 || (A == B && lb > la) ||
					    (!t.has_children[A] && la != 0) || (!t.has_children[B] && lb != 0))
						continue;
					bool can_approx = false;
					if (ca != cb) { ++t.mac_tests; can_approx = mac_accept(&t.geom[4 * (std::size_t) ca], &t.geom[4 * (std::size_t) cb], ratio_sq); }
					const bool can_reduce = !can_approx && (t.has_children[ca] || t.has_children[cb]);
					// (the reference drops a produced {0,0}: only possible for a childless root,
					//  where it is the single self leaf interaction; keep it as P2P so a
					//  sub-capacity system still has forces)
					if (can_reduce) { next.push_back(ca); next.push_back(cb); }
					else if (!can_approx) { t.p2p.push_back(ca); t.p2p.push_back(cb); }
					else { t.m2l.push_back(ca); t.m2l.push_back(cb); }
				}
		}
		cur.swap(next);
	}
	*n_m2l = t.m2l.size() / 2;
	*n_p2p = t.p2p.size() / 2;
}
void orc_get_lists(void* h, std::uint32_t* m2l_pairs, std::uint32_t* p2p_pairs) {
	Tree& t = *(Tree*) h;
	if (m2l_pairs) std::memcpy(m2l_pairs, t.m2l.data(), 4 * t.m2l.size());
	if (p2p_pairs) std::memcpy(p2p_pairs, t.p2p.data(), 4 * t.p2p.size());
}
std::uint64_t orc_traverse_rounds(void* h) { return ((Tree*) h)->rounds; }

// ---- FP64 direct sum ---------------------------------------------------------
// field g_i = sum_{j != i} q_j (x_j - x_i) / (|x_j - x_i|^2 + eps^2)^{3/2}
// (src/naive_simulation.cpp:7-46 with D1 fixed, softened per src/field.cl:22-24;
// acceleration a_i = G q_i/m_i g_i with the naive sign convention, G>0 attracts).
// posq: n_src records (x,y,z,q). targets: indices into posq (or NULL = 0..n_tgt-1).
void orc_direct_field(std::uint64_t n_src, const float* posq, std::uint64_t n_tgt, const std::uint32_t* targets,
                      double eps, double* g3, double* phi, int threads) {
	const double eps2 = eps * eps;
	parallel_for(n_tgt, threads, [&](std::uint64_t b, std::uint64_t e) {
		for (std::uint64_t ti = b; ti < e; ++ti) {
			const std::uint64_t i = targets ? targets[ti] : ti;
			const double xi = posq[4 * i], yi = posq[4 * i + 1], zi = posq[4 * i + 2];
			double gx = 0, gy = 0, gz = 0, ph = 0;
			for (std::uint64_t j = 0; j < n_src; ++j) {
				if (j == i) continue;
				const double dx = posq[4 * j] - xi, dy = posq[4 * j + 1] - yi, dz = posq[4 * j + 2] - zi;
				const double r2 = dx * dx + dy * dy + dz * dz + eps2;
				if (r2 == 0.0) continue;
				const double inv = 1.0 / std::sqrt(r2), w = posq[4 * j + 3] * inv * inv * inv;
				gx += w * dx; gy += w * dy; gz += w * dz; ph += posq[4 * j + 3] * inv;
			}
			g3[3 * ti] = gx; g3[3 * ti + 1] = gy; g3[3 * ti + 2] = gz;
			if (phi) phi[ti] = ph;
		}
	});
}

// ---- FP64 FMM over the reference-rule lists ---------------------------------
// posq is in SORTED (tree/leaf) order. Expansion centre = geometric cell centre
// (src/moment.cl:27). order p >= 0 (p = 0 reproduces the reference's monopole
// far field up to where it is evaluated). Outputs: g3 (field per particle),
// optional multipoles/locals (ncoef per node, DFS order) for kernel-level parity.
// Conventions: M_m = sum q (y-c)^m / m!;  Phi(x) = sum_n L_n (x-c)^n;
// field g = grad Phi with kernel 1/sqrt(r^2+eps^2).
std::uint32_t orc_ncoef(std::uint32_t p) { return (p + 1) * (p + 2) * (p + 3) / 6; }

// Multipole expansion centre: 0 = geometric cell centre (src/moment.cl:27),
// 1 = centre of |charge| of the node's particles (exploration / product option).
static int g_centre_mode = 0;
void orc_set_centre_mode(int mode) { g_centre_mode = mode; }
// Exploration: M2L pairs with ext2/d2 < g_low_tau are evaluated at order p-1 (0 = off).
static double g_low_tau = 0.0;
static unsigned long long g_low_count = 0, g_all_count = 0;
void orc_set_low_order_tau(double tau) { g_low_tau = tau; g_low_count = g_all_count = 0; }
unsigned long long orc_low_count(void) { return g_low_count; }
unsigned long long orc_all_count(void) { return g_all_count; }

void orc_fmm_field(void* h, std::uint64_t n, const float* posq, std::uint32_t order, double eps,
                   double* g3, double* phi, double* multipoles, double* locals, int threads) {
	Tree& t = *(Tree*) h;
	const Idx I((int) order);
	const int nc = I.n, p = (int) order;
	const std::size_t m = t.depth.size();
	const double eps2 = eps * eps;
	std::vector<double> M(m * nc, 0.0), L(m * nc, 0.0);
	auto centre = [&](std::size_t id, double c[3]) { for (int d = 0; d < 3; ++d) c[d] = t.geom[4 * id + d]; };
	// multipole centres
	std::vector<double> mc(3 * m);
	for (std::size_t id = 0; id < m; ++id) {
		centre(id, &mc[3 * id]);
		if (g_centre_mode == 1 && t.leaf_count[id] > 0) {
			double w = 0, s3[3] = {0, 0, 0};
			for (std::uint32_t q = t.leaf_index[id]; q < t.leaf_index[id] + t.leaf_count[id]; ++q) {
				const double a_ = std::fabs((double) posq[4 * q + 3]);
				w += a_; for (int d = 0; d < 3; ++d) s3[d] += a_ * posq[4 * q + d];
			}
			if (w > 0) for (int d = 0; d < 3; ++d) mc[3 * id + d] = s3[d] / w;
		}
	}
	auto mcentre = [&](std::size_t id, double c[3]) { for (int d = 0; d < 3; ++d) c[d] = mc[3 * id + d]; };
	// P2M on childless nodes (src/moment.cl:6-67), then M2M deepest-first. DFS
	// pre-order => children have larger ids than parents: sweep ids descending.
	for (std::size_t id = m; id-- > 0;) {
		double c[3]; mcentre(id, c);
		double* Mi = &M[id * nc];
		if (!t.has_children[id]) {
			for (std::uint32_t q = t.leaf_index[id]; q < t.leaf_index[id] + t.leaf_count[id]; ++q) {
				const double r[3] = {posq[4 * q] - c[0], posq[4 * q + 1] - c[1], posq[4 * q + 2] - c[2]};
				for (int a = 0; a < nc; ++a)
					Mi[a] += posq[4 * q + 3] * ipow(r[0], I.ex[a]) * ipow(r[1], I.ey[a]) * ipow(r[2], I.ez[a]) / I.fact[a];
			}
		} else {
			for (int k = 0; k < 8; ++k) {
				const std::size_t cid = id + t.child_off[9 * id + k];
				if (t.leaf_count[cid] == 0) continue;
				double cc[3]; mcentre(cid, cc);
				const double d[3] = {cc[0] - c[0], cc[1] - c[1], cc[2] - c[2]};
				const double* Mc = &M[cid * nc];
				for (int a = 0; a < nc; ++a)      // target index n
					for (int b = 0; b < nc; ++b) {  // source index m <= n
						const int ei = I.ex[a] - I.ex[b], ej = I.ey[a] - I.ey[b], ek = I.ez[a] - I.ez[b];
						if (ei < 0 || ej < 0 || ek < 0) continue;
						static const double f[] = {1, 1, 2, 6, 24, 120, 720, 5040, 40320, 362880, 3628800};
						Mi[a] += Mc[b] * ipow(d[0], ei) * ipow(d[1], ej) * ipow(d[2], ek) / (f[ei] * f[ej] * f[ek]);
					}
			}
		}
	}
	// M2L over the unordered list, both directions.
	std::vector<double> a(nc);
	auto m2l = [&](std::size_t tgt, std::size_t src) {
		double ct[3], cs[3]; centre(tgt, ct); mcentre(src, cs);
		const double x[3] = {ct[0] - cs[0], ct[1] - cs[1], ct[2] - cs[2]};
		kernel_taylor(I, x, eps2, a.data());
		const double* Ms = &M[src * nc];
		double* Lt = &L[tgt * nc];
		int pe = p;
		++g_all_count;
		if (g_low_tau > 0 && p >= 3) {
			// FP32, same operation order as the device (traverse.cu mac_classify): ext2 < tau * d2
			const float* ga = &t.geom[4 * tgt];
			const float* gb = &t.geom[4 * src];
			const float fx = gb[0] - ga[0], fy = gb[1] - ga[1], fz = gb[2] - ga[2];
			const float d2 = (fx * fx + fy * fy) + fz * fz;
			const float ext = ga[3] + gb[3];
			const float ext2 = (0.75f * ext) * ext;
			const float tau = (float) g_low_tau;
			if (ext2 < tau * d2) { pe = p - 1; ++g_low_count; }
		}
		for (int nn = 0; nn < nc; ++nn)
			for (int mm = 0; mm < nc; ++mm) {
				if (I.ord[nn] + I.ord[mm] > pe) continue;
				const int s = I.at(I.ex[nn] + I.ex[mm], I.ey[nn] + I.ey[mm], I.ez[nn] + I.ez[mm]);
				const double sign = (I.ord[mm] & 1) ? -1.0 : 1.0;
				// D_{n+m}/n! = (n+m)!/n! a_{n+m}
				Lt[nn] += sign * Ms[mm] * a[s] * I.fact[s] / I.fact[nn];
			}
	};
	for (std::size_t q = 0; q < t.m2l.size(); q += 2) { m2l(t.m2l[q], t.m2l[q + 1]); m2l(t.m2l[q + 1], t.m2l[q]); }
	if (multipoles) std::memcpy(multipoles, M.data(), 8 * M.size());
	// L2L parents-first (ascending DFS id).
	for (std::size_t id = 0; id < m; ++id) {
		if (!t.has_children[id]) continue;
		double c[3]; centre(id, c);
		const double* Lp = &L[id * nc];
		for (int k = 0; k < 8; ++k) {
			const std::size_t cid = id + t.child_off[9 * id + k];
			if (t.leaf_count[cid] == 0) continue;
			double cc[3]; centre(cid, cc);
			const double d[3] = {cc[0] - c[0], cc[1] - c[1], cc[2] - c[2]};
			double* Lc = &L[cid * nc];
			for (int nn = 0; nn < nc; ++nn)
				for (int mm = 0; mm < nc; ++mm) {
					const int ei = I.ex[mm] - I.ex[nn], ej = I.ey[mm] - I.ey[nn], ek = I.ez[mm] - I.ez[nn];
					if (ei < 0 || ej < 0 || ek < 0) continue;
					static const double f[] = {1, 1, 2, 6, 24, 120, 720, 5040, 40320, 362880, 3628800};
					// C(m,n) = m!/(n!(m-n)!)
					Lc[nn] += Lp[mm] * I.fact[mm] / (I.fact[nn] * f[ei] * f[ej] * f[ek]) * ipow(d[0], ei) * ipow(d[1], ej) * ipow(d[2], ek);
				}
		}
	}
	if (locals) std::memcpy(locals, L.data(), 8 * L.size());
	// L2P on childless nodes.
	for (std::uint64_t i = 0; i < 3 * n; ++i) g3[i] = 0.0;
	if (phi) for (std::uint64_t i = 0; i < n; ++i) phi[i] = 0.0;
	for (std::size_t id = 0; id < m; ++id) {
		if (t.has_children[id]) continue;
		double c[3]; centre(id, c);
		const double* Ll = &L[id * nc];
		for (std::uint32_t q = t.leaf_index[id]; q < t.leaf_index[id] + t.leaf_count[id]; ++q) {
			const double r[3] = {posq[4 * q] - c[0], posq[4 * q + 1] - c[1], posq[4 * q + 2] - c[2]};
			for (int nn = 0; nn < nc; ++nn) {
				const int e[3] = {I.ex[nn], I.ey[nn], I.ez[nn]};
				if (phi) phi[q] += Ll[nn] * ipow(r[0], e[0]) * ipow(r[1], e[1]) * ipow(r[2], e[2]);
				for (int d = 0; d < 3; ++d) {
					if (e[d] == 0) continue;
					int f[3] = {e[0], e[1], e[2]}; f[d] -= 1;
					g3[3 * q + d] += Ll[nn] * e[d] * ipow(r[0], f[0]) * ipow(r[1], f[1]) * ipow(r[2], f[2]);
				}
			}
		}
	}
	// P2P over the unordered leaf-pair list, both directions (src/field.cl:49-148:
	// self pairs take b < a only; here expressed as all ordered i != j).
	(void) threads;
	for (std::size_t q = 0; q < t.p2p.size(); q += 2) {
		const std::uint32_t A = t.p2p[q], B = t.p2p[q + 1];
		for (std::uint32_t i = t.leaf_index[A]; i < t.leaf_index[A] + t.leaf_count[A]; ++i)
			for (std::uint32_t j = t.leaf_index[B]; j < t.leaf_index[B] + t.leaf_count[B]; ++j) {
				if (A == B && j >= i) continue;
				const double dx = (double) posq[4 * j] - posq[4 * i], dy = (double) posq[4 * j + 1] - posq[4 * i + 1],
				             dz = (double) posq[4 * j + 2] - posq[4 * i + 2];
				const double r2 = dx * dx + dy * dy + dz * dz + eps2;
				if (r2 == 0.0) continue;
				const double inv = 1.0 / std::sqrt(r2), inv3 = inv * inv * inv;
				const double qi = posq[4 * i + 3], qj = posq[4 * j + 3];
				g3[3 * i] += qj * inv3 * dx; g3[3 * i + 1] += qj * inv3 * dy; g3[3 * i + 2] += qj * inv3 * dz;
				g3[3 * j] -= qi * inv3 * dx; g3[3 * j + 1] -= qi * inv3 * dy; g3[3 * j + 2] -= qi * inv3 * dz;
				if (phi) { phi[i] += qj * inv; phi[j] += qi * inv; }
			}
	}
}

// ---- the reference's naive step, as written ---------------------------------
// Bug-compatible FP32 restatement of src/naive_simulation.cpp:7-46 (x-only delta
// at :10-14, no softening at :16-17, kick-then-drift at :28-42, FP32 time at :44).
// particles: n records of 12 floats {pos[4], vel[4], mass, charge, pad, pad}.
float orc_naive_step_as_written(std::uint64_t n, float* P, float k, float dt, std::uint32_t steps) {
	float time = 0.0f;
	for (std::uint32_t s = 0; s < steps; ++s) {
		for (std::uint64_t i = 1; i < n; ++i)
			for (std::uint64_t j = 0; j < i; ++j) {
				float* a = P + 12 * i; float* b = P + 12 * j;
				const float d0 = b[0] - a[0], d1 = b[0] - a[0], d2 = b[0] - a[0];
				const float r = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
				const float cf = a[9] * b[9], ma = a[8], mb = b[8];
				const float f0 = k * cf * d0 / (r * r * r), f1 = k * cf * d1 / (r * r * r), f2 = k * cf * d2 / (r * r * r);
				a[4] += f0 / ma * dt; a[5] += f1 / ma * dt; a[6] += f2 / ma * dt;
				b[4] -= f0 / mb * dt; b[5] -= f1 / mb * dt; b[6] -= f2 / mb * dt;
			}
		for (std::uint64_t i = 0; i < n; ++i) {
			float* a = P + 12 * i;
			a[0] += a[4] * dt; a[1] += a[5] * dt; a[2] += a[6] * dt;
		}
		time += dt;
	}
	return time;
}

// Corrected + softened restatement in the same loop structure (FP32 state, FP64
// accumulation of the field): the semantics the product implements.
// integrator 0 = kick-drift (src/naive_simulation.cpp:28-42), 1 = explicit Euler
// with the old velocity (src/open_cl_simulation.cpp:602-607).
float orc_direct_step(std::uint64_t n, float* P, float G, float eps, float dt, std::uint32_t steps, int integrator, int threads) {
	float time = 0.0f;
	std::vector<float> posq(4 * n);
	std::vector<double> g(3 * n);
	for (std::uint32_t s = 0; s < steps; ++s) {
		for (std::uint64_t i = 0; i < n; ++i) { posq[4 * i] = P[12 * i]; posq[4 * i + 1] = P[12 * i + 1]; posq[4 * i + 2] = P[12 * i + 2]; posq[4 * i + 3] = P[12 * i + 9]; }
		orc_direct_field(n, posq.data(), n, nullptr, eps, g.data(), nullptr, threads);
		for (std::uint64_t i = 0; i < n; ++i) {
			float* a = P + 12 * i;
			const double s_ = (double) G * a[9] / a[8];
			for (int d = 0; d < 3; ++d) {
				const float v_old = a[4 + d];
				const float v_new = (float) (v_old + s_ * g[3 * i + d] * dt);
				a[4 + d] = v_new;
				a[d] = a[d] + (integrator == 0 ? v_new : v_old) * dt;
			}
		}
		time += dt;
	}
	return time;
}

// Variable time step. The reference has no implementation to follow: TODO:2-3 names "variable timestep" as a goal and
// every run uses the constant src/main.cpp:69. What is restated here is the rule documented in include/nbody_cuda.h
// (nbody_cuda_config::time_step_eta): step n uses dt_n, dt_0 = dt0, dt_{n+1} = clamp(eta sqrt(len / max_i |a_i|), lo, hi)
// with the accelerations step n computed (FP64 direct sum here; the product takes the maximum of its FP32 accelerations,
// so the two sequences agree to FP32 round-off, not bitwise). Returns the final time; dts[s] / amax[s] = the step taken
// at step s and the largest acceleration it found.
float orc_next_time_step(float eta, float len, float acc_max, float dt0, float dt_min, float dt_max) {
	if (!(eta > 0.0f) || !(acc_max > 0.0f) || !std::isfinite(acc_max)) return dt0;
	float dt = eta * std::sqrt(len / acc_max);
	const float hi = dt_max > 0.0f ? dt_max : dt0;
	if (!(dt < hi)) dt = hi;
	if (dt_min > 0.0f && dt < dt_min) dt = dt_min;
	return dt;
}

float orc_direct_step_adaptive(std::uint64_t n, float* P, float G, float eps, float len, float dt0, float eta, float dt_min, float dt_max,
                               std::uint32_t steps, int integrator, int threads, float* dts, float* amax) {
	float time = 0.0f, dt = dt0;
	std::vector<float> posq(4 * n);
	std::vector<double> g(3 * n);
	for (std::uint32_t s = 0; s < steps; ++s) {
		for (std::uint64_t i = 0; i < n; ++i) { posq[4 * i] = P[12 * i]; posq[4 * i + 1] = P[12 * i + 1]; posq[4 * i + 2] = P[12 * i + 2]; posq[4 * i + 3] = P[12 * i + 9]; }
		orc_direct_field(n, posq.data(), n, nullptr, eps, g.data(), nullptr, threads);
		double a2max = 0.0;
		for (std::uint64_t i = 0; i < n; ++i) {
			float* a = P + 12 * i;
			const double s_ = (double) G * a[9] / a[8];
			double a2 = 0.0;
			for (int d = 0; d < 3; ++d) {
				const double acc = s_ * g[3 * i + d];
				a2 += acc * acc;
				const float v_old = a[4 + d];
				const float v_new = (float) (v_old + acc * dt);
				a[4 + d] = v_new;
				a[d] = a[d] + (integrator == 0 ? v_new : v_old) * dt;
			}
			a2max = std::max(a2max, a2);
		}
		time += dt;
		const float am = (float) std::sqrt(a2max);
		if (dts) dts[s] = dt;
		if (amax) amax[s] = am;
		dt = orc_next_time_step(eta, len, am, dt0, dt_min, dt_max);
	}
	return time;
}

}  // extern "C"
