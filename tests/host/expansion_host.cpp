// Host build of nbody_b200/csrc/expansion.cuh for CPU-side unit tests (g++).
// Exposes each operator for P = 2,3,4 through a C ABI used by tests/test_expansion.py.
#include <cmath>
#include <cstdint>
// plain-C++ stand-ins for CUDA's float2 and sm_100's two-wide FP32 operations, so that Expansion::derivatives2 compiles here
#define NBODY_HOST_F32X2_SHIM 1
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#include "../../nbody_b200/csrc/expansion.cuh"

using namespace nbody;

template <int P>
static void run_chain(const float* src, int ns, const float* cB_child, const float* cB, const float* cA,
                      const float* cA_child, const float* tgt, int nt, float eps, float* field, float* M_out, float* L_out) {
	using E = Expansion<P>;
	float Mc[E::NC] = {0}, Mp[E::NC] = {0};
	for (int s = 0; s < ns; ++s) E::p2m(Mc, src[4 * s] - cB_child[0], src[4 * s + 1] - cB_child[1], src[4 * s + 2] - cB_child[2], src[4 * s + 3]);
	E::m2m(Mp, Mc, cB_child[0] - cB[0], cB_child[1] - cB[1], cB_child[2] - cB[2]);
	float D[E::NC], Lp[E::NC] = {0}, Lc[E::NC] = {0};
	E::derivatives(cA[0] - cB[0], cA[1] - cB[1], cA[2] - cB[2], eps * eps, D);
	E::template m2l<0>(Lp, Mp, D);
	E::template l2l<0>(Lc, Lp, cA_child[0] - cA[0], cA_child[1] - cA[1], cA_child[2] - cA[2]);
	for (int t = 0; t < nt; ++t) {
		float gx, gy, gz;
		E::l2p(Lc, tgt[3 * t] - cA_child[0], tgt[3 * t + 1] - cA_child[1], tgt[3 * t + 2] - cA_child[2], gx, gy, gz);
		field[4 * t] = gx; field[4 * t + 1] = gy; field[4 * t + 2] = gz;
		field[4 * t + 3] = E::l2p_potential(Lc, tgt[3 * t] - cA_child[0], tgt[3 * t + 1] - cA_child[1], tgt[3 * t + 2] - cA_child[2]);
	}
	for (int a = 0; a < E::NC; ++a) { M_out[a] = Mp[a]; L_out[a] = Lc[a]; }
}

extern "C" {
int exp_ncoef(int p) { return ncoef(p); }
int exp_index(int i, int j, int k) { return mi_index(i, j, k); }
// sources in child cell of B -> P2M -> M2M to B -> M2L to A -> L2L to child of A -> L2P at targets
void exp_chain(int p, const float* src, int ns, const float* cB_child, const float* cB, const float* cA, const float* cA_child,
               const float* tgt, int nt, float eps, float* field, float* M_out, float* L_out) {
	if (p == 2) run_chain<2>(src, ns, cB_child, cB, cA, cA_child, tgt, nt, eps, field, M_out, L_out);
	if (p == 3) run_chain<3>(src, ns, cB_child, cB, cA, cA_child, tgt, nt, eps, field, M_out, L_out);
	if (p == 4) run_chain<4>(src, ns, cB_child, cB, cA, cA_child, tgt, nt, eps, field, M_out, L_out);
}
void exp_derivatives(int p, float x, float y, float z, float eps2, float* D) {
	if (p == 2) { float d[ncoef(2)]; Expansion<2>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(2); ++a) D[a] = d[a]; }
	if (p == 3) { float d[ncoef(3)]; Expansion<3>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(3); ++a) D[a] = d[a]; }
	if (p == 4) { float d[ncoef(4)]; Expansion<4>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(4); ++a) D[a] = d[a]; }
}
}

// both halves of the packed derivative tensor (Expansion::derivatives2) for two separation vectors
template <int P>
static void run_derivatives2(const float* a, const float* b, float eps2, float* Da, float* Db) {
	using E = Expansion<P>;
	float2 D[E::NC];
	E::derivatives2(make_float2(a[0], b[0]), make_float2(a[1], b[1]), make_float2(a[2], b[2]), eps2, D);
	for (int n = 0; n < E::NC; ++n) { Da[n] = D[n].x; Db[n] = D[n].y; }
}
extern "C" void exp_derivatives2(int p, const float* a, const float* b, float eps2, float* Da, float* Db) {
	if (p == 2) run_derivatives2<2>(a, b, eps2, Da, Db);
	else if (p == 3) run_derivatives2<3>(a, b, eps2, Da, Db);
	else run_derivatives2<4>(a, b, eps2, Da, Db);
}

// One source against two targets the way k_m2l_pair does it (masked derivatives2 + broadcast contraction m2l_bc), and the same
// two interactions with the scalar operators. d0/d1 = target - source separations, M = the source's multipole (ncoef(p) floats),
// keep bit t = target t accepts. Outputs: L (ncoef(p) floats) per target, packed and scalar.
template <int P, int PE>
static void run_pair(const float* d0, const float* d1, const float* M, float eps2, unsigned keep, float* L0, float* L1, float* R0, float* R1) {
	using E = Expansion<P>;
	float2 D2[Expansion<PE>::NC];
	Expansion<PE>::template derivatives2<true>(make_float2(d0[0], d1[0]), make_float2(d0[1], d1[1]), make_float2(d0[2], d1[2]), eps2, D2,
	                                           make_float2((keep & 1u) ? 1.0f : 0.0f, (keep & 2u) ? 1.0f : 0.0f));
	float2 L2[E::NC];
	for (int n = 0; n < E::NC; ++n) L2[n] = make_float2(0.0f, 0.0f);
	E::template m2l_bc<1, PE>(L2, M, D2);
	for (int n = 0; n < E::NC; ++n) { L0[n] = L2[n].x; L1[n] = L2[n].y; }
	float Da[Expansion<PE>::NC], Db[Expansion<PE>::NC], La[E::NC] = {0}, Lb[E::NC] = {0};
	Expansion<PE>::derivatives(d0[0], d0[1], d0[2], eps2, Da);
	Expansion<PE>::derivatives(d1[0], d1[1], d1[2], eps2, Db);
	if (keep & 1u) E::template m2l<1, PE>(La, M, Da);
	if (keep & 2u) E::template m2l<1, PE>(Lb, M, Db);
	for (int n = 0; n < E::NC; ++n) { R0[n] = La[n]; R1[n] = Lb[n]; }
}
extern "C" void exp_pair(int p, int pe, const float* d0, const float* d1, const float* M, float eps2, unsigned keep, float* L0, float* L1,
                         float* R0, float* R1) {
	if (p == 4 && pe == 4) run_pair<4, 4>(d0, d1, M, eps2, keep, L0, L1, R0, R1);
	else if (p == 4 && pe == 3) run_pair<4, 3>(d0, d1, M, eps2, keep, L0, L1, R0, R1);
	else if (p == 3 && pe == 3) run_pair<3, 3>(d0, d1, M, eps2, keep, L0, L1, R0, R1);
	else if (p == 3 && pe == 2) run_pair<3, 2>(d0, d1, M, eps2, keep, L0, L1, R0, R1);
	else run_pair<2, 2>(d0, d1, M, eps2, keep, L0, L1, R0, R1);
}
