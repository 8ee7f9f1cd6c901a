// Host build of nbody_b200/csrc/expansion.cuh for CPU-side unit tests (g++).
// Exposes each operator for P = 2,3,4 through a C ABI used by tests/test_expansion.py.
#include <cmath>
#include <cstdint>
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
	if (p == 5) run_chain<5>(src, ns, cB_child, cB, cA, cA_child, tgt, nt, eps, field, M_out, L_out);
}
void exp_derivatives(int p, float x, float y, float z, float eps2, float* D) {
	if (p == 2) { float d[ncoef(2)]; Expansion<2>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(2); ++a) D[a] = d[a]; }
	if (p == 3) { float d[ncoef(3)]; Expansion<3>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(3); ++a) D[a] = d[a]; }
	if (p == 4) { float d[ncoef(4)]; Expansion<4>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(4); ++a) D[a] = d[a]; }
	if (p == 5) { float d[ncoef(5)]; Expansion<5>::derivatives(x, y, z, eps2, d); for (int a = 0; a < ncoef(5); ++a) D[a] = d[a]; }
}
}
