// Cartesian Taylor expansions of total order <= P for the softened kernel
// phi(x) = (|x|^2 + eps^2)^(-1/2), compile-time unrolled so every coefficient
// lives in a register. Replaces the reference's order-<=2 moments
// (src/moment.cl:27-54, never consumed: SURVEY D8) and its order-0 far field
// (src/field.cl:35-47,187-210) with real P2M/M2M/M2L/L2L/L2P operators.
//
// Conventions (identical to oracle/oracle.cpp orc_fmm_field, except that locals
// are stored as pure derivatives so that M2L is nothing but FMAs):
//   M_m  = sum_j q_j (y_j - c)^m / m!                         (multipole about c)
//   Lt_n = d^n Phi / dx^n (c_A) = sum_m (-1)^|m| M_m D_{n+m}(c_A - c_B),  |n|+|m| <= P
//   Phi(x) = sum_n Lt_n (x - c_A)^n / n!
// Multi-index order: total order ascending; inside an order i (x power)
// descending, then j descending — the same enumeration as the oracle's Idx.
//
// Everything here is __host__ __device__ so tests/ can compile it with g++ and
// compare against the FP64 oracle without a GPU.
#pragma once

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

namespace nbody {

NB_HD constexpr int mi_index(int i, int j, int k) {
	const int o = i + j + k, a = o - i;
	return o * (o + 1) * (o + 2) / 6 + a * (a + 1) / 2 + (a - j);
}
NB_HD constexpr int ncoef(int p) { return (p + 1) * (p + 2) * (p + 3) / 6; }
// Loop- and recursion-free so that, once the multi-index loops are unrolled, every
// coefficient folds to an immediate (a recursive constexpr is NOT inlined by nvcc in
// a non-constant-expression context and leaves CALLs and FP divisions in the SASS).
NB_HD constexpr int facti(int n) {
	return n <= 1 ? 1 : n == 2 ? 2 : n == 3 ? 6 : n == 4 ? 24 : n == 5 ? 120 : n == 6 ? 720 : n == 7 ? 5040 : 40320;
}
NB_HD constexpr float factf(int n) { return (float) facti(n); }
NB_HD constexpr float rfact3(int i, int j, int k) { return 1.0f / (float) (facti(i) * facti(j) * facti(k)); }

// Record stride of a multipole / local in memory: padded to a multiple of 4 floats
// so that records are 16-byte aligned (float4 / cp.async traffic).
NB_HD constexpr int coef_stride(int p) { return (ncoef(p) + 3) / 4 * 4; }

#define NB_FOR_ORDER(o, lo, hi) _Pragma("unroll") for (int o = (lo); o <= (hi); ++o)
#define NB_FOR_MI(o, i, j, k, lo, hi)                              \
	_Pragma("unroll") for (int o = (lo); o <= (hi); ++o)              \
	_Pragma("unroll") for (int i = o; i >= 0; --i)                    \
	_Pragma("unroll") for (int j = o - i; j >= 0; --j)                \
	if (const int k = o - i - j; true)

template <int P>
struct Expansion {
	static constexpr int NC = ncoef(P);

	// mono[n] = r^n (plain monomials) for |n| <= P
	NB_HD static void monomials(float x, float y, float z, float (&mono)[NC]) {
		mono[0] = 1.0f;
		NB_FOR_MI(o, i, j, k, 1, P) {
			if (i > 0) mono[mi_index(i, j, k)] = mono[mi_index(i - 1, j, k)] * x;
			else if (j > 0) mono[mi_index(i, j, k)] = mono[mi_index(i, j - 1, k)] * y;
			else mono[mi_index(i, j, k)] = mono[mi_index(i, j, k - 1)] * z;
		}
	}

	// P2M: M_m += q r^m / m!     (r = particle - centre)
	NB_HD static void p2m(float (&M)[NC], float x, float y, float z, float q) {
		float mono[NC];
		monomials(x, y, z, mono);
		NB_FOR_MI(o, i, j, k, 0, P) {
			M[mi_index(i, j, k)] += (q * rfact3(i, j, k)) * mono[mi_index(i, j, k)];
		}
	}

	// M2M: Mp_n += sum_{m <= n} Mc_m d^(n-m)/(n-m)!   (d = child centre - parent centre)
	NB_HD static void m2m(float (&Mp)[NC], const float (&Mc)[NC], float dx, float dy, float dz) {
		float mono[NC];
		monomials(dx, dy, dz, mono);
		NB_FOR_MI(o, i, j, k, 0, P) {
			float acc = 0.0f;
			NB_FOR_MI(o2, a, b, c, 0, o) {
				if (a <= i && b <= j && c <= k) {
					acc += Mc[mi_index(a, b, c)] * (rfact3(i - a, j - b, k - c) * mono[mi_index(i - a, j - b, k - c)]);
				}
			}
			Mp[mi_index(i, j, k)] += acc;
		}
	}

	// Derivative tensor D_n = d^n/dx^n (|x|^2+eps^2)^(-1/2), |n| <= P:
	//   D_n = sum_{2j<=n} c(n,j) g_{|n|-|j|} x^(n-2j),  g_k = (-1)^k (2k-1)!! R^-(2k+1),
	//   c(n,j) = prod_d n_d! / (j_d! (n_d-2j_d)! 2^j_d).
	NB_HD static void derivatives(float x, float y, float z, float eps2, float (&D)[NC]) {
		float mono[NC];
		monomials(x, y, z, mono);
		const float R2 = x * x + y * y + z * z + eps2;
#if defined(__CUDA_ARCH__)
		float inv;  // one MUFU.RSQ: R2 >= eps^2 or a cell separation squared, never denormal, so no fix-up code is wanted
		asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(R2));
#else
		const float inv = 1.0f / sqrtf(R2);
#endif
		const float inv2 = inv * inv;
		float g[P + 1];
		g[0] = inv;
		NB_FOR_ORDER(q, 1, P) g[q] = -(float) (2 * q - 1) * g[q - 1] * inv2;
		NB_FOR_MI(o, i, j, k, 0, P) {
			float acc = 0.0f;
			_Pragma("unroll") for (int a = 0; a <= P / 2; ++a)
			_Pragma("unroll") for (int b = 0; b <= P / 2; ++b)
			_Pragma("unroll") for (int c = 0; c <= P / 2; ++c)
			if (2 * a <= i && 2 * b <= j && 2 * c <= k) {
				// integer: number of ways to pair up 2a,2b,2c of the i,j,k derivatives
				const int coef = (facti(i) / (facti(a) * facti(i - 2 * a) << a)) * (facti(j) / (facti(b) * facti(j - 2 * b) << b)) *
				                 (facti(k) / (facti(c) * facti(k - 2 * c) << c));
				acc += ((float) coef * g[o - a - b - c]) * mono[mi_index(i - 2 * a, j - 2 * b, k - 2 * c)];
			}
			D[mi_index(i, j, k)] = acc;
		}
	}

	// M2L: Lt_n += sum_{|m| <= PE-|n|} (-1)^|m| M_m D_{n+m} for LO <= |n| <= PE.
	// LO = lowest local order kept: 1 when only the field (gradient) is needed, 0 to carry the
	// potential too. PE <= P is the order this pair is evaluated at (adaptive-order M2L: well
	// separated pairs run at P-1); D then only needs derivatives up to order PE.
	// Source-major loop order: one multipole coefficient is live at a time (it may come straight
	// from shared memory), the P-dependent live set is L and D only.
	template <int LO, int PE = P, typename MT, int ND>
	NB_HD static void m2l(float (&L)[NC], const MT& M, const float (&D)[ND]) {
		static_assert(PE <= P && ND >= ncoef(PE), "derivative tensor too short for the evaluation order");
		NB_FOR_MI(o2, a, b, c, 0, PE - LO) {
			const float mraw = M[mi_index(a, b, c)];
			const float m = (o2 & 1) ? -mraw : mraw;
			NB_FOR_MI(o, i, j, k, LO, PE - o2) {
				L[mi_index(i, j, k)] += m * D[mi_index(i + a, j + b, k + c)];
			}
		}
	}

	// L2L: Lc_n += sum_{|k| <= P-|n|} Lp_{n+k} d^k / k!    (d = child centre - parent centre)
	template <int LO>
	NB_HD static void l2l(float (&Lc)[NC], const float (&Lp)[NC], float dx, float dy, float dz) {
		float mono[NC];
		monomials(dx, dy, dz, mono);
		NB_FOR_MI(o, i, j, k, LO, P) {
			float acc = 0.0f;
			NB_FOR_MI(o2, a, b, c, 0, P - o) {
				acc += Lp[mi_index(i + a, j + b, k + c)] * (rfact3(a, b, c) * mono[mi_index(a, b, c)]);
			}
			Lc[mi_index(i, j, k)] += acc;
		}
	}

	// L2P: field g = grad Phi at offset r from the centre: g_x = sum_{n, i>=1} Lt_n r^(n-e_x)/(n-e_x)!
	NB_HD static void l2p(const float (&L)[NC], float x, float y, float z, float& gx, float& gy, float& gz) {
		float mono[NC];
		monomials(x, y, z, mono);
		float ax = 0.0f, ay = 0.0f, az = 0.0f;
		NB_FOR_MI(o, i, j, k, 1, P) {
			const float l = L[mi_index(i, j, k)];
			if (i > 0) ax += l * (rfact3(i - 1, j, k) * mono[mi_index(i - 1, j, k)]);
			if (j > 0) ay += l * (rfact3(i, j - 1, k) * mono[mi_index(i, j - 1, k)]);
			if (k > 0) az += l * (rfact3(i, j, k - 1) * mono[mi_index(i, j, k - 1)]);
		}
		gx = ax; gy = ay; gz = az;
	}

	// potential Phi at offset r (needs locals carried from order 0)
	NB_HD static float l2p_potential(const float (&L)[NC], float x, float y, float z) {
		float mono[NC];
		monomials(x, y, z, mono);
		float p = 0.0f;
		NB_FOR_MI(o, i, j, k, 0, P) p += L[mi_index(i, j, k)] * (rfact3(i, j, k) * mono[mi_index(i, j, k)]);
		return p;
	}
};

}  // namespace nbody
