// Microbenchmark: sustained FP32 FFMA issue rate on B200 for the operand patterns the
// FMM kernels use. Prints warp-instructions per cycle per SM sub-partition and TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a0, float b0, int iters) {
	float x[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) x[i] = a0 * (threadIdx.x + i);
	float a = a0, b = b0, c = a0 + b0, d = a0 - b0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int i = 0; i < 16; ++i) {
				if (MODE == 0) x[i] = fmaf(x[i], a, b);            // x = x*a + b   (3 distinct registers)
				if (MODE == 1) x[i] = fmaf(a, b, x[i]);            // x += a*b      (accumulate, shared multiplicands)
				if (MODE == 2) x[i] = fmaf(x[i], 1.0001f, 0.5f);   // immediates
				if (MODE == 3) x[i] = fmaf(x[(i + 5) & 15], a, x[i]);  // x_i += x_j * a   (M2L-like: acc + reg*shared)
				if (MODE == 4) x[i] = fmaf(x[(i + 5) & 15], x[(i + 11) & 15], x[i]);  // three distinct varying registers
			}
		}
		a += 1e-9f; b -= 1e-9f; c += d;
	}
	float s = c;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int blocks_per_sm) {
	int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
	const int blocks = p.multiProcessorCount * blocks_per_sm, iters = 4000;
	float* out; cudaMalloc(&out, (size_t) blocks * 256 * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, 10);
	cudaEventRecord(e0);
	k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double ffma = (double) blocks * 256 * iters * 8 * 16;
	int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	printf("%-40s occ=%d blk/SM  %.2f ms  %.2f TFLOP/s  (%.3f warp-FFMA/clk/SMSP at %d MHz nominal)\n", name, blocks_per_sm, ms,
	       2 * ffma / ms / 1e9, ffma / 32 / (ms * 1e-3) / (khz * 1e3) / (p.multiProcessorCount * 4), khz / 1000);
	cudaFree(out);
}

// ---- two-wide FP32 (FFMA2 / FADD2 / FMUL2, sm_100): does one packed instruction cost one issue slot and two pipe cycles? ----
// MODE 0: x2 += a2*b2 (pure FFMA2, register pairs);  MODE 1: x2 += s*b2 with a 32-bit broadcast operand (the P2P kernels' form);
// MODE 2: the P2P interaction chain, packed (3 FADD2, 6 FFMA2, 2 MUFU.RSQ, 3 FMUL2 per two pair evaluations);
// MODE 3: the same chain with scalar instructions (12 + 1 per evaluation) — the ceiling of the present kernels.
template <int MODE>
__global__ void __launch_bounds__(256) k2(float* out, float a0, float b0, int iters) {
	float2 x[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) x[i] = make_float2(a0 * (threadIdx.x + i), b0 * (threadIdx.x + i + 1));
	float2 a = make_float2(a0, b0), b = make_float2(b0, a0);
	float flops_guard = 0.f;
	__shared__ float4 tile[64 + 16];  // sources of the P2P chains: one broadcast LDS.128 per source, as in k_direct
	for (int i = threadIdx.x; i < 80; i += blockDim.x) tile[i] = make_float4(a0 * i, b0 * i, a0 + i, 1.0f);
	__syncthreads();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
			if (MODE == 0) {
#pragma unroll
				for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(a, b, x[i]);
			} else if (MODE == 1) {
#pragma unroll
				for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(make_float2(a0, a0), b, x[i]);
			} else if (MODE == 2) {
				// source (a.x, a.y, b.x, q = b.y) against target pairs x[0..2] = negated coordinates; accumulators x[3..5]; two chains
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					const float4 sv = tile[(it & 63) + 2 * r + c];
					const float sx = sv.x, sy = sv.y, sz = sv.z, q = sv.w;
					const float2 dx = __fadd2_rn(make_float2(sx, sx), x[0]), dy = __fadd2_rn(make_float2(sy, sy), x[1]), dz = __fadd2_rn(make_float2(sz, sz), x[2]);
					const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __ffma2_rn(dx, dx, make_float2(1e-4f, 1e-4f))));
					float2 inv;
					asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(r2.x));
					asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(r2.y));
					const float2 w = __fmul2_rn(__fmul2_rn(make_float2(q, q), inv), __fmul2_rn(inv, inv));
					x[3 + 0] = __ffma2_rn(w, dx, x[3 + 0]);
					x[3 + 1] = __ffma2_rn(w, dy, x[3 + 1]);
					x[3 + 2] = __ffma2_rn(w, dz, x[3 + 2]);
				}
			} else {
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const float4 sv = tile[(it & 63) + 2 * r + (c & 1)];
					const float sx = sv.x, sy = sv.y, sz = sv.z, q = sv.w;
					const float tx = (c & 2) ? x[0].y : x[0].x, ty = (c & 2) ? x[1].y : x[1].x, tz = (c & 2) ? x[2].y : x[2].x;
					const float dx = sx + tx, dy = sy + ty, dz = sz + tz;
					const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, 1e-4f)));
					float inv;
					asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2));
					const float w = (q * inv) * (inv * inv);
					float& ax = (c & 2) ? x[3].y : x[3].x;
					float& ay = (c & 2) ? x[4].y : x[4].x;
					float& az = (c & 2) ? x[5].y : x[5].x;
					ax = fmaf(w, dx, ax); ay = fmaf(w, dy, ay); az = fmaf(w, dz, az);
				}
			}
		}
		a.x += 1e-9f; b.y -= 1e-9f; flops_guard += a.x;
	}
	float s = flops_guard;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run2(const char* name, int blocks_per_sm) {
	int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
	const int blocks = p.multiProcessorCount * blocks_per_sm, iters = 4000;
	float* out; cudaMalloc(&out, (size_t) blocks * 256 * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k2<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, 10);
	cudaEventRecord(e0);
	k2<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	const double threads = (double) blocks * 256, reps = (double) iters * 8;
	if (MODE <= 1) {
		const double ffma = threads * reps * 8 * 2;  // scalar-equivalent FMAs
		printf("%-52s occ=%d blk/SM  %.2f ms  %.2f TFLOP/s  (%.3f warp-FFMA2/clk/SMSP at %d MHz nominal)\n", name, blocks_per_sm, ms, 2 * ffma / ms / 1e9,
		       ffma / 2 / 32 / (ms * 1e-3) / (khz * 1e3) / (p.multiProcessorCount * 4), khz / 1000);
	} else {
		const double evals = threads * reps * 4;  // pair evaluations
		printf("%-52s occ=%d blk/SM  %.2f ms  %.2f TFLOP/s by the 20-flop convention  (%.2f clk per warp pair evaluation per SMSP)\n", name, blocks_per_sm, ms,
		       20 * evals / ms / 1e9, (ms * 1e-3) * (khz * 1e3) * (p.multiProcessorCount * 4) / (evals / 32));
	}
	cudaFree(out);
}

// Dependent-issue latency: one chain per thread, one warp per SM sub-partition (128 threads per SM), clock64 around the loop.
template <bool PACKED>
__global__ void k_lat(float* out, long long* cycles, float a0, int iters) {
	float2 x = make_float2(a0 * threadIdx.x, a0);
	const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(0.5f, 0.25f);
	const long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r) {
			if (PACKED) x = __ffma2_rn(x, a, b);
			else x.x = fmaf(x.x, a.x, b.x);
		}
	}
	const long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = x.x + x.y;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <bool PACKED>
void run_lat(const char* name) {
	float* out; long long* cyc; cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&cyc, 8);
	const int iters = 4000;
	k_lat<PACKED><<<148, 128>>>(out, cyc, 1.0001f, iters);
	long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("%-52s %.2f cycles per dependent instruction\n", name, (double) h / (iters * 16.0));
	cudaFree(out); cudaFree(cyc);
}

int main() {
	run_lat<false>("FFMA dependent chain");
	run_lat<true>("FFMA2 dependent chain");
	for (int occ : {1, 2, 4}) {
		run2<0>("FFMA2 x2 += a2*b2", occ);
		run2<1>("FFMA2 x2 += s*b2 (32-bit broadcast operand)", occ);
		run2<2>("P2P chain, two-wide (12 packed + 2 MUFU per 2 evals)", occ);
		run2<3>("P2P chain, scalar (12 + 1 MUFU per eval)", occ);
	}
	for (int occ : {1, 2, 4}) {
		run<0>("x = x*a + b", occ);
		run<1>("x += a*b", occ);
		run<2>("x = x*imm + imm", occ);
		run<3>("x_i += x_j * a", occ);
		run<4>("x_i += x_j * x_k", occ);
	}
	return 0;
}
