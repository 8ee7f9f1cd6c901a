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

int main() {
	for (int occ : {1, 2, 4}) {
		run<0>("x = x*a + b", occ);
		run<1>("x += a*b", occ);
		run<2>("x = x*imm + imm", occ);
		run<3>("x_i += x_j * a", occ);
		run<4>("x_i += x_j * x_k", occ);
	}
	return 0;
}
