/* Plain C99 client of include/nbody_cuda.h: proves the header is C (not C++), that POD structs and sizes are what the
 * bindings assume, and exercises every entry point that needs no device (defaults, the time-step rule, the rebalancing
 * rule, checkpoint write / info / read, argument validation). Exit code 0 = all checks passed. A maintainer binding the
 * library from another language (cgo, JNI, ctypes) can read this file as the reference call sequence.
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -Iinclude examples/abi_host_check.c -Lnbody_b200 -lnbody_cuda \
 *       -Wl,-rpath,$PWD/nbody_b200 -o abi_host_check && ./abi_host_check /tmp/check.ckp */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nbody_cuda.h"

#define CHECK(cond)                                                                      \
	do {                                                                                    \
		if (!(cond)) {                                                                        \
			fprintf(stderr, "%s:%d: check failed: %s (last error: %s)\n", __FILE__, __LINE__, #cond, nbody_cuda_last_error()); \
			return 1;                                                                           \
		}                                                                                     \
	} while (0)

int main(int argc, char** argv) {
	const char* path = argc > 1 ? argv[1] : "abi_host_check.ckp";
	enum { N = 37 };
	nbody_cuda_config cfg;
	nbody_particle particles[N], back[N];
	uint32_t orig[N], orig_back[N];
	nbody_checkpoint_header hdr, info;
	uint32_t bounds_in[3] = {0u, 600u, 1000u}, bounds_out[3];
	float work[2] = {3.0f, 1.0f};
	int i;

	/* layouts the bindings rely on */
	CHECK(sizeof(nbody_particle) == 48);
	CHECK(sizeof(nbody_cuda_config) == 92);
	CHECK(sizeof(nbody_checkpoint_header) == 152);
	CHECK(sizeof(nbody_cuda_stats) == 12 * 8 + 10 * 4 + 3 * 8 + 4 * 4);

	nbody_cuda_default_config(&cfg);
	CHECK(cfg.abi_version == NBODY_CUDA_ABI_VERSION && cfg.leaf_capacity == 8 && cfg.order == 4 && cfg.max_depth == 21);
	nbody_cuda_tuned_config(&cfg);
	CHECK(cfg.abi_version == NBODY_CUDA_ABI_VERSION && cfg.leaf_capacity == 48 && cfg.order == 4 && cfg.mac_ratio == 0.5f);
	nbody_cuda_default_config(&cfg);
	CHECK(cfg.time_step == 0.001f && cfg.softening == 0.01f && cfg.mac_ratio == 0.5f && cfg.time_step_eta == 0.0f);

	/* the variable-time-step rule: off by default, eta * sqrt(softening / a_max) clamped otherwise */
	CHECK(nbody_cuda_next_time_step(&cfg, 100.0f) == cfg.time_step);
	cfg.time_step_eta = 0.1f;
	cfg.time_step = 1.0f;
	CHECK(fabsf(nbody_cuda_next_time_step(&cfg, 400.0f) / (0.1f * sqrtf(0.01f / 400.0f)) - 1.0f) < 1e-6f);
	cfg.time_step_min = 0.01f;
	CHECK(nbody_cuda_next_time_step(&cfg, 1e12f) == 0.01f);
	nbody_cuda_default_config(&cfg);

	/* per-step load rebalancing, host rule: rank 0 was three times slower, so its slice shrinks */
	CHECK(nbody_cuda_rebalance(2, bounds_in, work, 1.0f, bounds_out) == NBODY_OK);
	CHECK(bounds_out[0] == 0u && bounds_out[2] == 1000u && bounds_out[1] == 400u);
	CHECK(nbody_cuda_rebalance(0, bounds_in, work, 1.0f, bounds_out) == NBODY_ERR_INVALID);

	/* checkpoint round trip through files, host only */
	memset(particles, 0, sizeof(particles));
	for (i = 0; i < N; ++i) {
		particles[i].position[0] = 0.01f * (float) i;
		particles[i].velocity[2] = -0.5f * (float) i;
		particles[i].mass = 1.0f + (float) i;
		particles[i].charge = 2.0f;
		orig[i] = (uint32_t) (N - 1 - i);
	}
	memset(&hdr, 0, sizeof(hdr));
	hdr.n_particles = N;
	hdr.steps_done = 12;
	hdr.time = 0.012f;
	hdr.next_time_step = 5e-4f;
	hdr.config = cfg;
	CHECK(nbody_cuda_checkpoint_write(path, &hdr, particles, orig) == NBODY_OK);
	CHECK(nbody_cuda_checkpoint_info(path, &info) == NBODY_OK);
	CHECK(info.magic == NBODY_CHECKPOINT_MAGIC && info.version == NBODY_CHECKPOINT_VERSION && info.header_bytes == sizeof(info));
	CHECK(info.n_particles == N && info.steps_done == 12 && info.time == 0.012f && info.next_time_step == 5e-4f);
	CHECK(info.config.order == 4 && info.checksum != 0);
	CHECK(nbody_cuda_checkpoint_read(path, back, orig_back, N) == NBODY_OK);
	CHECK(memcmp(back, particles, sizeof(particles)) == 0 && memcmp(orig_back, orig, sizeof(orig)) == 0);
	CHECK(nbody_cuda_checkpoint_read(path, back, orig_back, N - 1) == NBODY_ERR_INVALID);
	CHECK(nbody_cuda_checkpoint_info("/nonexistent/dir/x.ckp", &info) == NBODY_ERR_INVALID);
	CHECK(strlen(nbody_cuda_last_error()) > 0);

	/* argument validation happens before any device is touched */
	{
		nbody_cuda_sim* sim = NULL;
		CHECK(nbody_cuda_create(&cfg, NULL, N, &sim) == NBODY_ERR_INVALID && sim == NULL);
		cfg.order = 9;
		CHECK(nbody_cuda_create(&cfg, particles, N, &sim) == NBODY_ERR_INVALID);
		cfg.order = 4;
		CHECK(nbody_cuda_num_particles(NULL) == 0);
		CHECK(nbody_cuda_step(NULL, NULL) == NBODY_ERR_INVALID);
		nbody_cuda_destroy(NULL);
	}
	remove(path);
	printf("abi_host_check: ok\n");
	return 0;
}
