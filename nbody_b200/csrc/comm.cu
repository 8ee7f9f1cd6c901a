// Multi-GPU plumbing (one process per GPU). Filled in by the distributed build;
// single-GPU simulations never touch it.
#include "common.cuh"

namespace nbody {
int comm_step_exchange(Sim&) { return NBODY_OK; }
int comm_partition(Sim&) { return NBODY_OK; }
void comm_destroy(Sim&) {}
}  // namespace nbody

extern "C" {
int nbody_cuda_comm_unique_id(uint8_t*) { nbody::set_error("distributed mode is not built"); return NBODY_ERR_COMM; }
int nbody_cuda_create_distributed(const nbody_cuda_config*, const nbody_particle*, uint64_t, uint64_t, uint64_t, int, int,
                                  const uint8_t*, nbody_cuda_sim**) {
	nbody::set_error("distributed mode is not built");
	return NBODY_ERR_COMM;
}
}
