// NCCL bound at run time (shared by comm.cu and let.cu).
#pragma once
#include <nccl.h>  // types and enums only: the library itself is resolved with dlopen (nccl_load, comm.cu)

#include "common.cuh"

#include <string>

namespace nbody {

// NCCL is resolved with dlopen at the first distributed call instead of being a link-time
// dependency: a host process that also runs PyTorch already carries its own libnccl.so.2
// (a newer one than the system's), and two copies of one SONAME cannot coexist. RTLD_NOLOAD
// first re-uses whatever the process has loaded; a single-GPU run never touches NCCL at all.
struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;  // optional (NCCL >= 2.18); last argument: ncclConfig_t*
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
extern NcclApi g_nccl;
bool nccl_load();

#define NB_NCCL_CHECK(expr)                                                                  \
	do {                                                                                        \
		ncclResult_t _r = (expr);                                                                 \
		if (_r != ncclSuccess) {                                                                  \
			::nbody::set_error(std::string(#expr) + ": " + ::nbody::g_nccl.GetErrorString(_r));       \
			return NBODY_ERR_COMM;                                                                  \
		}                                                                                         \
	} while (0)

}  // namespace nbody
