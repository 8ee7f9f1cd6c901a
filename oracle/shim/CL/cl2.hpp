// Test-infrastructure shim (NOT product code): lets the reference's CPU-only
// naive_simulation.cpp compile without an OpenCL SDK. The reference pulls the
// OpenCL scalar/vector typedefs into its CPU path through
// include/nbody/device/types.h:12-15,45-74; only these names are needed.
#ifndef ORACLE_SHIM_CL2_HPP_
#define ORACLE_SHIM_CL2_HPP_
#include <cstdint>
typedef std::uint32_t cl_uint;
typedef std::int32_t  cl_int;
typedef float         cl_float;
typedef std::uint8_t  cl_uchar;
typedef std::uint64_t cl_ulong;
typedef union alignas(16) { cl_float s[4]; } cl_float4;
#endif
