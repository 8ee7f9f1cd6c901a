// Test-infrastructure shim (NOT product code): the handful of OpenCL C features the reference's
// device kernels use (/root/reference/src/{interaction,field,force,moment,verify}.cl and
// include/nbody/device/types.h:45-51 in its __KERNEL__ branch), defined for a host C++ compiler so
// that oracle/Makefile can compile those kernel sources WHERE THEY LIE, unmodified, into
// oracle/_ref/libclref.so. oracle/ref_cl_harness.cpp includes this header inside `namespace refcl`
// and then the .cl files themselves.
//
// What OpenCL leaves to the implementation and what is chosen here (all stated in DESIGN.md 3):
//  * float arithmetic is IEEE binary32, one rounding per operation, no contraction (the Makefile
//    passes -ffp-contract=off); sqrt and / are correctly rounded (OpenCL 1.2 allows 3 / 2.5 ulp);
//  * dot(float4, float4) is evaluated left to right over x, y, z, w;
//  * work items of an NDRange run sequentially in ascending (group, local id 1, local id 0) order.
#ifndef ORACLE_SHIM_OPENCL_C_HOST_H_
#define ORACLE_SHIM_OPENCL_C_HOST_H_

#ifndef REFCL_INSIDE_NAMESPACE
#error "include from oracle/ref_cl_harness.cpp, inside namespace refcl"
#endif

typedef unsigned int uint;
typedef unsigned char uchar;

// OpenCL's float4: 16 bytes, 16-byte aligned, components x y z w, component-wise operators,
// scalars widen to vectors (`vector_t v = 0.0;`, src/moment.cl:23-25).
struct alignas(16) float4 {
	float x, y, z, w;
	float4() = default;
	float4(float s) : x(s), y(s), z(s), w(s) {}
	float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
};
inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
inline float4 operator*(float s, float4 a) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline float4 operator*(float4 a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator/(float4 a, float s) { return float4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline float4 operator/(float4 a, int s) { return a / (float) s; }  // `dimensions / 2`, src/interaction.cl:66
inline float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
inline float dot(float4 a, float4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
// (vector_t)(a, b, c, d) literals (src/moment.cl:43-54, src/force.cl:35,66) are the one construct a C++ compiler reads
// differently (a cast of a comma expression): interaction.cl, field.cl and verify.cl have none and compile as they are;
// for moment.cl and force.cl the Makefile writes a transient copy into oracle/_ref/gen/ (removed after the link) with exactly the token sequence
// "(vector_t) (" replaced by "make_vector_t(" and nothing else changed.
inline float4 make_vector_t(float a, float b, float c, float d) { return float4(a, b, c, d); }

inline float sqrt(float v) { return __builtin_sqrtf(v); }
inline uint min(uint a, uint b) { return b < a ? b : a; }

// work-item functions: the harness sets these before every kernel invocation
struct WorkItem { std::size_t group[3], local[3], local_size[3]; };
extern thread_local WorkItem g_work_item;
inline std::size_t get_group_id(uint d) { return g_work_item.group[d]; }
inline std::size_t get_local_id(uint d) { return g_work_item.local[d]; }
inline std::size_t get_local_size(uint d) { return g_work_item.local_size[d]; }
inline std::size_t get_global_id(uint d) { return g_work_item.group[d] * g_work_item.local_size[d] + g_work_item.local[d]; }

// atomics on global memory (sequential execution: plain read-modify-write)
inline uint atomic_inc(uint* p) { const uint old = *p; *p = old + 1u; return old; }
inline uint atomic_max(uint* p, uint v) { const uint old = *p; if (v > old) *p = v; return old; }

// address-space and function qualifiers
#define kernel
#define global

#endif
