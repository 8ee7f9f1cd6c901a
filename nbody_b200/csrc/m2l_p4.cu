// Stage 4 at expansion order 4 (kernels: m2l_kernels.cuh).
#include "m2l_kernels.cuh"

namespace nbody {

void launch_m2l_p4(Sim& s) { m2l_t<4>(s); }
void launch_l2l_p4(Sim& s) { l2l_t<4>(s); }

}  // namespace nbody
