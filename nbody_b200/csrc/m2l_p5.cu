// Stage 4 at expansion order 5 (kernels: m2l_kernels.cuh).
#include "m2l_kernels.cuh"

namespace nbody {

void launch_m2l_p5(Sim& s) { m2l_t<5>(s); }
void launch_l2l_p5(Sim& s) { l2l_t<5>(s); }

}  // namespace nbody
