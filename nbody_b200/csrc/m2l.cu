// Stage 4 launchers: orders 2 and 3 are instantiated here, 4 and 5 in their own translation units (m2l_kernels.cuh has the kernels).
#include "m2l_kernels.cuh"

namespace nbody {

void launch_m2l_p4(Sim& s);  // m2l_p4.cu
void launch_l2l_p4(Sim& s);
void launch_m2l_p5(Sim& s);  // m2l_p5.cu
void launch_l2l_p5(Sim& s);

void launch_m2l(Sim& s) {
	switch (s.cfg.order) {
		case 2: m2l_t<2>(s); break;
		case 3: m2l_t<3>(s); break;
		case 5: launch_m2l_p5(s); break;
		default: launch_m2l_p4(s); break;
	}
}
void launch_l2l(Sim& s) {
	switch (s.cfg.order) {
		case 2: l2l_t<2>(s); break;
		case 3: l2l_t<3>(s); break;
		case 5: launch_l2l_p5(s); break;
		default: launch_l2l_p4(s); break;
	}
}

}  // namespace nbody
