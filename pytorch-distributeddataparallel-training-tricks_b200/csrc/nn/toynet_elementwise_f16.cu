// fp16-activation build of the bandwidth-bound Toy_Net kernels: the same source as toynet_elementwise.cu with
// act_t = __half (see the "activation precision" block at the top of that file).  Provides the __half overloads of the
// launchers declared in include/ddpt_toynet.h and launch_head_f16.
#define DDPT_ACT_F16 1
#include "toynet_elementwise.cu"
