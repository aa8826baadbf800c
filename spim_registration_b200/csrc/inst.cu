// One group of the engine kernels' launch instantiations (see instances.h); compiled once per group by build.py:
//   nvcc ... -DSPIM_SPLIT_BUILD -DSPIM_INST_GROUP=COL_A -c inst.cu -o inst_COL_A.o
#include "instances.h"

#ifdef SPIM_INST_GROUP      // compiled without a group this file contributes nothing
namespace spim {
#define SPIM_DEFINE_INST(Body, MAXT, MINB) \
    template void rt::launch<Body, MAXT, MINB>(const Body::Params&, long long, int, size_t, rt::Stream);
#define SPIM_CAT_(a, b) a##b
#define SPIM_GROUP_LIST(g) SPIM_CAT_(SPIM_INSTANCES_, g)
SPIM_GROUP_LIST(SPIM_INST_GROUP)(SPIM_DEFINE_INST)
}
#endif
