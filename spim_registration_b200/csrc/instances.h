// The engine kernels' launch instantiations, grouped by translation unit.  The product library is built from several
// translation units compiled in parallel (build.py): spim_b200.cu declares every launch below `extern template` and
// inst.cu, compiled once per group with -DSPIM_INST_GROUP=<group>, holds the explicit instantiations -- and with them the
// device code -- of its group.  Without -DSPIM_SPLIT_BUILD (the kernel emulator, a plain single-file nvcc build) everything
// is instantiated implicitly in spim_b200.cu as before.
#pragma once
#include "kernels.h"
#include "runtime.h"

namespace spim {
typedef XInvT<EPI_STORE, MATH_IEEE> XInvStore;
typedef XInvT<EPI_RATIO, MATH_IEEE> XInvRatioIeee;
typedef XInvT<EPI_RATIO, MATH_FAST> XInvRatioFast;
typedef XInvT<EPI_UPDATE, MATH_IEEE> XInvUpdateIeee;
typedef XInvT<EPI_UPDATE, MATH_FAST> XInvUpdateFast;
typedef XInvT<EPI_UPDATE, MATH_EXACT64> XInvUpdateExact64;
typedef XInvT<EPI_RATIO, MATH_FAST, true> XInvRatioFastFuse;       // brick mode with mapped peers: the epilogue also stores the neighbours' halo voxels
typedef XInvT<EPI_UPDATE, MATH_FAST, true> XInvUpdateFastFuse;
}

// (body, __launch_bounds__ max threads, min blocks per SM): exactly the launches of engine.h
#define SPIM_INSTANCES_COL_A(X) X(ColPass, 256, 1) X(ColPass, 384, 1)
#define SPIM_INSTANCES_COL_B(X) X(ColPass, 128, 5) X(ColPassR8, 128, 6)
#define SPIM_INSTANCES_COL_C(X) X(ColPassNarrow, 256, 1) X(ColPassNarrow, 128, 5)
#define SPIM_INSTANCES_COL_D(X) X(ColPassT, 512, 1)
#define SPIM_INSTANCES_X_A(X) X(XFwd, 192, 4) X(XInvStore, 256, 1)
#define SPIM_INSTANCES_X_B(X) X(XInvRatioFast, 256, 1) X(XInvRatioFast, 192, 4)
#define SPIM_INSTANCES_X_C(X) X(XInvRatioIeee, 256, 1) X(XInvRatioIeee, 128, 6)
#define SPIM_INSTANCES_X_D(X) X(XInvUpdateFast, 256, 1) X(XInvUpdateFast, 128, 5)
#define SPIM_INSTANCES_X_E(X) X(XInvUpdateIeee, 256, 1) X(XInvUpdateExact64, 256, 1)
#define SPIM_INSTANCES_X_F(X) X(XFwdT, 256, 3) X(XFwdT, 384, 2) X(XFwdT, 768, 1)
#define SPIM_INSTANCES_X_G(X) X(XFwdTNarrow, 256, 3) X(XFwdTNarrow, 384, 2)
#define SPIM_INSTANCES_X_H(X) X(XInvRatioFastFuse, 256, 1) X(XInvRatioFastFuse, 192, 4)
#define SPIM_INSTANCES_X_I(X) X(XInvUpdateFastFuse, 256, 1) X(XInvUpdateFastFuse, 128, 5)
#define SPIM_INSTANCE_GROUPS "COL_A", "COL_B", "COL_C", "COL_D", "X_A", "X_B", "X_C", "X_D", "X_E", "X_F", "X_G", "X_H", "X_I"
#define SPIM_INSTANCES_ALL(X)                                                                                      \
    SPIM_INSTANCES_COL_A(X) SPIM_INSTANCES_COL_B(X) SPIM_INSTANCES_COL_C(X) SPIM_INSTANCES_COL_D(X)                \
    SPIM_INSTANCES_X_A(X) SPIM_INSTANCES_X_B(X) SPIM_INSTANCES_X_C(X) SPIM_INSTANCES_X_D(X) SPIM_INSTANCES_X_E(X) SPIM_INSTANCES_X_F(X) SPIM_INSTANCES_X_G(X) SPIM_INSTANCES_X_H(X) SPIM_INSTANCES_X_I(X)

#if defined(SPIM_SPLIT_BUILD) && !defined(SPIM_HOST_EMU)
namespace spim {
#define SPIM_EXTERN_INST(Body, MAXT, MINB) \
    extern template void rt::launch<Body, MAXT, MINB>(const Body::Params&, long long, int, size_t, rt::Stream);
SPIM_INSTANCES_ALL(SPIM_EXTERN_INST)
#undef SPIM_EXTERN_INST
}
#endif
