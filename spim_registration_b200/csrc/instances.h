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
typedef XInvT<EPI_UPDATE, MATH_FAST, 5> XInvUpdateFastR5;     // last stage compiled for radices <= 5 only (engine.h, SPIM_XINV_R0)
typedef XInvT<EPI_UPDATE, MATH_FAST, 7> XInvUpdateFastR7;
typedef XInvP<EPI_STORE, MATH_IEEE> XInvPStore;               // persistent, TMA-fed (the default x-inverse kernels)
typedef XInvP<EPI_RATIO, MATH_IEEE> XInvPRatioIeee;
typedef XInvP<EPI_RATIO, MATH_FAST> XInvPRatioFast;
typedef XInvP<EPI_UPDATE, MATH_IEEE> XInvPUpdateIeee;
typedef XInvP<EPI_UPDATE, MATH_FAST> XInvPUpdateFast;
typedef XInvP<EPI_UPDATE, MATH_EXACT64> XInvPUpdateExact64;
}

#define SPIM_INSTANCES_COL_A(X) X(ColPass, 256, 1) X(ColPass, 384, 1)
#define SPIM_INSTANCES_COL_B(X) X(ColPass, 128, 5) X(ColPass, 256, 3)
#define SPIM_INSTANCES_COL_C(X) X(ColPassNarrow, 256, 1) X(ColPassNarrow, 128, 5)
#define SPIM_INSTANCES_COL_D(X) X(ColPassT, 512, 1) X(ColPassW, 256, 1)
#define SPIM_INSTANCES_COL_E(X) X(ColPass, 192, 3) X(ColPass, 128, 6)
#define SPIM_INSTANCES_COL_F(X) X(ColPassR8, 128, 6) X(ColPassR8, 256, 1)
#define SPIM_INSTANCES_X_A(X) X(XFwd, 256, 1) X(XFwd, 192, 4) X(XInvStore, 256, 1) X(XInvRatioIeee, 256, 1) X(XInvRatioIeee, 128, 6)
#define SPIM_INSTANCES_X_B(X) X(XInvRatioFast, 256, 1) X(XInvRatioFast, 128, 6) X(XInvUpdateFast, 256, 1)
#define SPIM_INSTANCES_X_C(X) X(XInvUpdateIeee, 256, 1) X(XInvUpdateExact64, 256, 1)
#define SPIM_INSTANCES_X_D(X) X(XInvUpdateFast, 128, 5) X(XInvUpdateFastR5, 128, 6) X(XInvUpdateFastR7, 128, 5)
#define SPIM_INSTANCES_X_E(X) X(XFwd, 160, 5) X(XInvRatioFast, 160, 5) X(XInvUpdateFastR5, 160, 5)
#define SPIM_INSTANCES_X_F(X) X(XFwdT, 256, 3) X(XFwdT, 384, 2) X(XFwdT, 512, 1)
#define SPIM_INSTANCES_X_G(X) X(XInvPRatioFast, 256, 3) X(XInvPRatioFast, 512, 1) X(XInvPRatioFast, 256, 2)
#define SPIM_INSTANCES_X_H(X) X(XInvPUpdateFast, 192, 3) X(XInvPUpdateFast, 160, 3) X(XInvPUpdateFast, 512, 1) X(XInvPUpdateFast, 256, 2)
#define SPIM_INSTANCES_X_I(X) X(XInvPStore, 256, 2) X(XInvPRatioIeee, 256, 2)
#define SPIM_INSTANCES_X_J(X) X(XInvPUpdateIeee, 256, 2) X(XInvPUpdateExact64, 256, 2)
#define SPIM_INSTANCE_GROUPS "X_G", "X_H", "X_I", "X_J", "X_F", "COL_A", "COL_B", "COL_C", "COL_D", "COL_E", "COL_F", "X_A", "X_B", "X_C", "X_D", "X_E"
#define SPIM_INSTANCES_ALL(X)                                                                                      \
    SPIM_INSTANCES_COL_A(X) SPIM_INSTANCES_COL_B(X) SPIM_INSTANCES_COL_C(X) SPIM_INSTANCES_COL_D(X)                \
    SPIM_INSTANCES_COL_E(X) SPIM_INSTANCES_COL_F(X)                                                                                        \
    SPIM_INSTANCES_X_A(X) SPIM_INSTANCES_X_B(X) SPIM_INSTANCES_X_C(X) SPIM_INSTANCES_X_D(X) SPIM_INSTANCES_X_E(X) SPIM_INSTANCES_X_F(X) \
    SPIM_INSTANCES_X_G(X) SPIM_INSTANCES_X_H(X) SPIM_INSTANCES_X_I(X) SPIM_INSTANCES_X_J(X)

#if defined(SPIM_SPLIT_BUILD) && !defined(SPIM_HOST_EMU)
namespace spim {
#define SPIM_EXTERN_INST(Body, MAXT, MINB) \
    extern template void rt::launch<Body, MAXT, MINB>(const Body::Params&, long long, int, size_t, rt::Stream);
SPIM_INSTANCES_ALL(SPIM_EXTERN_INST)
#undef SPIM_EXTERN_INST
}
#endif
