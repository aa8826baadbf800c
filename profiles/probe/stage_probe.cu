// Static SASS probe: one radix stage of the column pass compiled on its own, so that the instruction mix of a work item can
// be read off `cuobjdump -sass` in seconds (profiles/probe/count.sh).  Not part of the library.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -cubin -DPR=8 -DPINV=0 -DPTW=1 -DPSRC=0 -DPDST=0 -DPW=8
#include "../../spim_registration_b200/csrc/kernels.h"
using namespace spim;
#ifndef PR
#define PR 8
#endif
#ifndef PINV
#define PINV 0
#endif
#ifndef PTW
#define PTW 1
#endif
#ifndef PSRC
#define PSRC 0
#endif
#ifndef PDST
#define PDST 0
#endif
#ifndef PW
#define PW 8
#endif
struct ProbeParams { FftPlanDev pl; GRows g; int s; };
extern "C" __global__ void __launch_bounds__(128) probe_stage(const __grid_constant__ ProbeParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TG tg = tg_cta();
    stage_tile<PR, PINV != 0, PTW != 0, PW>(tg, p.pl, p.s, reinterpret_cast<float4*>(smem), 0, PSRC, PDST, p.g);
}
#ifdef PMID
struct MidParams { FftPlanDev pl; GRows g; const float2* kh; long long ks4; };
extern "C" __global__ void __launch_bounds__(128) probe_mid(const __grid_constant__ MidParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TG tg = tg_cta();
    mid_tile<PR, PW>(tg, p.pl, reinterpret_cast<float4*>(smem), PSRC, PDST, p.g, p.kh, nullptr, p.ks4);
}
#endif
#ifdef PX
// x kernels: last inverse stage + fused epilogue, pre-split, first forward stage, split step
#ifndef PEPI
#define PEPI 2
#endif
#ifndef PMATH
#define PMATH 2
#endif
#ifndef PVEC
#define PVEC 1
#endif
extern "C" __global__ void __launch_bounds__(128) probe_xinv_stage0(const __grid_constant__ XInvParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* tile2 = reinterpret_cast<float2*>(smem);
    long long* srcoff = reinterpret_cast<long long*>(tile2 + (size_t)p.plan.n * TC);
    EpiAcc acc; acc.sum = 0.0; acc.mx = 0.f;
    xinv_stage0<PR, PEPI, PMATH, PVEC != 0>(p, tile2, srcoff + 2 * TC, srcoff + TC, acc);
    if (PEPI == EPI_UPDATE) stats_commit(p, tile2, acc);
}
extern "C" __global__ void __launch_bounds__(128) probe_xinv_presplit(const __grid_constant__ XInvParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* tile2 = reinterpret_cast<float2*>(smem);
    long long* srcoff = reinterpret_cast<long long*>(tile2 + (size_t)p.plan.n * TC);
    xinv_presplit(p, reinterpret_cast<float4*>(smem), srcoff, p.plan.n);
}
extern "C" __global__ void __launch_bounds__(192) probe_xfwd_stage0(const __grid_constant__ XFwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* tile2 = reinterpret_cast<float2*>(smem);
    long long* srcoff = reinterpret_cast<long long*>(tile2 + (size_t)p.plan.n * TC);
    xfwd_stage0<PR, PVEC != 0>(p, reinterpret_cast<float4*>(smem), srcoff);
}
extern "C" __global__ void __launch_bounds__(192) probe_xfwd_split(const __grid_constant__ XFwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* tile2 = reinterpret_cast<float2*>(smem);
    long long* srcoff = reinterpret_cast<long long*>(tile2 + (size_t)p.plan.n * TC);
    xfwd_split(p, reinterpret_cast<const float4*>(smem), srcoff + TC, p.plan.n);
}
#endif
