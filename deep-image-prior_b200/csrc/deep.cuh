// Persistent deep-level kernel (deep.cu): the levels whose tensors are so small that every launch is pure latency
// (>= level 2 at 512x512: 110 of 177 launches below 20 us) run as ONE kernel per pass -- a list of ops executed by all
// CTAs with grid-wide barriers in between (reference stages models/skip.py:64-91).  The ops are the very
// same device code as the stand-alone kernels (conv_tc.cu, kernels_mem.cu).
#pragma once
#include "conv_tc.cuh"
#include "kernels.cuh"

namespace dip {

enum DeepOpType {
  DO_CONV = 1, DO_WGRAD, DO_SKINNY_FWD, DO_BN_ACT_WRITE, DO_CAT_STATS, DO_CAT_WRITE, DO_BN_BWD_REDUCE, DO_BN_BWD_APPLY,
  DO_CAT_BWD_REDUCE, DO_CAT_BWD_APPLY, DO_UPADJ, DO_SKINNY_BWD
};

struct DeepSkinnyFwd { const float* x; int ldx, x_rs; const float* w; const float* b; int C, N, H, W; float* y; int mode; double* stats; int cw; };
struct DeepBnActWrite { const float* raw; int ld_in; BnRef bn; int H, W; float* dst; int ld_out, pad, act; };
struct DeepCat { CatArgs a; double* fwd; BnRef bn_cat; float* dst; };
struct DeepBnBwd { const float* raw; int ld_raw; BnRef bn; int act; GradSrc src; int H, W; double* bwd; float* draw; float* zs; double* dbias; };
struct DeepCatBwd { const float* pcat; BnRef bn_cat; const float* gp; int ld, H, W; double* bwd; float* dcat; };
struct DeepUpadj { const float* D; int ld, coff, h, w, C, bilinear; float* dst; };
struct DeepSkinnyBwd { const float* x; int ldx, x_rs; const float* w; int C, N, H, W; const float* dy; const float* out_nchw; int mode; float* dx; double* dw; double* db; int cw; };

struct alignas(128) DeepOp {
  int type;
  int sync;      // 1: grid-wide barrier after this op (0: the next op is independent of it)
  int VL, PPB;   // lane geometry of the vector kernels (as their stand-alone launchers compute it)
  int pad_[28];
  union U {
    TcConvParams conv;
    TcWgradParams wg;
    DeepSkinnyFwd skf;
    DeepBnActWrite bnw;
    DeepCat cat;
    DeepBnBwd bnb;
    DeepCatBwd catb;
    DeepUpadj up;
    DeepSkinnyBwd skb;
    U() {}
  } u;
  DeepOp() : type(0), sync(1), VL(0), PPB(0) {}
};

// dynamic shared memory available to the conv phases of the deep kernel (227 KB minus the kernel's static shared memory)
size_t deep_dyn_smem();
cudaError_t deep_kernels_init();
// ops: device array; bar: device counter (zeroed here, on the stream); grid: CTAs (<= SMs)
cudaError_t launch_deep(const DeepOp* ops, int nops, unsigned* bar, int grid, cudaStream_t s);

}  // namespace dip
