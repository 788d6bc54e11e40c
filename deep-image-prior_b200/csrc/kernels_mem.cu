// HBM-bound kernels of the dip-b200 engine: coalesced, 128-bit vectorised NHWC fp32.
//
// Each kernel replaces a chain of torch ops of the reference's skip network
// (models/skip.py:41-100 built from models/common.py:76-124):
//   input_pad        : net_input perturbation + nn.ReflectionPad2d(1)            (denoising.ipynb c10:12-13, common.py:117)
//   bn_act_write     : nn.BatchNorm2d (training mode) + nn.LeakyReLU(0.2) + nn.ReflectionPad2d(1)   (common.py:96,82,117)
//   cat_stats/write  : nn.Upsample(x2) + Concat + nn.BatchNorm2d(132) + pad     (skip.py:81,50-55; common.py:19-39)
//   bn_bwd_*/cat_bwd_*: autograd adjoints of the above
//   skinny_*         : 1x1 convs with <= 4 outputs (skip branches, RGB head + nn.Sigmoid, skip.py:57-60,96-98)
//   mse / adam / noise: torch.nn.MSELoss, torch.optim.Adam.step, noise.normal_()  (common_utils.py:225-230)
#include "kernels.cuh"

#include <math.h>

namespace dip {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4fma(float w, float4 a, float4 acc) {
  return make_float4(fmaf(w, a.x, acc.x), fmaf(w, a.y, acc.y), fmaf(w, a.z, acc.z), fmaf(w, a.w, acc.w));
}
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float lrelu(float y) { return y > 0.f ? y : kLreluSlope * y; }
__device__ __forceinline__ float4 lrelu4(float4 y) { return make_float4(lrelu(y.x), lrelu(y.y), lrelu(y.z), lrelu(y.w)); }

// per-thread BN coefficients for channels 4v..4v+3
struct Bn4 {
  float4 mean, rstd, scale, shift, gamma;
};
__device__ __forceinline__ Bn4 bn_coef(const BnRef& bn, int v) {
  float mean[4], rstd[4], g[4], b[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = 4 * v + e;
    const int ct = (c + bn.rot) % bn.C;
    const double m = bn.fwd[c] * static_cast<double>(bn.inv_n);
    double var = bn.fwd[bn.C + c] * static_cast<double>(bn.inv_n) - m * m;
    if (var < 0.0) var = 0.0;
    mean[e] = static_cast<float>(m);
    rstd[e] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
    g[e] = bn.gamma[ct];
    b[e] = bn.beta[ct];
  }
  Bn4 r;
  r.mean = make_float4(mean[0], mean[1], mean[2], mean[3]);
  r.rstd = make_float4(rstd[0], rstd[1], rstd[2], rstd[3]);
  r.gamma = make_float4(g[0], g[1], g[2], g[3]);
  r.scale = f4mul(r.gamma, r.rstd);
  r.shift = make_float4(b[0] - mean[0] * r.scale.x, b[1] - mean[1] * r.scale.y, b[2] - mean[2] * r.scale.z,
                        b[3] - mean[3] * r.scale.w);
  return r;
}
__device__ __forceinline__ float4 bn_apply(const Bn4& c, float4 x) {
  return make_float4(fmaf(x.x, c.scale.x, c.shift.x), fmaf(x.y, c.scale.y, c.shift.y), fmaf(x.z, c.scale.z, c.shift.z),
                     fmaf(x.w, c.scale.w, c.shift.w));
}
__device__ __forceinline__ float4 bn_xhat(const Bn4& c, float4 x) {
  return make_float4((x.x - c.mean.x) * c.rstd.x, (x.y - c.mean.y) * c.rstd.y, (x.z - c.mean.z) * c.rstd.z,
                     (x.w - c.mean.w) * c.rstd.w);
}

// Launch geometry for "vec-lane per 4 channels" kernels: thread = (pixel slot, v); v fixed per thread.
struct VecGeom {
  int VL, PPB, threads, blocks;
};
static VecGeom vec_geom(int C, long long npix) {
  VecGeom g;
  g.VL = C / 4;
  g.PPB = 256 / g.VL;
  if (g.PPB < 1) g.PPB = 1;
  g.threads = g.VL * g.PPB;
  long long nb = (npix + g.PPB - 1) / g.PPB;
  const long long cap = 148LL * 8;
  g.blocks = static_cast<int>(nb < cap ? nb : cap);
  if (g.blocks < 1) g.blocks = 1;
  return g;
}

// Block reduction of K float4 accumulators over the PPB pixel slots, then fp64 atomics: dst[k][4v+e].
template <int K>
__device__ __forceinline__ void block_reduce_atomic(const float4 (&acc)[K], int VL, int PPB, double* const (&dst)[K]) {
  extern __shared__ float4 red_smem[];
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
#pragma unroll
  for (int k = 0; k < K; ++k) red_smem[k * nthr + tid] = acc[k];
  __syncthreads();
  if (tid < VL) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (dst[k] == nullptr) continue;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int pp = 0; pp < PPB; ++pp) {
        const float4 t = red_smem[k * nthr + pp * VL + tid];
        s0 += t.x; s1 += t.y; s2 += t.z; s3 += t.w;
      }
      atomicAdd(dst[k] + 4 * tid + 0, s0);
      atomicAdd(dst[k] + 4 * tid + 1, s1);
      atomicAdd(dst[k] + 4 * tid + 2, s2);
      atomicAdd(dst[k] + 4 * tid + 3, s3);
    }
  }
}

// ------------------------------------------------------------------------------------------------ input_pad
__global__ void k_input_pad(const float* __restrict__ z, const float* __restrict__ noise, float sigma,
                            float* __restrict__ dst, int C, int H, int W) {
  __shared__ float tile[32][33];
  const int Wp = W + 2;
  const int yy = blockIdx.y;
  const int sy = reflect_idx(yy - 1, H);
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int xx = blockIdx.x * 32 + tx;
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k;
      float val = 0.f;
      if (xx < Wp && c < C) {
        const int sx = reflect_idx(xx - 1, W);
        const size_t off = (static_cast<size_t>(c) * H + sy) * W + sx;
        val = z[off];
        if (noise != nullptr) val = fmaf(noise[off], sigma, val);
      }
      tile[ty + 8 * k][tx] = val;
    }
    __syncthreads();
    for (int k = 0; k < 4; ++k) {
      const int xo = blockIdx.x * 32 + ty + 8 * k;
      const int c = c0 + tx;
      if (xo < Wp && c < C) dst[(static_cast<size_t>(yy) * Wp + xo) * C + c] = tile[tx][ty + 8 * k];
    }
    __syncthreads();
  }
}
void launch_input_pad(const float* z, const float* noise, float sigma, float* dst, int C, int H, int W,
                      cudaStream_t s) {
  dim3 grid((W + 2 + 31) / 32, H + 2), block(32, 8);
  k_input_pad<<<grid, block, 0, s>>>(z, noise, sigma, dst, C, H, W);
}

// ------------------------------------------------------------------------------------------------ channel_stats
__global__ void k_channel_stats(const float* __restrict__ x, int ld, int VL, int PPB, long long npix,
                                double* __restrict__ fwd, int C) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  float4 acc[2] = {f4zero(), f4zero()};
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const float4 t = ld4(x + p * ld + 4 * v);
    acc[0] = f4add(acc[0], t);
    acc[1] = f4fma(1.f, f4mul(t, t), acc[1]);
  }
  double* const dst[2] = {fwd, fwd + C};
  block_reduce_atomic<2>(acc, VL, PPB, dst);
}
void launch_channel_stats(const float* x, int ld, int C, int npix, double* fwd, cudaStream_t s) {
  VecGeom g = vec_geom(C, npix);
  k_channel_stats<<<g.blocks, g.threads, 2 * g.threads * sizeof(float4), s>>>(x, ld, g.VL, g.PPB, npix, fwd, C);
}

// ------------------------------------------------------------------------------------------------ bn_act_write
__global__ void k_bn_act_write(const float* __restrict__ raw, int ld_in, BnRef bn, int H, int W,
                               float* __restrict__ dst, int ld_out, int pad, int act, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const Bn4 cf = bn_coef(bn, v);
  const int Ho = H + 2 * pad, Wo = W + 2 * pad;
  const long long nout = static_cast<long long>(Ho) * Wo;
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < nout; p += static_cast<long long>(gridDim.x) * PPB) {
    const int yo = static_cast<int>(p / Wo), xo = static_cast<int>(p % Wo);
    const int yi = pad ? reflect_idx(yo - 1, H) : yo;
    const int xi = pad ? reflect_idx(xo - 1, W) : xo;
    float4 y = bn_apply(cf, ld4(raw + (static_cast<long long>(yi) * W + xi) * ld_in + 4 * v));
    if (act) y = lrelu4(y);
    st4(dst + p * ld_out + 4 * v, y);
  }
}
void launch_bn_act_write(const float* raw, int ld_in, BnRef bn, int H, int W, float* dst, int ld_out, int pad,
                         int act, cudaStream_t s) {
  VecGeom g = vec_geom(bn.C, static_cast<long long>(H + 2 * pad) * (W + 2 * pad));
  k_bn_act_write<<<g.blocks, g.threads, 0, s>>>(raw, ld_in, bn, H, W, dst, ld_out, pad, act, g.VL, g.PPB);
}

// ------------------------------------------------------------------------------------------------ concat stage
// Pre-BN value of the concat tensor at interior pixel (i, j) for vec lane v (v < Cu/4: upsampled, else skip branch).
__device__ __forceinline__ float4 up2x_value(const float* __restrict__ U, int Cu, int h, int w, int i, int j, int v,
                                             int bilinear) {
  if (!bilinear) return ld4(U + (static_cast<long long>(i >> 1) * w + (j >> 1)) * Cu + 4 * v);
  const int iy = i >> 1, jx = j >> 1;
  int y0, y1, x0, x1;
  float wy0, wx0;
  if (i & 1) { y0 = iy; y1 = min(iy + 1, h - 1); wy0 = 0.75f; } else { y0 = max(iy - 1, 0); y1 = iy; wy0 = 0.25f; }
  if (j & 1) { x0 = jx; x1 = min(jx + 1, w - 1); wx0 = 0.75f; } else { x0 = max(jx - 1, 0); x1 = jx; wx0 = 0.25f; }
  const float wy1 = 1.f - wy0, wx1 = 1.f - wx0;
  const float4 a = ld4(U + (static_cast<long long>(y0) * w + x0) * Cu + 4 * v);
  const float4 b = ld4(U + (static_cast<long long>(y0) * w + x1) * Cu + 4 * v);
  const float4 c = ld4(U + (static_cast<long long>(y1) * w + x0) * Cu + 4 * v);
  const float4 d = ld4(U + (static_cast<long long>(y1) * w + x1) * Cu + 4 * v);
  float4 r = f4zero();
  r = f4fma(wy0 * wx0, a, r);
  r = f4fma(wy0 * wx1, b, r);
  r = f4fma(wy1 * wx0, c, r);
  r = f4fma(wy1 * wx1, d, r);
  return r;
}
struct CatLane {
  int is_up;
  Bn4 bs;  // skip-branch BN (valid when !is_up)
};
__device__ __forceinline__ CatLane cat_lane(const CatArgs& a, int v) {
  CatLane l;
  l.is_up = v < a.Cu / 4;
  if (!l.is_up) l.bs = bn_coef(a.bn_s, v - a.Cu / 4);
  return l;
}
__device__ __forceinline__ float4 cat_value(const CatArgs& a, const CatLane& l, int i, int j, int v) {
  if (l.is_up) return up2x_value(a.U, a.Cu, a.H >> 1, a.W >> 1, i, j, v, a.bilinear);
  const float4 x = ld4(a.raw_s + (static_cast<long long>(i) * a.W + j) * a.Cs + 4 * (v - a.Cu / 4));
  return lrelu4(bn_apply(l.bs, x));
}

__global__ void k_cat_stats(CatArgs a, double* __restrict__ fwd, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const CatLane l = cat_lane(a, v);
  const long long npix = static_cast<long long>(a.H) * a.W;
  float4 acc[2] = {f4zero(), f4zero()};
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const float4 t = cat_value(a, l, static_cast<int>(p / a.W), static_cast<int>(p % a.W), v);
    acc[0] = f4add(acc[0], t);
    acc[1] = f4fma(1.f, f4mul(t, t), acc[1]);
  }
  double* const dst[2] = {fwd, fwd + (a.Cu + a.Cs)};
  block_reduce_atomic<2>(acc, VL, PPB, dst);
}
void launch_cat_stats(CatArgs a, double* fwd_cat, cudaStream_t s) {
  VecGeom g = vec_geom(a.Cu + a.Cs, static_cast<long long>(a.H) * a.W);
  k_cat_stats<<<g.blocks, g.threads, 2 * g.threads * sizeof(float4), s>>>(a, fwd_cat, g.VL, g.PPB);
}

__global__ void k_cat_write(CatArgs a, BnRef bn_cat, float* __restrict__ dst, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const CatLane l = cat_lane(a, v);
  const Bn4 cf = bn_coef(bn_cat, v);
  const int Wo = a.W + 2;
  const long long nout = static_cast<long long>(a.H + 2) * Wo;
  const int ld = a.Cu + a.Cs;
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < nout; p += static_cast<long long>(gridDim.x) * PPB) {
    const int yo = static_cast<int>(p / Wo), xo = static_cast<int>(p % Wo);
    const int i = reflect_idx(yo - 1, a.H), j = reflect_idx(xo - 1, a.W);
    st4(dst + p * ld + 4 * v, bn_apply(cf, cat_value(a, l, i, j, v)));
  }
}
void launch_cat_write(CatArgs a, BnRef bn_cat, float* dst, cudaStream_t s) {
  VecGeom g = vec_geom(a.Cu + a.Cs, static_cast<long long>(a.H + 2) * (a.W + 2));
  k_cat_write<<<g.blocks, g.threads, 0, s>>>(a, bn_cat, dst, g.VL, g.PPB);
}

// ------------------------------------------------------------------------------------------------ gradient sources
// fold: adjoint of ReflectionPad2d(1). Interior (i,j) <- padded (i+1,j+1) plus mirrored halo rows/cols.
__device__ __forceinline__ float4 fold_read(const float* __restrict__ gp, int ld, int coff, int H, int W, int i, int j,
                                            int v) {
  int rows[3], cols[3];
  int nr = 0, nc = 0;
  rows[nr++] = i + 1;
  if (i == 1) rows[nr++] = 0;
  if (i == H - 2) rows[nr++] = H + 1;
  cols[nc++] = j + 1;
  if (j == 1) cols[nc++] = 0;
  if (j == W - 2) cols[nc++] = W + 1;
  const int Wp = W + 2;
  float4 r = f4zero();
  for (int a = 0; a < nr; ++a)
    for (int b = 0; b < nc; ++b) r = f4add(r, ld4(gp + (static_cast<long long>(rows[a]) * Wp + cols[b]) * ld + coff + 4 * v));
  return r;
}
// adjoint of x2 upsampling: D is [2H][2W][ld]
__device__ __forceinline__ float4 upadj_read(const float* __restrict__ D, int ld, int coff, int H, int W, int i, int j,
                                             int v, int bilinear) {
  const int H2 = 2 * H, W2 = 2 * W;
  float4 r = f4zero();
  if (!bilinear) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        r = f4add(r, ld4(D + (static_cast<long long>(2 * i + a) * W2 + (2 * j + b)) * ld + coff + 4 * v));
    return r;
  }
  const float wgt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int y = min(max(2 * i - 1 + a, 0), H2 - 1);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int x = min(max(2 * j - 1 + b, 0), W2 - 1);
      r = f4fma(wgt[a] * wgt[b], ld4(D + (static_cast<long long>(y) * W2 + x) * ld + coff + 4 * v), r);
    }
  }
  return r;
}
template <int KIND>
__device__ __forceinline__ float4 grad_read(const GradSrc& s, int H, int W, int i, int j, int v) {
  if (KIND == 0) return ld4(s.g + (static_cast<long long>(i) * W + j) * s.ld + s.coff + 4 * v);
  if (KIND == 1) {
    float4 r = fold_read(s.g, s.ld, s.coff, H, W, i, j, v);
    if (s.g2 != nullptr) r = f4add(r, ld4(s.g2 + (static_cast<long long>(i) * W + j) * s.ld2 + 4 * v));
    return r;
  }
  return upadj_read(s.g, s.ld, s.coff, H, W, i, j, v, s.bilinear);
}
__device__ __forceinline__ float4 lrelu_bwd4(float4 y, float4 g) {
  return make_float4(y.x > 0.f ? g.x : kLreluSlope * g.x, y.y > 0.f ? g.y : kLreluSlope * g.y,
                     y.z > 0.f ? g.z : kLreluSlope * g.z, y.w > 0.f ? g.w : kLreluSlope * g.w);
}

// ------------------------------------------------------------------------------------------------ BN(+LReLU) backward
template <int KIND>
__global__ void k_bn_bwd_reduce(const float* __restrict__ raw, int ld_raw, BnRef bn, int act, GradSrc src, int H, int W,
                                double* __restrict__ bwd, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const Bn4 cf = bn_coef(bn, v);
  const long long npix = static_cast<long long>(H) * W;
  float4 acc[2] = {f4zero(), f4zero()};
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const int i = static_cast<int>(p / W), j = static_cast<int>(p % W);
    const float4 x = ld4(raw + p * ld_raw + 4 * v);
    float4 dz = grad_read<KIND>(src, H, W, i, j, v);
    if (act) dz = lrelu_bwd4(bn_apply(cf, x), dz);
    acc[0] = f4add(acc[0], dz);
    acc[1] = f4add(acc[1], f4mul(dz, bn_xhat(cf, x)));
  }
  double* const dst[2] = {bwd, bwd + bn.C};
  block_reduce_atomic<2>(acc, VL, PPB, dst);
}
void launch_bn_bwd_reduce(const float* raw, int ld_raw, BnRef bn, int act, GradSrc src, int H, int W, double* bwd,
                          cudaStream_t s) {
  VecGeom g = vec_geom(bn.C, static_cast<long long>(H) * W);
  const size_t sm = 2 * g.threads * sizeof(float4);
  if (src.kind == 0) k_bn_bwd_reduce<0><<<g.blocks, g.threads, sm, s>>>(raw, ld_raw, bn, act, src, H, W, bwd, g.VL, g.PPB);
  else if (src.kind == 1) k_bn_bwd_reduce<1><<<g.blocks, g.threads, sm, s>>>(raw, ld_raw, bn, act, src, H, W, bwd, g.VL, g.PPB);
  else k_bn_bwd_reduce<2><<<g.blocks, g.threads, sm, s>>>(raw, ld_raw, bn, act, src, H, W, bwd, g.VL, g.PPB);
}

template <int KIND>
__global__ void k_bn_bwd_apply(const float* __restrict__ raw, int ld_raw, BnRef bn, int act, GradSrc src, int H, int W,
                               const double* __restrict__ bwd, float* __restrict__ draw, float* __restrict__ zs,
                               double* __restrict__ dbias, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const Bn4 cf = bn_coef(bn, v);
  const int C = bn.C;
  float4 m1, m2;
  m1.x = static_cast<float>(bwd[4 * v + 0] * bn.inv_n); m1.y = static_cast<float>(bwd[4 * v + 1] * bn.inv_n);
  m1.z = static_cast<float>(bwd[4 * v + 2] * bn.inv_n); m1.w = static_cast<float>(bwd[4 * v + 3] * bn.inv_n);
  m2.x = static_cast<float>(bwd[C + 4 * v + 0] * bn.inv_n); m2.y = static_cast<float>(bwd[C + 4 * v + 1] * bn.inv_n);
  m2.z = static_cast<float>(bwd[C + 4 * v + 2] * bn.inv_n); m2.w = static_cast<float>(bwd[C + 4 * v + 3] * bn.inv_n);
  const long long npix = static_cast<long long>(H) * W;
  float4 acc[1] = {f4zero()};
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const int i = static_cast<int>(p / W), j = static_cast<int>(p % W);
    const float4 x = ld4(raw + p * ld_raw + 4 * v);
    float4 dz = grad_read<KIND>(src, H, W, i, j, v);
    if (act) dz = lrelu_bwd4(bn_apply(cf, x), dz);
    const float4 xh = bn_xhat(cf, x);
    float4 dx;
    dx.x = cf.scale.x * (dz.x - m1.x - xh.x * m2.x);
    dx.y = cf.scale.y * (dz.y - m1.y - xh.y * m2.y);
    dx.z = cf.scale.z * (dz.z - m1.z - xh.z * m2.z);
    dx.w = cf.scale.w * (dz.w - m1.w - xh.w * m2.w);
    st4(draw + p * C + 4 * v, dx);
    if (zs != nullptr) st4(zs + (static_cast<long long>(2 * i) * (2 * W) + 2 * j) * C + 4 * v, dx);
    acc[0] = f4add(acc[0], dx);
  }
  double* const dst[1] = {dbias};
  block_reduce_atomic<1>(acc, VL, PPB, dst);
}
void launch_bn_bwd_apply(const float* raw, int ld_raw, BnRef bn, int act, GradSrc src, int H, int W,
                         const double* bwd, float* draw, float* zs, double* dbias, cudaStream_t s) {
  VecGeom g = vec_geom(bn.C, static_cast<long long>(H) * W);
  const size_t sm = g.threads * sizeof(float4);
  if (src.kind == 0) k_bn_bwd_apply<0><<<g.blocks, g.threads, sm, s>>>(raw, ld_raw, bn, act, src, H, W, bwd, draw, zs, dbias, g.VL, g.PPB);
  else if (src.kind == 1) k_bn_bwd_apply<1><<<g.blocks, g.threads, sm, s>>>(raw, ld_raw, bn, act, src, H, W, bwd, draw, zs, dbias, g.VL, g.PPB);
  else k_bn_bwd_apply<2><<<g.blocks, g.threads, sm, s>>>(raw, ld_raw, bn, act, src, H, W, bwd, draw, zs, dbias, g.VL, g.PPB);
}

// ------------------------------------------------------------------------------------------------ concat-BN backward
__global__ void k_cat_bwd_reduce(CatArgs a, BnRef bn_cat, const float* __restrict__ gp, int ld_gp,
                                 double* __restrict__ bwd, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const CatLane l = cat_lane(a, v);
  const Bn4 cf = bn_coef(bn_cat, v);
  const long long npix = static_cast<long long>(a.H) * a.W;
  float4 acc[2] = {f4zero(), f4zero()};
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const int i = static_cast<int>(p / a.W), j = static_cast<int>(p % a.W);
    const float4 x = cat_value(a, l, i, j, v);
    const float4 dz = fold_read(gp, ld_gp, 0, a.H, a.W, i, j, v);
    acc[0] = f4add(acc[0], dz);
    acc[1] = f4add(acc[1], f4mul(dz, bn_xhat(cf, x)));
  }
  double* const dst[2] = {bwd, bwd + bn_cat.C};
  block_reduce_atomic<2>(acc, VL, PPB, dst);
}
void launch_cat_bwd_reduce(CatArgs a, BnRef bn_cat, const float* gp, int ld_gp, double* bwd, cudaStream_t s) {
  VecGeom g = vec_geom(a.Cu + a.Cs, static_cast<long long>(a.H) * a.W);
  k_cat_bwd_reduce<<<g.blocks, g.threads, 2 * g.threads * sizeof(float4), s>>>(a, bn_cat, gp, ld_gp, bwd, g.VL, g.PPB);
}
__global__ void k_cat_bwd_apply(CatArgs a, BnRef bn_cat, const float* __restrict__ gp, int ld_gp,
                                const double* __restrict__ bwd, float* __restrict__ dcat, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const CatLane l = cat_lane(a, v);
  const Bn4 cf = bn_coef(bn_cat, v);
  const int C = bn_cat.C;
  float m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    m1[e] = static_cast<float>(bwd[4 * v + e] * bn_cat.inv_n);
    m2[e] = static_cast<float>(bwd[C + 4 * v + e] * bn_cat.inv_n);
  }
  const long long npix = static_cast<long long>(a.H) * a.W;
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const int i = static_cast<int>(p / a.W), j = static_cast<int>(p % a.W);
    const float4 xh = bn_xhat(cf, cat_value(a, l, i, j, v));
    const float4 dz = fold_read(gp, ld_gp, 0, a.H, a.W, i, j, v);
    float4 dx;
    dx.x = cf.scale.x * (dz.x - m1[0] - xh.x * m2[0]);
    dx.y = cf.scale.y * (dz.y - m1[1] - xh.y * m2[1]);
    dx.z = cf.scale.z * (dz.z - m1[2] - xh.z * m2[2]);
    dx.w = cf.scale.w * (dz.w - m1[3] - xh.w * m2[3]);
    st4(dcat + p * C + 4 * v, dx);
  }
}
void launch_cat_bwd_apply(CatArgs a, BnRef bn_cat, const float* gp, int ld_gp, const double* bwd, float* dcat,
                          cudaStream_t s) {
  VecGeom g = vec_geom(a.Cu + a.Cs, static_cast<long long>(a.H) * a.W);
  k_cat_bwd_apply<<<g.blocks, g.threads, 0, s>>>(a, bn_cat, gp, ld_gp, bwd, dcat, g.VL, g.PPB);
}

// ------------------------------------------------------------------------------------------------ skinny 1x1 convs
// VL = C/4 lanes per pixel (power of two <= 32); warp-shuffle reduction over the pixel's lanes.
__global__ void k_skinny_fwd(const float* __restrict__ x, int ldx, int x_rs, const float* __restrict__ w,
                             const float* __restrict__ b, int C, int N, int H, int W, float* __restrict__ y, int mode) {
  const int VL = C / 4;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long p = gid / VL;
  const int v = static_cast<int>(gid % VL);
  const long long npix = static_cast<long long>(H) * W;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (p < npix) {
    const int i = static_cast<int>(p / W), j = static_cast<int>(p % W);
    const float4 xv = ld4(x + (static_cast<long long>(i) * x_rs + j) * ldx + 4 * v);
    for (int n = 0; n < N; ++n) {
      const float4 wv = ld4(w + n * C + 4 * v);
      acc[n] = xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
    }
  }
  for (int o = VL >> 1; o > 0; o >>= 1)
    for (int n = 0; n < 4; ++n) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (p < npix && v == 0) {
    for (int n = 0; n < N; ++n) {
      float o = acc[n] + (b != nullptr ? b[n] : 0.f);
      if (mode == 0) y[p * N + n] = o;
      else y[n * npix + p] = (mode == 1) ? 1.f / (1.f + expf(-o)) : o;
    }
  }
}
void launch_skinny_fwd(const float* x, int ldx, int x_rs, const float* w, const float* b, int C, int N, int H,
                       int W, float* y, int mode, cudaStream_t s) {
  const long long total = static_cast<long long>(H) * W * (C / 4);
  k_skinny_fwd<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(x, ldx, x_rs, w, b, C, N, H, W, y, mode);
}

__global__ void k_skinny_bwd(const float* __restrict__ x, int ldx, int x_rs, const float* __restrict__ w, int C, int N,
                             int H, int W, const float* __restrict__ dy, const float* __restrict__ out_nchw, int mode,
                             float* __restrict__ dx, double* __restrict__ dw, double* __restrict__ db, int VL, int PPB) {
  const int v = threadIdx.x % VL, slot = threadIdx.x / VL;
  const long long npix = static_cast<long long>(H) * W;
  float4 wv[4];
  for (int n = 0; n < 4; ++n) wv[n] = n < N ? ld4(w + n * C + 4 * v) : f4zero();
  float4 acc[5] = {f4zero(), f4zero(), f4zero(), f4zero(), f4zero()};  // dw rows 0..3, db (lane v == 0 only)
  for (long long p = static_cast<long long>(blockIdx.x) * PPB + slot; p < npix; p += static_cast<long long>(gridDim.x) * PPB) {
    const int i = static_cast<int>(p / W), j = static_cast<int>(p % W);
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n = 0; n < N; ++n) {
      if (mode == 0) g[n] = dy[p * N + n];
      else {
        const float d = dy[n * npix + p];
        if (mode == 1) { const float o = out_nchw[n * npix + p]; g[n] = d * o * (1.f - o); } else g[n] = d;
      }
    }
    const float4 xv = ld4(x + (static_cast<long long>(i) * x_rs + j) * ldx + 4 * v);
    float4 d = f4zero();
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      acc[n] = f4fma(g[n], xv, acc[n]);
      d = f4fma(g[n], wv[n], d);
    }
    if (v == 0) acc[4] = f4add(acc[4], make_float4(g[0], g[1], g[2], g[3]));
    if (dx != nullptr) st4(dx + p * C + 4 * v, d);
  }
  // dw[n][4v+e]
  extern __shared__ float4 red_smem[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int k = 0; k < 5; ++k) red_smem[k * nthr + tid] = acc[k];
  __syncthreads();
  if (tid < VL) {
    for (int n = 0; n < N; ++n) {
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int pp = 0; pp < PPB; ++pp) {
        const float4 t = red_smem[n * nthr + pp * VL + tid];
        s0 += t.x; s1 += t.y; s2 += t.z; s3 += t.w;
      }
      atomicAdd(dw + n * C + 4 * tid + 0, s0);
      atomicAdd(dw + n * C + 4 * tid + 1, s1);
      atomicAdd(dw + n * C + 4 * tid + 2, s2);
      atomicAdd(dw + n * C + 4 * tid + 3, s3);
    }
    if (tid == 0) {
      double s[4] = {0, 0, 0, 0};
      for (int pp = 0; pp < PPB; ++pp) {
        const float4 t = red_smem[4 * nthr + pp * VL];
        s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w;
      }
      for (int n = 0; n < N; ++n) atomicAdd(db + n, s[n]);
    }
  }
}
void launch_skinny_bwd(const float* x, int ldx, int x_rs, const float* w, int C, int N, int H, int W,
                       const float* dy, const float* out_nchw, int mode, float* dx, double* dw, double* db,
                       cudaStream_t s) {
  VecGeom g = vec_geom(C, static_cast<long long>(H) * W);
  k_skinny_bwd<<<g.blocks, g.threads, 5 * g.threads * sizeof(float4), s>>>(x, ldx, x_rs, w, C, N, H, W, dy, out_nchw,
                                                                          mode, dx, dw, db, g.VL, g.PPB);
}

// ------------------------------------------------------------------------------------------------ MSE loss
__global__ void k_mse(const float* __restrict__ out, const float* __restrict__ target, const float* __restrict__ mask,
                      int C, int HW, double* __restrict__ loss, float* __restrict__ dout) {
  const long long n = static_cast<long long>(C) * HW;
  const float inv_n = 1.f / static_cast<float>(n);
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float m = mask != nullptr ? mask[i % HW] : 1.f;
    const float d = m * (out[i] - target[i]);
    acc = fmaf(d, d, acc);
    if (dout != nullptr) dout[i] = 2.f * m * d * inv_n;
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) atomicAdd(loss, static_cast<double>(t) * static_cast<double>(inv_n));
  }
}
void launch_mse(const float* out, const float* target, const float* mask, int C, int HW, double* loss, float* dout,
                cudaStream_t s) {
  const long long n = static_cast<long long>(C) * HW;
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_mse<<<blocks, 256, 0, s>>>(out, target, mask, C, HW, loss, dout);
}

// ------------------------------------------------------------------------------------------------ Philox noise
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void k_noise(const float* __restrict__ z0, float* __restrict__ z, float sigma, uint64_t seed,
                        uint64_t offset, size_t n4) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint32_t c[4] = {static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(offset),
                     static_cast<uint32_t>(offset >> 32)};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    // Box-Muller on two pairs of uniforms in (0, 1]
    const float u0 = (static_cast<float>(c[0]) + 1.0f) * 2.3283064365386963e-10f;
    const float u1 = static_cast<float>(c[1]) * 2.3283064365386963e-10f;
    const float u2 = (static_cast<float>(c[2]) + 1.0f) * 2.3283064365386963e-10f;
    const float u3 = static_cast<float>(c[3]) * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    float4 zv = ld4(z0 + 4 * i);
    zv.x = fmaf(sigma, r0 * c0, zv.x);
    zv.y = fmaf(sigma, r0 * s0, zv.y);
    zv.z = fmaf(sigma, r1 * c1, zv.z);
    zv.w = fmaf(sigma, r1 * s1, zv.w);
    st4(z + 4 * i, zv);
  }
}
void launch_noise(const float* z0, float* z, float sigma, uint64_t seed, uint64_t offset, size_t n, cudaStream_t s) {
  const size_t n4 = n / 4;
  int blocks = static_cast<int>((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_noise<<<blocks, 256, 0, s>>>(z0, z, sigma, seed, offset, n4);
}

// ------------------------------------------------------------------------------------------------ weight packing
__global__ void k_pack_fprop(const float* __restrict__ w, int N, int C, int kh, int kw, int rot, float* __restrict__ dst,
                             int n_rows, int c_pad) {
  const long long total = static_cast<long long>(kh) * kw * n_rows * c_pad;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c_pad);
    const int n = static_cast<int>((i / c_pad) % n_rows);
    const int tap = static_cast<int>(i / (static_cast<long long>(c_pad) * n_rows));
    float val = 0.f;
    if (n < N && c < C) val = w[(static_cast<long long>(n) * C + (c + rot) % C) * (kh * kw) + tap];
    dst[i] = val;
  }
}
void launch_pack_fprop(const float* w, int N, int C, int kh, int kw, int rot, float* dst, int n_rows, int c_pad,
                       cudaStream_t s) {
  const long long total = static_cast<long long>(kh) * kw * n_rows * c_pad;
  k_pack_fprop<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(w, N, C, kh, kw, rot, dst, n_rows, c_pad);
}
__global__ void k_pack_dgrad(const float* __restrict__ w, int N, int C, int kh, int kw, int rot, float* __restrict__ dst,
                             int c_rows, int n_pad) {
  const long long total = static_cast<long long>(kh) * kw * c_rows * n_pad;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % n_pad);
    const int c = static_cast<int>((i / n_pad) % c_rows);
    const int tapf = static_cast<int>(i / (static_cast<long long>(n_pad) * c_rows));
    const int tap = kh * kw - 1 - tapf;  // (kh-1-r', kw-1-s')
    float val = 0.f;
    if (n < N && c < C) val = w[(static_cast<long long>(n) * C + (c + rot) % C) * (kh * kw) + tap];
    dst[i] = val;
  }
}
void launch_pack_dgrad(const float* w, int N, int C, int kh, int kw, int rot, float* dst, int c_rows, int n_pad,
                       cudaStream_t s) {
  const long long total = static_cast<long long>(kh) * kw * c_rows * n_pad;
  k_pack_dgrad<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(w, N, C, kh, kw, rot, dst, c_rows, n_pad);
}
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int ksplits, int N, int C, int kh, int kw, int rot,
                               int c_pad, float* __restrict__ dw) {
  const int taps = kh * kw;
  const long long total = static_cast<long long>(taps) * N * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int n = static_cast<int>((i / C) % N);
    const int tap = static_cast<int>(i / (static_cast<long long>(C) * N));
    double s = 0.0;
    for (int k = 0; k < ksplits; ++k)
      s += partial[((static_cast<long long>(k) * taps + tap) * 128 + n) * c_pad + c];
    dw[(static_cast<long long>(n) * C + (c + rot) % C) * taps + tap] = static_cast<float>(s);
  }
}
void launch_wgrad_reduce(const float* partial, int ksplits, int N, int C, int kh, int kw, int rot, int c_pad,
                         float* dw, cudaStream_t s) {
  const long long total = static_cast<long long>(kh) * kw * N * C;
  k_wgrad_reduce<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(partial, ksplits, N, C, kh, kw, rot, c_pad, dw);
}
__global__ void k_cvt_f64_f32(const double* __restrict__ src, float* __restrict__ dst, int n, int rot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(i + rot) % n] = static_cast<float>(src[i]);
}
void launch_cvt_f64_f32(const double* src, float* dst, int n, int rot, cudaStream_t s) {
  k_cvt_f64_f32<<<(n + 127) / 128, 128, 0, s>>>(src, dst, n, rot);
}
__global__ void k_bn_running(const double* __restrict__ fwd, int C, int rot, float n, float* __restrict__ rm,
                             float* __restrict__ rv, long long* __restrict__ nb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const int ct = (c + rot) % C;
    const double m = fwd[c] / n;
    double var = fwd[C + c] / n - m * m;
    if (var < 0) var = 0;
    const double unb = n > 1.f ? var * n / (n - 1.0) : var;
    rm[ct] = 0.9f * rm[ct] + 0.1f * static_cast<float>(m);
    rv[ct] = 0.9f * rv[ct] + 0.1f * static_cast<float>(unb);
  }
  if (c == 0 && nb != nullptr) *nb += 1;
}
void launch_bn_running(const double* fwd, int C, int rot, float n, float* running_mean, float* running_var,
                       long long* num_batches, cudaStream_t s) {
  k_bn_running<<<(C + 127) / 128, 128, 0, s>>>(fwd, C, rot, n, running_mean, running_var, num_batches);
}

// ------------------------------------------------------------------------------------------------ Adam
// Arithmetic order follows torch.optim.Adam (single-tensor path): m = lerp(m, g, 1-b1); v = v*b2 + (1-b2) g^2;
// p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps).
static constexpr int kAdamChunk = 2048;
__global__ void k_adam(AdamTable t, float step_size, float w1, float b2, float w2, float bc2_sqrt, float eps) {
  const int ti = t.blk_tensor[blockIdx.x];
  const int start = t.blk_start[blockIdx.x];
  const int n = t.numel[ti];
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const int end = min(start + kAdamChunk, n);
  for (int i = start + threadIdx.x; i < end; i += blockDim.x) {
    const float gi = g[i];
    float mi = m[i];
    mi = mi + (gi - mi) * w1;
    const float vi = v[i] * b2 + w2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}
void launch_adam(AdamTable t, double lr, double b1, double b2, double eps, int step, cudaStream_t s) {
  const double bc1 = 1.0 - pow(b1, step);
  const double bc2 = 1.0 - pow(b2, step);
  k_adam<<<t.nblocks, 256, 0, s>>>(t, static_cast<float>(lr / bc1), static_cast<float>(1.0 - b1),
                                   static_cast<float>(b2), static_cast<float>(1.0 - b2),
                                   static_cast<float>(sqrt(bc2)), static_cast<float>(eps));
}
int adam_chunk() { return kAdamChunk; }

}  // namespace dip
