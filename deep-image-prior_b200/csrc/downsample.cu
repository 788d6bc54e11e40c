// Anti-aliased strided downsampling operator of the super-resolution path and its adjoint.
//
// Replaces models/downsampler.py:58-71 of the reference (Downsampler.forward): nn.ReplicationPad2d(pad) followed by a
// dense nn.Conv2d(n_planes, n_planes, K x K, stride=factor) whose weight is diagonal over the planes (the same K x K
// Lanczos / Gauss / box filter for every plane, models/downsampler.py:44-56) -- i.e. a depthwise stencil.  The
// reference pays C*C*K*K multiplies per output for a tensor that is zero off the diagonal; here each plane is filtered
// with the K x K taps directly.  Tensors are torch-layout planes [C][H][W] fp32 (the layout of the network output).
//
//   fwd:  y[c][oy][ox] = sum_{r,s} k[r][s] * x[c][clamp(oy*f + r - pad)][clamp(ox*f + s - pad)]
//   bwd:  dx[c][i][j]  = sum over padded positions (py,px) that clamp to (i,j) of
//                        sum_{oy,ox} dy[c][oy][ox] * k[py - oy*f][px - ox*f]          (gather form: no atomics)
//
// Both are HBM/L2-bound streaming kernels: fwd stages the (TO-1)*f+K square input patch of a 16x16 output tile in
// shared memory with coalesced row loads; bwd reads the (small) low-resolution gradient through L2.
#include "kernels.cuh"

namespace dip {

static constexpr int kDownTile = 16;   // output tile edge of the forward kernel (256 threads)

// Patch layout in shared memory: phase-split columns, s_x[col % f][row][col / f] with row pitch `pq`, so that the 16
// threads of an output row (column stride f in the image) read CONSECUTIVE words, and pq is chosen such that the two
// output rows of a warp (f patch rows apart) fall on the other half of the banks -> conflict-free inner loop.
__host__ __device__ inline int down_patch_edge(int K, int f) { return (kDownTile - 1) * f + K; }
__host__ __device__ inline int down_patch_pitch(int K, int f) {
  const int eq = (down_patch_edge(K, f) + f - 1) / f;
  for (int p = eq; p < eq + 32; ++p)
    if ((f * p) % 32 == 16) return p;
  return eq | 1;
}

__global__ void __launch_bounds__(256) k_down_fwd(const float* __restrict__ x, int H, int W, const float* __restrict__ kern,
                                                  int K, int f, int pad, float* __restrict__ y, int Ho, int Wo) {
  pdl_enter();
  extern __shared__ float ds_smem[];
  const int E = down_patch_edge(K, f), pq = down_patch_pitch(K, f);
  float* s_k = ds_smem;                      // [K*K]
  float* s_x = ds_smem + K * K;              // [f][E][pq]
  const int c = blockIdx.z;
  const int oy0 = blockIdx.y * kDownTile, ox0 = blockIdx.x * kDownTile;
  const float* xc = x + static_cast<size_t>(c) * H * W;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) s_k[i] = kern[i];
  const int y_base = oy0 * f - pad, x_base = ox0 * f - pad;
  for (int i = threadIdx.x; i < E * E; i += blockDim.x) {
    const int r = i / E, q = i - r * E;
    int sy = y_base + r, sx = x_base + q;
    sy = sy < 0 ? 0 : (sy > H - 1 ? H - 1 : sy);      // nn.ReplicationPad2d (models/downsampler.py:61-64)
    sx = sx < 0 ? 0 : (sx > W - 1 ? W - 1 : sx);
    s_x[((q % f) * E + r) * pq + q / f] = xc[static_cast<size_t>(sy) * W + sx];
  }
  __syncthreads();
  const int tx = threadIdx.x % kDownTile, ty = threadIdx.x / kDownTile;
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= Ho || ox >= Wo) return;
  float acc = 0.f;
  for (int r = 0; r < K; ++r) {
    const float* prow = s_x + (ty * f + r) * pq + tx;
    for (int q = 0; q < K; ++q)   // image column tx*f + q -> phase q % f, word tx + q / f
      acc = fmaf(s_k[r * K + q], prow[(q % f) * E * pq + q / f], acc);
  }
  y[(static_cast<size_t>(c) * Ho + oy) * Wo + ox] = acc;
}

__global__ void __launch_bounds__(256) k_down_bwd(const float* __restrict__ dy, int Ho, int Wo, const float* __restrict__ kern,
                                                  int K, int f, int pad, float* __restrict__ dx, int H, int W) {
  pdl_enter();
  extern __shared__ float ds_smem[];
  float* s_k = ds_smem;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) s_k[i] = kern[i];
  __syncthreads();
  const int c = blockIdx.z;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= H || j >= W) return;
  const float* g = dy + static_cast<size_t>(c) * Ho * Wo;
  // padded rows / columns that replicate input row i / column j
  const int py0 = i == 0 ? 0 : i + pad, py1 = i == H - 1 ? H - 1 + 2 * pad : i + pad;
  const int px0 = j == 0 ? 0 : j + pad, px1 = j == W - 1 ? W - 1 + 2 * pad : j + pad;
  float acc = 0.f;
  for (int py = py0; py <= py1; ++py) {
    int oy_lo = py - K + 1; oy_lo = oy_lo <= 0 ? 0 : (oy_lo + f - 1) / f;
    int oy_hi = py / f; if (oy_hi > Ho - 1) oy_hi = Ho - 1;
    for (int px = px0; px <= px1; ++px) {
      int ox_lo = px - K + 1; ox_lo = ox_lo <= 0 ? 0 : (ox_lo + f - 1) / f;
      int ox_hi = px / f; if (ox_hi > Wo - 1) ox_hi = Wo - 1;
      for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        const float* grow = g + static_cast<size_t>(oy) * Wo;
        const float* krow = s_k + (py - oy * f) * K + px;
        for (int ox = ox_lo; ox <= ox_hi; ++ox) acc = fmaf(grow[ox], krow[-ox * f], acc);
      }
    }
  }
  dx[(static_cast<size_t>(c) * H + i) * W + j] = acc;
}

int down_out_size(int n, int K, int f, int pad) {
  const int span = n + 2 * pad - K;
  return span < 0 ? 0 : span / f + 1;
}
size_t down_fwd_smem(int K, int f) {
  return (static_cast<size_t>(K) * K + static_cast<size_t>(f) * down_patch_edge(K, f) * down_patch_pitch(K, f)) * sizeof(float);
}

cudaError_t launch_down_fwd(const float* x, int C, int H, int W, const float* kern, int K, int f, int pad, float* y,
                            cudaStream_t s) {
  const int Ho = down_out_size(H, K, f, pad), Wo = down_out_size(W, K, f, pad);
  if (Ho < 1 || Wo < 1) return cudaErrorInvalidValue;
  const size_t smem = down_fwd_smem(K, f);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  dim3 grid((Wo + kDownTile - 1) / kDownTile, (Ho + kDownTile - 1) / kDownTile, C);
  return launch_k(k_down_fwd, grid, dim3(256), smem, s, 1, x, H, W, kern, K, f, pad, y, Ho, Wo);
}

// function attributes are per device: called from engine_init() once for every device the library is used on
cudaError_t down_kernels_init() {
  return cudaFuncSetAttribute(k_down_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_down_bwd(const float* dy, int C, int H, int W, const float* kern, int K, int f, int pad, float* dx,
                            cudaStream_t s) {
  const int Ho = down_out_size(H, K, f, pad), Wo = down_out_size(W, K, f, pad);
  if (Ho < 1 || Wo < 1) return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(K) * K * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  dim3 grid((W + 31) / 32, (H + 7) / 8, C);
  return launch_k(k_down_bwd, grid, dim3(256), smem, s, 1, dy, Ho, Wo, kern, K, f, pad, dx, H, W);
}

}  // namespace dip
