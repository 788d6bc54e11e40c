// Host-side descriptors for the tcgen05 implicit-GEMM convolution kernels (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dip {

// Implicit-GEMM forward/dgrad conv:  D[pixel][n] = sum_{tap, c} A[pixel (+) tap][c] * Wp[tap][n][c]  (+ bias[n])
//   A  : NHWC activation (fp32, or its bf16 twin when bf16 = 1) seen through a 5-D tensor map (C, px, X, py, Y)   (parity
//        dims px/py have extent 1 for stride-1 convs and 2 for stride-2 convs; out-of-bounds coordinates read as zero)
//   Wp : packed weights [tap][n_rows][c_pad] (fp32 or bf16), K-major, through a 2-D map (c_pad, taps*n_rows)
//   D  : NHWC fp32 output through a 3-D map (C_out, W_out, H_out); partial tiles and channels >= C_out are clipped by TMA.
struct TcConvParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmD;
  int tiles_x, tiles_y;    // output tile grid
  int bw, bh;              // tile = bw x bh output pixels, bw*bh == 128
  int out_w, out_h;        // valid output extent (for masked statistics)
  int kh, kw;              // filter taps
  int stride;              // 1 or 2 (spatial stride of A reads)
  int offx, offy;          // input coordinate of tap (0,0) for output pixel (0,0)
  int bf16;                // 1: bf16 operands (kind::f16): A / Wp are bf16, a K block is 64 channels (still 128 bytes), K = 16 per MMA
  int kblocks;             // K blocks (128-byte operand rows: 32 fp32 or 64 bf16 channels) per tap
  int tail_mmas;           // number of MMAs (K = 8 fp32 / 16 bf16 channels) issued for the last K block of a tap (1..4)
  int n_mma;               // UMMA N of this CTA (multiple of 16, <= 160): all output channels, or N / n_split of them
  int n_chunks;            // output 32-channel chunks written (ceil(n_mma/32))
  int stages;              // smem pipeline depth
  int patch;               // 1: 3x3 stride-1 patch mode (tile 8 x 16; one (bw+2) x (bh+2) input patch serves all taps)
  int pw, ph;              // patch extent in pixels
  int tps;                 // patch mode: filter taps per weight stage (1..3; 3 = one filter row per barrier round)
  int csize;               // thread-block cluster size (1, 2, 4): the weight tile is multicast across the cluster
  int pair;                // patch mode, n_mma == 128: every CTA iteration computes TWO vertically adjacent tiles from one
                           // (bw+2) x (2*bh+2) input patch and ONE stream of weight tiles (two accumulators per TMEM buffer):
                           // half the L2->SM weight traffic and half the barrier rounds per FLOP
  int n_split;             // 1, 2 or 4 CTAs per pixel tile, each computing n_mma = N / n_split output channels (small levels)
  // Stride-2 input gradient as its 4 sub-pixel phases in ONE launch (nphase = 4, per-tap mode): output pixel (2i+a, 2j+b)
  // of the padded input gradient only receives the taps r = a (mod 2), s = b (mod 2), i.e. a (2-a) x (2-b) stride-1
  // correlation over dY.  Work item = (phase, tile of the (h+1) x (w+1) phase grid); tmD is then the 5-D parity view
  // (C, px, X, py, Y) of the padded gradient buffer and the tile is stored at parity (opx, opy).  nphase = 0: one phase
  // described by kh / kw / offx / offy above, 3-D tmD.
  int vgrid;               // deep-level kernel only: number of CTAs that work on this conv (= grid of the stand-alone launch)
  int nphase;
  struct Phase { int kh, kw, offx, offy, tap0, opx, opy; } phs[4];   // tap0: first packed weight tap of the phase
  int n_valid;             // output channels that exist (bias / statistics are only read / written below it); 0: all n_mma * n_split
  const float* bias;       // [n_valid] or nullptr
  double* stats;           // [2][stats_ld] per-channel sum / sum of squares (fp64 atomics) or nullptr
  int stats_ld;
  int dbg_shift;           // experiment: start the A descriptor `dbg_shift` 128-byte rows into the stage
  int dbg_flags;           // experiments: 1 skip A loads, 2 skip B loads, 4 skip epilogue, 8 skip MMAs
  int dbg_nmma;            // experiment: MMAs issued per k-block (patch mode)
  int dbg_bo;              // experiment: set the descriptor's base_offset field to ((addr >> 7) & 7)
};

// Weight-gradient GEMM:  dW[tap][n][c] = sum_{pixels} dY[pixel][n] * X[pixel (+) tap][c]
//   dY : NHWC [H][W][N] (fp32 or bf16 twin; N <= 128 output channels, channels >= N read as zero) through a 3-D map (N, W, H)
//   X  : conv input through the same 5-D view as in TcConvParams
//   out: atomic = 1 (engine): every split-K CTA adds its tile into ONE fp32 accumulator [tap][128][c_pad] with vector
//        reductions at the L2 (red.global.add.v4.f32; the 0.6 MB accumulator never leaves the L2) -- no partials in
//        HBM, no reduction kernel;  atomic = 0: deterministic partials [ksplit][tap][128][c_pad] + follow-up reduce
struct TcWgradParams {
  CUtensorMap tmY;
  CUtensorMap tmX;
  float* partial;          // atomic: [kh*kw][128][c_pad] accumulator (zeroed by the caller); else [ksplits][kh*kw][128][c_pad]
  int atomic;
  int kh, kw, stride, offx, offy;
  int px_blocks_x;         // W / kp
  int px_blocks;           // total pixel blocks = H * px_blocks_x
  int kp;                  // pixels per K block (box width): 16 or 32
  int bf16;                // 1: dY / X are bf16 (chunks of 64 channels, standard 128B swizzle, K = 16 pixels per MMA)
  int c_chunks;            // 128-byte channel chunks of X (32 fp32 / 64 bf16 channels each)
  int n_cols;              // UMMA N = accumulator columns per tap = row stride of `partial` (0: 32 * c_chunks)
  int xshare;              // 1: stride-1 taps of a filter row share one (kp + kw - 1)-pixel X tile
  int ksplits;             // CTAs per tap row
  int stages;
};

size_t tc_conv_smem_bytes(const TcConvParams& p);
size_t tc_wgrad_smem_bytes(const TcWgradParams& p);
cudaError_t tc_conv_launch(const TcConvParams& p, int num_sms, cudaStream_t s);
cudaError_t tc_wgrad_launch(const TcWgradParams& p, cudaStream_t s);
int tc_conv_grid(const TcConvParams& p, int num_sms);
cudaError_t tc_kernels_init();

}  // namespace dip
