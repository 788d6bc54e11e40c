// dip-b200 engine: plan builder (shapes -> HBM buffers, TMA tensor maps, kernel schedule) and the C ABI of
// libdip.so (include/dip.h).  Replaces the execution of the reference's skip network
// (models/skip.py:41-100, module tree interpreted by torch.nn.Sequential) + autograd backward + Adam.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>
#include <map>
#include <string>
#include <vector>

#include "../../include/dip.h"
#include "conv_tc.cuh"
#include "deep.cuh"
#include "kernels.cuh"

namespace dip {

static thread_local std::string g_err;
static int fail(const std::string& m) {
  g_err = m;
  return -1;
}
#define DIP_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
  } while (0)
#define DIP_CHECK(expr)        \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 0;
static unsigned long long g_inited_mask = 0;   // one bit per device: function attributes (dynamic smem opt-in) are per device

static int engine_init() {
  int dev = 0;
  DIP_CUDA(cudaGetDevice(&dev));
  if (dev >= 64) return fail("dip-b200: device ordinal >= 64 not supported");
  if (g_inited_mask & (1ull << dev)) return 0;
  cudaDeviceProp prop;
  DIP_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail("dip-b200 requires an sm_100 (B200) device; found sm_" + std::to_string(prop.major * 10 + prop.minor));
  g_num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  DIP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (fn == nullptr || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  DIP_CUDA(tc_kernels_init());
  DIP_CUDA(down_kernels_init());
  DIP_CUDA(deep_kernels_init());
  g_inited_mask |= 1ull << dev;
  return 0;
}

static bool is_tc(int prec) { return prec == DIP_PRECISION_TF32 || prec == DIP_PRECISION_BF16; }   // tensor-core paths

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, bool atom32 = false, bool bf16 = false) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides_b, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0],
             box[1], rank > 2 ? box[2] : 0);
    return fail(buf);
  }
  return 0;
}
// activation [rows][cols][ld] (c valid channels) as the 5-D view (C, px, X, py, Y) used by the conv kernels
// bf16 = true: the tensor holds bf16 (ld in elements); a box row is still 128 bytes = 64 channels
static int map_act5(CUtensorMap* m, const void* base, int rows, int cols, int ld, int c, int stride, int bw, int bh,
                    bool atom32 = false, bool bf16 = false) {
  const cuuint64_t e = bf16 ? 2 : sizeof(float);
  cuuint64_t dims[5], str[4];
  if (stride == 1) {
    dims[0] = c; dims[1] = 1; dims[2] = cols; dims[3] = 1; dims[4] = rows;
    str[0] = ld * e; str[1] = ld * e; str[2] = (cuuint64_t)cols * ld * e; str[3] = (cuuint64_t)cols * ld * e;
  } else {
    dims[0] = c; dims[1] = 2; dims[2] = cols / 2; dims[3] = 2; dims[4] = rows / 2;
    str[0] = ld * e; str[1] = 2 * ld * e; str[2] = (cuuint64_t)cols * ld * e; str[3] = 2 * (cuuint64_t)cols * ld * e;
  }
  cuuint32_t box[5] = {bf16 ? 64u : 32u, 1, (cuuint32_t)bw, 1, (cuuint32_t)bh};
  return encode_map(m, base, 5, dims, str, box, atom32, bf16);
}
static int map_act3(CUtensorMap* m, const void* base, int rows, int cols, int ld, int c, int bw, int bh, bool atom32 = false,
                    bool bf16 = false) {
  const cuuint64_t e = bf16 ? 2 : sizeof(float);
  cuuint64_t dims[3] = {(cuuint64_t)c, (cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[2] = {ld * e, (cuuint64_t)cols * ld * e};
  cuuint32_t box[3] = {bf16 ? 64u : 32u, (cuuint32_t)bw, (cuuint32_t)bh};
  return encode_map(m, base, 3, dims, str, box, atom32, bf16);
}
static int map_w2(CUtensorMap* m, const void* base, int rows_total, int kcols, int box_rows, bool bf16 = false) {
  const cuuint64_t e = bf16 ? 2 : sizeof(float);
  cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows_total};
  cuuint64_t str[1] = {kcols * e};
  cuuint32_t box[2] = {bf16 ? 64u : 32u, (cuuint32_t)box_rows};
  return encode_map(m, base, 2, dims, str, box, false, bf16);
}

static void pick_tile(int w, int h, int* bw, int* bh) {
  const int cand[5][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}, {8, 16}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    const long long cw = (w + cand[i][0] - 1) / cand[i][0] * cand[i][0];
    const long long ch = (h + cand[i][1] - 1) / cand[i][1] * cand[i][1];
    // prefer square-ish tiles on ties (smaller halo re-reads for 3x3 taps)
    const long long cost = cw * ch * 64 + (cand[i][0] + cand[i][1]);
    if (best < 0 || cost < best) { best = cost; *bw = cand[i][0]; *bh = cand[i][1]; }
  }
}
static int round_up(int x, int m) { return (x + m - 1) / m * m; }
// cluster size for the weight multicast: only when there is at least a full wave of tiles; the weight-tile slice of each
// CTA must be a whole number of 8-row swizzle atoms
static int pick_csize(int tiles, int n_rows) {
  int want = 1;  // measured: multicast does not pay at cluster sizes <= 4 (per-SM ingest, not L2 reads, is the limit)
  if (const char* e = getenv("DIP_CSIZE")) want = atoi(e);
  if (tiles < 2 * 148) return 1;
  while (want > 1 && (n_rows % (8 * want) != 0)) want >>= 1;
  return want < 1 ? 1 : want;
}

// ------------------------------------------------------------------------------------------------ kernel timing
// Optional CUDA-event brackets around every tensor-core launch (bench.py roofline: algorithmic FLOPs / device time).
struct Timer {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cls; double flops; size_t e0, e1; };
  std::vector<Rec> recs;
  size_t get(cudaStream_t s) {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    cudaEventRecord(pool[used], s);
    return used++;
  }
  void reset() { used = 0; recs.clear(); }
};
struct TimeScope {
  Timer* t; int cls; double flops; size_t e0 = 0; cudaStream_t s;
  TimeScope(Timer* t_, int cls_, double flops_, cudaStream_t s_) : t(t_ && t_->on ? t_ : nullptr), cls(cls_), flops(flops_), s(s_) {
    if (t) e0 = t->get(s);
  }
  ~TimeScope() { if (t) { size_t e1 = t->get(s); t->recs.push_back({cls, flops, e0, e1}); } }
};

// HBM-bound launches are recorded as class 16 + 8 * kernel id + sub-kind; the "flops" field of the record then carries the
// launch's ALGORITHMIC bytes (unique elements its contract reads + writes, x 4 B) -- bench.py's HBM rooflines.
enum HbmId { H_INPUT_PAD = 0, H_NOISE, H_SKINNY_FWD, H_BN_ACT_WRITE, H_BN_ACT_HEAD, H_CAT_STATS, H_CAT_WRITE, H_BN_BWD_REDUCE,
             H_BN_BWD_APPLY, H_CAT_BWD_REDUCE, H_CAT_BWD_APPLY, H_UPADJ, H_SKINNY_BWD, H_MSE, H_ADAM, H_HEAD_DLOGIT, H_DOWN_FWD,
             H_DOWN_BWD, H_PACK, H_WGRAD_REDUCE };
#define HBM_T(timer, id, sub, bytes, s, stmt)                                              \
  do {                                                                                     \
    TimeScope _ts((timer), 16 + 8 * (int)(id) + (int)(sub), (double)(bytes), (s));         \
    stmt;                                                                                  \
  } while (0)

// ------------------------------------------------------------------------------------------------ conv op
// Filter taps per weight stage of the patch-mode convs: a barrier round costs the MMA-issuing thread a fixed ~0.2 us, so a
// whole filter row (3 taps = 12 MMAs) per round when the stages fit, else 2 (env DIP_TPS overrides).
static int pick_tps() {
  if (const char* e = getenv("DIP_TPS")) { const int t = atoi(e); if (t >= 1 && t <= 3) return t; }
  return 3;
}
// CTAs per pixel tile (output channels split across them) for launches with fewer tiles than SMs: n_rows % (32 * split) == 0
static int pick_nsplit(int tiles, int n_rows) {
  int mx = 4;
  if (const char* e = getenv("DIP_NSPLIT_MAX")) mx = atoi(e);
  int sp = 1;
  while (sp * 2 <= mx && tiles * sp * 2 <= 148 && n_rows % (32 * sp * 2) == 0) sp *= 2;
  return sp;
}
// Tile pairs (TcConvParams::pair): 128 output channels (4 accumulators of 128 TMEM columns) or 144 (3 of 160), and only when
// the wave quantisation does not eat the gain: a pair iteration costs ~1.6 single-tile iterations (measured on the level-0
// 3x3 conv: 184.8 -> 145.9 us), so pair when ceil(pairs / SMs) * 1.6 < ceil(tiles / SMs)   (tiles_y = rows of 8 x 16 tiles)
static int pick_pair(int tiles_x, int tiles_y, int n_rows) {
  static const bool off = getenv("DIP_NO_PAIR") != nullptr;
  if (off || (n_rows != 128 && n_rows != 144)) return 0;   // 144: the 132-channel dgrad (three rotating accumulators, no statistics)
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  const int tiles = tiles_x * tiles_y, pairs = tiles_x * ((tiles_y + 1) / 2);
  return ((pairs + sms - 1) / sms) * 16 < ((tiles + sms - 1) / sms) * 10 ? 1 : 0;
}
static void fit_stages(TcConvParams& p, size_t budget = 232448) {
  for (;;) {
    p.stages = 6;
    while (tc_conv_smem_bytes(p) > budget && p.stages > 2) p.stages--;
    if (tc_conv_smem_bytes(p) <= budget || p.tps <= 1) return;
    p.tps--;  // two stages of 3 taps do not fit (wide dgrad tiles): fall back to 2 taps per stage
  }
}
struct ConvOp {
  Timer* timer = nullptr;
  double alg_flops() const { return 2.0 * out_h * out_w * (double)N * C * k * k; }
  // N output channels (a multiple of 8, <= 128: num_channels_down / num_channels_up of the level), C input channels
  int N = 128, C = 0, k = 1, stride = 1, rot = 0;
  int Np = 128;      // fprop UMMA N: N rounded up to 16 (rows per tap of the fprop pack; rows >= N are zero)
  int n_pad = 128;   // dgrad K extent per tap: N rounded up to 32 (columns of the dgrad pack), n_pad16: to 64 (bf16)
  int n_pad16 = 128;
  // An op may cover a SLICE [coff, coff + C) of the (engine-order) input channels of a wider convolution whose weight has
  // Ctot input channels (the 256-channel up conv of the skip=128 configuration: fprop runs as one op with 8 K blocks,
  // dgrad / wgrad as two 128-channel halves -- TMEM holds 512 accumulator columns).  Ctot == 0: the op is the whole conv.
  int Ctot = 0, coff = 0;
  bool do_fprop = true, do_wgrad = true;
  int dg_ld = 0;   // channel stride of dg_out (0: C)
  int c_pad = 0;   // fprop K extent per tap (multiple of 32)
  int crows = 0;   // dgrad UMMA N (input channels rounded to 16)
  // precision mode bf16: the tensor-core kernels read bf16 twins of the conv input / of dY (written by the producer kernels
  // next to -- or instead of -- the fp32 tensors) and bf16 weight packs; outputs and accumulators stay fp32
  bool bf16 = false;
  int c_pad16 = 0;                                   // fprop K extent per tap in bf16 (multiple of 64)
  const uint16_t* in16 = nullptr; int in_ld16 = 0;   // twin of `in`
  const uint16_t* dg_in16 = nullptr;                 // twin of dg_in (ld 128)
  const uint16_t* wg_dy16 = nullptr;                 // twin of wg_dy (ld 128)
  // forward
  const float* in = nullptr; int in_rows = 0, in_cols = 0, in_ld = 0; int offx = 0, offy = 0;
  float* out = nullptr; int out_h = 0, out_w = 0;
  double* stats = nullptr;
  float* wp_f = nullptr; float* wp_d = nullptr;
  float* wacc = nullptr;   // plan-owned weight-gradient accumulator [tap][128][c_pad] (tensor-core path; zeroed once per backward)
  // dgrad: dg_in [dg_in_h][dg_in_w][128] -> dg_out [dg_out_h][dg_out_w][C]
  bool has_dgrad = false;
  bool dg_s2 = false;   // tensor-core dgrad of a stride-2 3x3 conv as its 4 sub-pixel phases (dg_in = dY [h][w][128], not zero-stuffed)
  const float* dg_in = nullptr; int dg_in_h = 0, dg_in_w = 0;
  float* dg_out = nullptr; int dg_out_h = 0, dg_out_w = 0; int dg_off = 0;
  // wgrad: dY [wg_h][wg_w][128]
  const float* wg_dy = nullptr; int wg_h = 0, wg_w = 0;
  // param slots
  int p_w = -1, p_b = -1;
  TcConvParams fp{}, dg{};
  TcWgradParams wg{};
  int simt_ksplits = 1;

  void set_shapes() {
    c_pad = round_up(C, 32);
    c_pad16 = round_up(C, 64);
    crows = round_up(C, 16);
    Np = round_up(N, 16); n_pad = round_up(N, 32); n_pad16 = round_up(N, 64);
    if (Ctot == 0) Ctot = C;
    if (dg_ld == 0) dg_ld = C;
  }
  size_t wp_f_elems() const { return (size_t)k * k * Np * c_pad; }
  size_t wp_d_elems() const { return (size_t)k * k * crows * n_pad; }
  size_t wacc_elems() const { return (size_t)k * k * 128 * c_pad; }
  // pixels per K block of the weight-gradient GEMM (TMA box width): 64 where the row length allows it -- half the barrier
  // rounds per FLOP (bf16: 12 MMAs per round instead of 6; tf32: 24 instead of 12).  Measured (profiles/r02_wgrad_kp_ab.txt):
  // 512^2 denoise 376.6 -> 377.7 it/s, SR 1024^2 122.0 -> 122.6 (tf32) / 150.2 -> 150.7 (bf16).  DIP_WGRAD_KP=32|64: A/B switch.
  int wg_kp() const {
    static const int forced = getenv("DIP_WGRAD_KP") ? atoi(getenv("DIP_WGRAD_KP")) : 0;
    const bool want64 = forced ? forced == 64 : true;
    return (want64 && wg_w % 64 == 0) ? 64 : (wg_w % 32 == 0) ? 32 : 16;
  }
  int tc_ksplits() const {
    const int kp = wg_kp();
    const int blocks = wg_h * ((wg_w + kp - 1) / kp);
    int ks = 148 / k;
    if (const char* e = getenv("DIP_WGRAD_KS")) { const int cap = atoi(e); if (cap >= 1 && cap < ks) ks = cap; }   // experiment
    // every split-K CTA adds a whole [3 taps][128][c_pad] slab to the accumulator with L2 reductions: at the deep levels a CTA
    // with one or two pixel blocks costs more in reductions than in MMAs, so a CTA gets at least `minblk` pixel blocks
    static const int minblk = getenv("DIP_WGRAD_MINBLK") ? atoi(getenv("DIP_WGRAD_MINBLK")) : 1;
    if (minblk > 1 && ks > blocks / minblk) ks = blocks / minblk;
    if (ks > blocks) ks = blocks;
    return ks < 1 ? 1 : ks;
  }
  size_t partial_elems(int prec) const {
    const int ks = is_tc(prec) ? 1 : simt_ksplits;   // tensor-core path: one accumulator (atomic split-K)
    return (size_t)ks * k * k * 128 * c_pad;
  }

  int build_tc(float* partial) {
    // ---- fprop
    int bw, bh;
    pick_tile(out_w, out_h, &bw, &bh);
    fp = TcConvParams{};
    const bool patch_ok = getenv("DIP_NO_PATCH") == nullptr;
    if (!do_fprop) {
    } else if (k == 3 && stride == 1 && patch_ok) {
      // patch mode: tile 8 wide x 16 tall, one 10 x 18 input patch per 32-channel block feeds all nine taps
      bw = 8; bh = 16;
      fp.patch = 1; fp.pw = bw + 2; fp.ph = bh + 2;
      fp.pair = pick_pair((out_w + bw - 1) / bw, (out_h + bh - 1) / bh, Np);
      if (fp.pair) fp.ph = 2 * bh + 2;
      DIP_CHECK(bf16 ? map_act5(&fp.tmA, in16, in_rows, in_cols, in_ld16, C, 1, fp.pw, fp.ph, false, true)
                     : map_act5(&fp.tmA, in, in_rows, in_cols, in_ld, C, 1, fp.pw, fp.ph));
    } else {
      DIP_CHECK(bf16 ? map_act5(&fp.tmA, in16, in_rows, in_cols, in_ld16, C, stride, bw, bh, false, true)
                     : map_act5(&fp.tmA, in, in_rows, in_cols, in_ld, C, stride, bw, bh));
    }
    if (do_fprop) {
    fp.csize = fp.pair ? 1 : pick_csize(((out_w + bw - 1) / bw) * ((out_h + bh - 1) / bh), Np);
    fp.tps = (fp.patch && fp.csize == 1) ? pick_tps() : 1;
    fp.n_split = fp.pair ? 1 : fp.csize == 1 ? pick_nsplit(((out_w + bw - 1) / bw) * ((out_h + bh - 1) / bh), Np) : 1;
    DIP_CHECK(map_w2(&fp.tmB, wp_f, k * k * Np, bf16 ? c_pad16 : c_pad, Np / fp.csize / fp.n_split, bf16));
    DIP_CHECK(map_act3(&fp.tmD, out, out_h, out_w, N, N, bw, bh));
    fp.tiles_x = (out_w + bw - 1) / bw; fp.tiles_y = (out_h + bh - 1) / bh;
    fp.bw = bw; fp.bh = bh; fp.out_w = out_w; fp.out_h = out_h;
    fp.kh = fp.kw = k; fp.stride = stride; fp.offx = offx; fp.offy = offy;
    fp.bf16 = bf16 ? 1 : 0;
    fp.kblocks = bf16 ? c_pad16 / 64 : c_pad / 32;
    fp.tail_mmas = bf16 ? ((C % 64 == 0) ? 4 : (C % 64 + 15) / 16) : ((C % 32 == 0) ? 4 : (C % 32 + 7) / 8);
    fp.n_mma = Np / fp.n_split; fp.n_chunks = (fp.n_mma + 31) / 32;
    fp.n_valid = N;
    fp.bias = nullptr; fp.stats = stats; fp.stats_ld = N;
    fit_stages(fp);
    }
    // ---- dgrad
    if (has_dgrad && dg_s2) {
      // phase grid: (dg_out_h / 2) x (dg_out_w / 2) positions per parity class of the padded gradient
      const int gh = dg_out_h / 2, gw = dg_out_w / 2;
      pick_tile(gw, gh, &bw, &bh);
      dg = TcConvParams{};
      DIP_CHECK(bf16 ? map_act5(&dg.tmA, dg_in16, dg_in_h, dg_in_w, N, N, 1, bw, bh, false, true)
                     : map_act5(&dg.tmA, dg_in, dg_in_h, dg_in_w, N, N, 1, bw, bh));
      dg.csize = 1; dg.tps = 1;
      dg.n_split = pick_nsplit(4 * ((gw + bw - 1) / bw) * ((gh + bh - 1) / bh), crows);
      DIP_CHECK(map_w2(&dg.tmB, wp_d, k * k * crows, bf16 ? n_pad16 : n_pad, crows / dg.n_split, bf16));
      DIP_CHECK(map_act5(&dg.tmD, dg_out, dg_out_h, dg_out_w, dg_ld, C, 2, bw, bh));   // parity view of the padded gradient
      dg.tiles_x = (gw + bw - 1) / bw; dg.tiles_y = (gh + bh - 1) / bh;
      dg.bw = bw; dg.bh = bh; dg.out_w = gw; dg.out_h = gh;
      dg.kh = dg.kw = 2; dg.stride = 1; dg.offx = dg.offy = -1;
      dg.nphase = 4;
      int t0 = 0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          TcConvParams::Phase& q = dg.phs[a * 2 + b];
          q.kh = 2 - a; q.kw = 2 - b; q.offy = a == 0 ? -1 : 0; q.offx = b == 0 ? -1 : 0; q.tap0 = t0; q.opx = b; q.opy = a;
          t0 += q.kh * q.kw;
        }
      dg.bf16 = bf16 ? 1 : 0;
      dg.kblocks = bf16 ? n_pad16 / 64 : n_pad / 32;   // K = the N channels of dY
      dg.tail_mmas = bf16 ? ((N % 64 == 0) ? 4 : (N % 64 + 15) / 16) : ((N % 32 == 0) ? 4 : (N % 32 + 7) / 8);
      dg.n_mma = crows / dg.n_split; dg.n_chunks = (dg.n_mma + 31) / 32;
      dg.bias = nullptr; dg.stats = nullptr; dg.stats_ld = 0;
      fit_stages(dg);
    } else if (has_dgrad) {
      pick_tile(dg_out_w, dg_out_h, &bw, &bh);
      dg = TcConvParams{};
      if (k == 3 && patch_ok) {
        bw = 8; bh = 16;
        dg.patch = 1; dg.pw = bw + 2; dg.ph = bh + 2;
        dg.pair = pick_pair((dg_out_w + bw - 1) / bw, (dg_out_h + bh - 1) / bh, crows);
        if (dg.pair) dg.ph = 2 * bh + 2;
        DIP_CHECK(bf16 ? map_act5(&dg.tmA, dg_in16, dg_in_h, dg_in_w, N, N, 1, dg.pw, dg.ph, false, true)
                       : map_act5(&dg.tmA, dg_in, dg_in_h, dg_in_w, N, N, 1, dg.pw, dg.ph));
      } else {
        DIP_CHECK(bf16 ? map_act5(&dg.tmA, dg_in16, dg_in_h, dg_in_w, N, N, 1, bw, bh, false, true)
                       : map_act5(&dg.tmA, dg_in, dg_in_h, dg_in_w, N, N, 1, bw, bh));
      }
      dg.csize = dg.pair ? 1 : pick_csize(((dg_out_w + bw - 1) / bw) * ((dg_out_h + bh - 1) / bh), crows);
      dg.tps = (dg.patch && dg.csize == 1) ? pick_tps() : 1;
      dg.n_split = dg.pair ? 1 : dg.csize == 1 ? pick_nsplit(((dg_out_w + bw - 1) / bw) * ((dg_out_h + bh - 1) / bh), crows) : 1;
      DIP_CHECK(map_w2(&dg.tmB, wp_d, k * k * crows, bf16 ? n_pad16 : n_pad, crows / dg.csize / dg.n_split, bf16));
      DIP_CHECK(map_act3(&dg.tmD, dg_out, dg_out_h, dg_out_w, dg_ld, C, bw, bh));
      dg.tiles_x = (dg_out_w + bw - 1) / bw; dg.tiles_y = (dg_out_h + bh - 1) / bh;
      dg.bw = bw; dg.bh = bh; dg.out_w = dg_out_w; dg.out_h = dg_out_h;
      dg.kh = dg.kw = k; dg.stride = 1; dg.offx = dg.offy = dg_off;
      dg.bf16 = bf16 ? 1 : 0;
      dg.kblocks = bf16 ? n_pad16 / 64 : n_pad / 32;   // K = the N channels of dY
      dg.tail_mmas = bf16 ? ((N % 64 == 0) ? 4 : (N % 64 + 15) / 16) : ((N % 32 == 0) ? 4 : (N % 32 + 7) / 8);
      dg.n_mma = crows / dg.n_split; dg.n_chunks = (dg.n_mma + 31) / 32;
      dg.bias = nullptr; dg.stats = nullptr; dg.stats_ld = 0;
      fit_stages(dg);
    }
    // ---- wgrad
    wg = TcWgradParams{};
    if (!do_wgrad) return 0;
    wg.kp = wg_kp();
    wg.bf16 = bf16 ? 1 : 0;
    wg.xshare = (stride == 1 && k == 3 && getenv("DIP_NO_XSHARE") == nullptr) ? 1 : 0;
    if (bf16) {
      DIP_CHECK(map_act3(&wg.tmY, wg_dy16, wg_h, wg_w, N, N, wg.kp, 1, false, true));
      DIP_CHECK(map_act5(&wg.tmX, in16, in_rows, in_cols, in_ld16, C, stride, wg.xshare ? wg.kp + k - 1 : wg.kp, 1, false, true));
    } else {
      DIP_CHECK(map_act3(&wg.tmY, wg_dy, wg_h, wg_w, N, N, wg.kp, 1, true));
      DIP_CHECK(map_act5(&wg.tmX, in, in_rows, in_cols, in_ld, C, stride, wg.xshare ? wg.kp + k - 1 : wg.kp, 1, true));
    }
    wg.partial = partial;
    wg.kh = wg.kw = k; wg.stride = stride; wg.offx = offx; wg.offy = offy;
    wg.px_blocks_x = (wg_w + wg.kp - 1) / wg.kp;
    wg.px_blocks = wg_h * wg.px_blocks_x;
    wg.c_chunks = bf16 ? c_pad16 / 64 : c_pad / 32;
    wg.n_cols = c_pad;
    wg.ksplits = tc_ksplits();
    wg.stages = 6;
    while (tc_wgrad_smem_bytes(wg) > 232448 && wg.stages > 1) wg.stages--;
    if (tc_wgrad_smem_bytes(wg) > 232448) return fail("wgrad stage does not fit in shared memory");
    return 0;
  }

  int run_fprop(int prec, const float* bias, cudaStream_t s) {
    if (is_tc(prec)) {
      TcConvParams p = fp;
      p.bias = bias;
      if (const char* e = getenv("DIP_DBG_SHIFT")) p.dbg_shift = atoi(e);
      if (const char* e = getenv("DIP_DBG_BO")) p.dbg_bo = atoi(e);
      if (const char* e = getenv("DIP_DBG_FLAGS")) p.dbg_flags = atoi(e);
      if (const char* e = getenv("DIP_DBG_NMMA")) p.dbg_nmma = atoi(e);
      if (const char* e = getenv("DIP_DBG_STAGES")) { const int st = atoi(e); if (st >= 1 && st < p.stages) p.stages = st; }
      TimeScope ts(timer, 0, alg_flops(), s);
      DIP_CUDA(tc_conv_launch(p, g_num_sms, s));
    } else {
      SimtConvArgs a{};
      a.A = in; a.a_h = in_rows; a.a_w = in_cols; a.a_ld = in_ld; a.a_c = C;
      a.Wp = wp_f; a.n_rows = Np; a.c_pad = c_pad;
      a.D = out; a.d_h = out_h; a.d_w = out_w; a.d_ld = N; a.d_c = N;
      a.kh = a.kw = k; a.stride = stride; a.offx = offx; a.offy = offy; a.bias = bias;
      launch_simt_conv(a, s);
      if (stats != nullptr) launch_channel_stats(out, N, N, out_h * out_w, stats, s);
      DIP_CUDA(cudaGetLastError());
    }
    return 0;
  }
  int run_dgrad(int prec, cudaStream_t s) {
    if (is_tc(prec)) {
      TimeScope ts(timer, 1, alg_flops(), s);
      DIP_CUDA(tc_conv_launch(dg, g_num_sms, s));
    } else {
      SimtConvArgs a{};
      a.A = dg_in; a.a_h = dg_in_h; a.a_w = dg_in_w; a.a_ld = N; a.a_c = N;
      a.Wp = wp_d; a.n_rows = crows; a.c_pad = n_pad;
      a.D = dg_out; a.d_h = dg_out_h; a.d_w = dg_out_w; a.d_ld = dg_ld; a.d_c = C;
      a.kh = a.kw = k; a.stride = 1; a.offx = a.offy = dg_off; a.bias = nullptr;
      launch_simt_conv(a, s);
      DIP_CUDA(cudaGetLastError());
    }
    return 0;
  }
  int run_wgrad(int prec, float* partial, float* dw, cudaStream_t s) {
    static const bool dbg_skip = getenv("DIP_DBG_SKIP_WGRAD") != nullptr;   // timing diagnostic only: gradients are wrong
    if (dbg_skip) return 0;
    int ks;
    if (is_tc(prec)) {
      // split-K CTAs add their tiles into one accumulator with vector reductions at the L2.  Plan-owned accumulator
      // (wacc): zeroed by one memset per backward, unpacked to OIHW by one table kernel at the end of the backward pass.
      // Single-op entry points: zero + launch + unpack here.
      TcWgradParams p = wg;
      p.atomic = 1;
      p.partial = wacc != nullptr ? wacc : partial;
      if (wacc == nullptr) DIP_CUDA(cudaMemsetAsync(partial, 0, wacc_elems() * sizeof(float), s));
      {
        TimeScope ts(timer, 2, alg_flops(), s);
        DIP_CUDA(tc_wgrad_launch(p, s));
      }
      if (wacc != nullptr) return 0;
      launch_wgrad_reduce(partial, 1, N, C, k, k, rot, c_pad, dw, s, Ctot, coff);
      DIP_CUDA(cudaGetLastError());
      return 0;
    } else {
      SimtWgradArgs a{};
      a.dY = wg_dy; a.h = wg_h; a.w = wg_w; a.dy_ld = N; a.n = N;
      a.X = in; a.x_h = in_rows; a.x_w = in_cols; a.x_ld = in_ld; a.x_c = C;
      a.kh = a.kw = k; a.stride = stride; a.offx = offx; a.offy = offy;
      a.partial = partial; a.c_pad = c_pad; a.ksplits = ks = simt_ksplits;
      launch_simt_wgrad(a, s);
    }
    HBM_T(timer, H_WGRAD_REDUCE, 0, ((double)ks + 1.0) * k * k * 128.0 * c_pad * sizeof(float), s,
          launch_wgrad_reduce(partial, ks, N, C, k, k, rot, c_pad, dw, s, Ctot, coff));
    DIP_CUDA(cudaGetLastError());
    return 0;
  }
};

// ------------------------------------------------------------------------------------------------ table kernels
struct PackEntry {
  const float* w; float* dst_f; float* dst_d;
  int N, C, k, rot, n_rows, c_pad, c_rows;
  int Ctot, coff;   // weight has Ctot input channels; this entry packs engine channels [coff, coff + C)
  int s2;           // dgrad pack of a stride-2 3x3 conv: taps in sub-pixel phase order (kS2Taps), not flipped
  int bf16;         // packs hold bf16 (same buffers); the fprop pack then has rows of c_pad16 (multiple of 64) channels
  int c_pad16;
  int n_pad;        // columns of the dgrad pack: N rounded up to 32 (bf16: to 64)
};
// packed tap t of the 4-phase stride-2 dgrad -> filter tap r * 3 + s.  Phase (a, b) = parity of the padded gradient pixel;
// its taps are r in {2, 0} (a = 0: dY rows i-1, i) or {1} (a = 1), same for s.  Phases in the order (0,0) (0,1) (1,0) (1,1).
__constant__ int kS2Taps[9] = {2 * 3 + 2, 2 * 3 + 0, 0 * 3 + 2, 0 * 3 + 0,   // (0,0): (r', s') = (0,0) (0,1) (1,0) (1,1)
                               2 * 3 + 1, 0 * 3 + 1,                           // (0,1): s = 1
                               1 * 3 + 2, 1 * 3 + 0,                           // (1,0): r = 1
                               1 * 3 + 1};                                     // (1,1)
__global__ void k_pack_table(const PackEntry* __restrict__ tab) {
  pdl_enter();
  const PackEntry e = tab[blockIdx.y];
  const int taps = e.k * e.k;
  const int c_pad = e.bf16 ? e.c_pad16 : e.c_pad;
  __nv_bfloat16* const f16 = reinterpret_cast<__nv_bfloat16*>(e.dst_f);
  __nv_bfloat16* const d16 = reinterpret_cast<__nv_bfloat16*>(e.dst_d);
  const long long nf = e.dst_f != nullptr ? (long long)taps * e.n_rows * c_pad : 0;
  const int n_pad = e.n_pad > 0 ? e.n_pad : 128;
  const long long nd = e.dst_d != nullptr ? (long long)taps * e.c_rows * n_pad : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nd; i += (long long)gridDim.x * blockDim.x) {
    if (i < nf) {
      const int c = (int)(i % c_pad), n = (int)((i / c_pad) % e.n_rows), tap = (int)(i / ((long long)c_pad * e.n_rows));
      float v = 0.f;
      if (n < e.N && c < e.C) v = e.w[((long long)n * e.Ctot + (c + e.coff + e.rot) % e.Ctot) * taps + tap];
      if (e.bf16) f16[i] = __float2bfloat16_rn(v); else e.dst_f[i] = v;
    } else {
      const long long j = i - nf;
      const int n = (int)(j % n_pad), c = (int)((j / n_pad) % e.c_rows), tapf = (int)(j / ((long long)n_pad * e.c_rows));
      const int tap = e.s2 ? kS2Taps[tapf] : taps - 1 - tapf;
      float v = 0.f;
      if (n < e.N && c < e.C) v = e.w[((long long)n * e.Ctot + (c + e.coff + e.rot) % e.Ctot) * taps + tap];
      if (e.bf16) d16[j] = __float2bfloat16_rn(v); else e.dst_d[j] = v;
    }
  }
}
// accumulators [tap][128][c_pad] of all tensor-core weight gradients -> OIHW gradients (one launch per backward pass)
struct UnpackEntry {
  const float* acc; float* dw;
  int N, C, taps, rot, c_pad, Ctot, coff;
};
__global__ void k_wgrad_unpack_table(const UnpackEntry* __restrict__ tab) {
  pdl_enter();
  const UnpackEntry e = tab[blockIdx.y];
  const int total = e.N * e.C * e.taps;   // dw elements this entry owns: (n, engine channel c, tap)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % e.taps, c = (i / e.taps) % e.C, n = i / (e.taps * e.C);
    e.dw[((size_t)n * e.Ctot + (c + e.coff + e.rot) % e.Ctot) * e.taps + tap] = e.acc[((size_t)tap * 128 + n) * e.c_pad + c];
  }
}
struct CvtEntry {
  const double* src; float* dst; int n, rot;
  int row, row_ld;   // dst rows of `row` elements come from accumulator rows of `row_ld` (0: contiguous)
};
__global__ void k_cvt_table(const CvtEntry* __restrict__ tab) {
  pdl_enter();
  const CvtEntry e = tab[blockIdx.x];
  for (int i = threadIdx.x; i < e.n; i += blockDim.x) {
    const int si = e.row > 0 ? (i / e.row) * e.row_ld + i % e.row : i;
    e.dst[(i + e.rot) % e.n] = (float)acc_get(e.src + (size_t)si * kAccS);
  }
}
struct RunEntry {
  const double* fwd; float* rm; float* rv; void* nb; int C, rot; float n; int nb_is_float;
};
__global__ void k_running_table(const RunEntry* __restrict__ tab) {
  pdl_enter();
  const RunEntry e = tab[blockIdx.x];
  if (e.rm == nullptr) return;
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    const int ct = (c + e.rot) % e.C;
    const double m = acc_get(e.fwd + c * kAccS) / e.n;
    double var = acc_get(e.fwd + (e.C + c) * kAccS) / e.n - m * m;
    if (var < 0) var = 0;
    const double unb = e.n > 1.f ? var * e.n / (e.n - 1.0) : var;
    e.rm[ct] = 0.9f * e.rm[ct] + 0.1f * (float)m;
    e.rv[ct] = 0.9f * e.rv[ct] + 0.1f * (float)unb;
  }
  if (threadIdx.x == 0 && e.nb != nullptr) {
    // Module.type(torch.cuda.FloatTensor) (every notebook does this) also casts num_batches_tracked to float32
    if (e.nb_is_float) *reinterpret_cast<float*>(e.nb) += 1.f; else *reinterpret_cast<long long*>(e.nb) += 1;
  }
}

// ------------------------------------------------------------------------------------------------ plan
struct Arena {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* get(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};
struct BufInfo {
  void* ptr; int rows, cols, ld, c;
};
struct BnLayer {
  int C = 0, rot = 0; float n = 1.f;
  double* fwd = nullptr; double* bwd = nullptr; double* dbias = nullptr;
  int p_gamma = -1, p_beta = -1, p_bias = -1, idx = -1;
};
struct Level {
  int H, W, h, w, Cin;   // Cin: stored depth of the level input (power of two >= 4)
  int Cin_act;           // depth of the conv weights that read it (differs at level 0 for input depths like 3)
  int bilinear;          // x2 upsampling into this level: 1 bilinear, 0 nearest (skip.py:81, upsample_mode[i])
  // channel widths (models/skip.py:5-7): nd = num_channels_down[l] (both down convs), nu = num_channels_up[l] (up conv and 1x1),
  // ns = num_channels_skip[l], cu = depth of the tensor upsampled into this level's concat (nu of the level below, or nd at
  // the deepest level).  128 / 128 / {0, 4, 128} / 128 in every BASELINE configuration.
  int nd = 128, nu = 128, ns = 0, cu = 128;
  // downsample_mode 'avg': the first down conv runs at stride 1 into rawF [H][W][nd], raw_d1 = AvgPool2d(2, 2)(rawF);
  // dRawF [H][W][nd] = the pooling adjoint of dRaw_d1 (input of that conv's dgrad / wgrad)
  float *rawF = nullptr, *dRawF = nullptr;
  uint16_t* dRawF16 = nullptr;
  float *Pin, *raw_s, *raw_d1, *P_d1, *raw_d2, *P_d2, *P_cat, *raw_u, *A_u, *raw_v, *U;
  // bf16 twins (precision mode bf16; ld = the fp32 tensor's depth rounded up to 8)
  uint16_t *Pin16 = nullptr, *P_d1_16 = nullptr, *P_d2_16 = nullptr, *P_cat16 = nullptr, *A_u16 = nullptr;
  uint16_t *dRaw_v16 = nullptr, *dRaw_u16 = nullptr, *dRaw_d2_16 = nullptr, *dRaw_d1_16 = nullptr, *dRaw_s16 = nullptr;
  int Pin_ld16 = 0, cat_ld16 = 0;
  float *dUp;  // [h][w][128] adjoint of the upsampling applied to dCat
  float *dS;   // skip=128: [H][W][Cin] input gradient of the (tensor-core) skip conv, levels > 0
  float *dRaw_v, *dA_u, *dRaw_u, *dP_cat, *dCat, *dRaw_s, *dRaw_d2, *dP_d1, *dRaw_d1, *ZS, *dPin;
  BnLayer bn_s, bn_d1, bn_d2, bn_cat, bn_u, bn_v;
  int p_skip_w, p_skip_b;
  double* dw_s;
  ConvOp d1, d2, up, c11;
  ConvOp sk, up_a, up_b;   // skip=128 only: 1x1 skip conv Cin -> 128; channel halves of the 256 -> 128 up conv (dgrad / wgrad)
};

}  // namespace dip

using namespace dip;

struct dip_plan {
  dip_net_desc desc;
  int H, W;
  bool dry = false;
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<Level> lv;
  std::vector<long long> numel;
  std::vector<float*> params, grads;
  std::vector<void*> running;
  std::vector<BnLayer*> bns;
  std::vector<ConvOp*> convs;
  std::map<std::string, BufInfo> bufs;
  // head
  int p_head_w = -1, p_head_b = -1;
  double* dw_head = nullptr; double* db_head = nullptr; double* db_scratch = nullptr;
  float* out_saved = nullptr;  // [C_out][H][W] (sigmoid output, needed by backward)
  // accumulators
  double* acc_fwd = nullptr; size_t acc_fwd_n = 0;
  double* acc_bwd = nullptr; size_t acc_bwd_n = 0;
  float* partial = nullptr;
  // runner scratch
  float* zbuf = nullptr; float* dout = nullptr; float* dl4 = nullptr;
  // super-resolution operator applied between the network output and the loss (dip_plan_set_downsampler)
  static constexpr int kDownMaxK = 64;
  float* ds_kern = nullptr;      // [K][K] taps
  float* ds_y = nullptr;         // [C_out][Ho][Wo] downsampled output
  float* ds_dy = nullptr;        // [C_out][Ho][Wo] dL/d(ds_y)
  int ds_K = 0, ds_f = 1, ds_pad = 0, ds_Ho = 0, ds_Wo = 0;
  static constexpr int kLossRing = 65536;
  double* loss_ring = nullptr;   // [kLossRing] loss slots of the runner when the caller passes no history buffer
  int* it_dev = nullptr;         // [2] device counters: {global Adam step, iteration index of this call}
  // identity of everything the captured step bakes in; compared field by field (the struct has padding).  adam_id is a
  // process-wide serial number: a new dip_adam allocated at the address of a destroyed one must not match.
  struct GraphKey {
    const void *z0 = nullptr, *target = nullptr, *mask = nullptr, *out = nullptr, *slots = nullptr;
    unsigned long long adam_id = 0, adam_bind = 0;
    float sigma = 0.f; uint64_t seed = 0; double lr = 0.0;
    bool operator==(const GraphKey& o) const {
      return z0 == o.z0 && target == o.target && mask == o.mask && out == o.out && slots == o.slots && adam_id == o.adam_id &&
             adam_bind == o.adam_bind && sigma == o.sigma && seed == o.seed && lr == o.lr;
    }
  };
  GraphKey gkey{};
  cudaGraphExec_t gexec = nullptr;
  // notebook path (dip_forward / dip_backward called once per closure): each is replayed as its own CUDA graph over the
  // plan's fixed staging buffers (zbuf / out_saved / dout); rebuilt when the bound pointers change
  cudaGraphExec_t gfwd = nullptr, gbwd = nullptr;
  cudaStream_t gstream = nullptr;
  cudaEvent_t gev_in = nullptr, gev_out = nullptr;
  // weight-gradient chain runs on a private side stream, forked/joined with events (also inside graph capture)
  cudaStream_t wstream = nullptr;
  // the skip branches (1x1 conv + BN of every level, forward and backward) are independent of the deeper-level chain:
  // they run on a second side stream
  cudaStream_t sstream = nullptr;
  std::vector<cudaEvent_t> wev;
  size_t wev_used = 0;
  bool side_on = false;
  // weight-gradient GEMMs of the outer levels deferred until the main chain is inside the (latency-bound, SM-starved)
  // deep levels, where a full-GPU tensor-core kernel on the side stream costs the least
  std::vector<dip::ConvOp*> deferred;
  // runner: the input perturbation is generated inside the level-0 input transform (k_noise_pad) -- set around plan_forward
  struct { bool on = false; const float* z0 = nullptr; float sigma = 0.f; uint64_t seed = 0, offset = 0; const int* it_dev = nullptr; } fnoise;
  bool prepacked = false;   // the runner already issued the weight repack of this forward (beside the noise kernel)
  // tables
  PackEntry* d_pack = nullptr; CvtEntry* d_cvt = nullptr; RunEntry* d_run = nullptr; UnpackEntry* d_unpack = nullptr;
  int n_pack = 0, n_cvt = 0, n_run = 0, n_unpack = 0;
  float* wacc_base = nullptr; size_t wacc_bytes = 0;   // all weight-gradient accumulators, contiguous
  // persistent deep-level kernel (deep.cu): levels >= deep_from run as one launch per pass
  static constexpr int kDeepMaxOps = 96;
  DeepOp* d_deep_fwd = nullptr; DeepOp* d_deep_bwd = nullptr;
  int n_deep_fwd = 0, n_deep_bwd = 0, deep_grid = 0, deep_from = 2;
  unsigned* deep_bar = nullptr;
  long long pack_max = 0;
  bool bound = false;
  int nbt_is_float = 0;
  int launches_fwd = 0, launches_bwd = 0;
  Timer timer;
};

namespace dip {

static BnRef bn_ref(const dip_plan* P, const BnLayer& b) {
  BnRef r;
  r.fwd = b.fwd; r.gamma = P->params[b.p_gamma]; r.beta = P->params[b.p_beta];
  r.C = b.C; r.rot = b.rot; r.inv_n = 1.f / b.n;
  return r;
}

static int build_plan(dip_plan* P, Arena& A) {
  const dip_net_desc& d = P->desc;
  const int L = d.num_scales;
  if (L < 1 || L > 8) return fail("dip-b200: 1..8 scales supported");
  // widths: one value for every scale (channels / skip_channels), or per scale (channels == 0: channels_down / channels_up /
  // channels_skip, denoising.ipynb c8:17-23 "snail": [8, 16, 32, 64, 128] with skips [0, 0, 0, 4, 4])
  std::vector<int> ND(L), NU(L), NS(L);
  for (int l = 0; l < L; ++l) {
    ND[l] = d.channels > 0 ? d.channels : d.channels_down[l];
    NU[l] = d.channels > 0 ? d.channels : d.channels_up[l];
    NS[l] = d.channels > 0 ? d.skip_channels : d.channels_skip[l];
    for (int w : {ND[l], NU[l]})
      if (w < 8 || w > 128 || w % 8 != 0)
        return fail("dip-b200: num_channels_down / num_channels_up must be multiples of 8 in [8, 128] (128 in the BASELINE configurations)");
  }
  bool wide = true, uniform = true;
  for (int l = 0; l < L; ++l) {
    wide = wide && NS[l] == 128 && ND[l] == 128 && NU[l] == 128;
    uniform = uniform && ND[l] == 128 && NU[l] == 128 && NS[l] == NS[0];
  }
  for (int l = 0; l < L; ++l)
    if (!(wide || NS[l] == 0 || NS[l] == 4))
      return fail("dip-b200: num_channels_skip must be 0 or 4 per scale, or 128 at every scale of a 128-wide network");
  const int CS = NS[0];   // the uniform skip width where the code below asks for it (wide: 128)
  if (d.downsample_mode != 0 && d.downsample_mode != 1) return fail("dip-b200: downsample_mode must be 'stride' (0) or 'avg' (1)");
  const bool avg = d.downsample_mode == 1;
  // parameters of the skip branch (conv w, b, BN gamma, beta): absent where num_channels_skip[l] = 0
  // (inpainting.ipynb c14:11-16 "vase": models/skip.py:50-53 then adds `deeper` alone, no Concat)
  auto psl = [&](int l) { return NS[l] > 0 ? 4 : 0; };
  // wide: skip branch on the tensor cores, 256-channel concat (all widths 128, all skips 128)
  if (d.in_channels < 1 || d.in_channels > 128) return fail("dip-b200: input depth must be in [1,128]");
  // level-0 activations are stored with the input depth rounded up to a power of two >= 4 (zero channels); the conv
  // weights keep their real depth (TMA zero-fills the missing channels, the 1x1 skip conv reads its rows by element)
  int cin_eng = 4;
  while (cin_eng < d.in_channels) cin_eng *= 2;
  if (d.out_channels < 1 || d.out_channels > 4) return fail("dip-b200: num_output_channels must be <= 4");
  if (P->H % (1 << L) || P->W % (1 << L)) return fail("dip-b200: H and W must be divisible by 2^num_scales");
  if ((P->H >> L) < 2 || (P->W >> L) < 2)
    return fail("dip-b200: H and W must be at least 2 * 2^num_scales (ReflectionPad2d(1) in front of the deepest 3x3 conv needs 2 "
                "pixels per side: torch raises for the reference's network as well)");
  const int prec = d.precision;
  if (prec != DIP_PRECISION_TF32 && prec != DIP_PRECISION_FP32 && prec != DIP_PRECISION_BF16) return fail("dip-b200: unknown precision");
  const bool bf = prec == DIP_PRECISION_BF16;
  P->lv.resize(L);
  int pidx = 0;
  // parameter slots in net.parameters() order: assigned recursively (pre-order part, then post-order part)
  std::vector<int> pre(L), post(L);
  {
    int idx = 0;
    for (int l = 0; l < L; ++l) { pre[l] = idx; idx += 8 + psl(l); }
    for (int l = L - 1; l >= 0; --l) { post[l] = idx; idx += 10; }
    P->p_head_w = idx; P->p_head_b = idx + 1;
    pidx = idx + 2;
  }
  P->numel.assign(pidx, 0);
  P->bns.clear(); P->convs.clear();
  size_t acc_f = 0, acc_b = 0;
  auto bn_init = [&](BnLayer& b, int C, int rot, int n, int pg, int pbias) {
    b.C = C; b.rot = rot; b.n = (float)n; b.p_gamma = pg; b.p_beta = pg + 1; b.p_bias = pbias;
    P->numel[pg] = C; P->numel[pg + 1] = C;
    acc_f += 2 * C; acc_b += 3 * C;
  };
  for (int l = 0; l < L; ++l) {
    Level& v = P->lv[l];
    v.H = P->H >> l; v.W = P->W >> l; v.h = v.H / 2; v.w = v.W / 2;
    v.nd = ND[l]; v.nu = NU[l]; v.ns = NS[l]; v.cu = l == L - 1 ? ND[l] : NU[l + 1];
    const int CS = v.ns, ps = psl(l);   // (shadow the network-wide values inside the level loop)
    v.Cin = l == 0 ? cin_eng : ND[l - 1];
    v.Cin_act = l == 0 ? d.in_channels : ND[l - 1];
    v.bilinear = d.upsample_bilinear < 0 ? (d.upsample_mask >> l) & 1 : (d.upsample_bilinear != 0);
    const int b0 = pre[l], b1 = post[l];
    // skip conv 1x1 Cin -> CS
    if (CS > 0) {
      v.p_skip_w = b0; v.p_skip_b = b0 + 1;
      P->numel[b0] = (long long)CS * v.Cin_act; P->numel[b0 + 1] = CS;
      bn_init(v.bn_s, CS, 0, v.H * v.W, b0 + 2, b0 + 1);
    } else {
      v.p_skip_w = v.p_skip_b = -1;
      v.bn_s = BnLayer{};
    }
    // down1 3x3 s2
    v.d1.N = v.nd; v.d1.C = v.Cin_act; v.d1.k = 3; v.d1.stride = 2; v.d1.p_w = b0 + ps; v.d1.p_b = b0 + ps + 1;
    P->numel[b0 + ps] = (long long)v.nd * v.Cin_act * 9; P->numel[b0 + ps + 1] = v.nd;
    bn_init(v.bn_d1, v.nd, 0, v.h * v.w, b0 + ps + 2, b0 + ps + 1);
    // down2 3x3
    v.d2.N = v.nd; v.d2.C = v.nd; v.d2.k = 3; v.d2.stride = 1; v.d2.p_w = b0 + ps + 4; v.d2.p_b = b0 + ps + 5;
    P->numel[b0 + ps + 4] = (long long)v.nd * v.nd * 9; P->numel[b0 + ps + 5] = v.nd;
    bn_init(v.bn_d2, v.nd, 0, v.h * v.w, b0 + ps + 6, b0 + ps + 5);
    // concat BN (torch channel order [skip | up], engine order [up | skip])
    bn_init(v.bn_cat, v.cu + CS, CS, v.H * v.W, b1 + 0, -1);
    // up 3x3 (cu+CS) -> nu
    v.up.N = v.nu; v.up.C = v.cu + CS; v.up.k = 3; v.up.stride = 1; v.up.rot = CS; v.up.p_w = b1 + 2; v.up.p_b = b1 + 3;
    if (wide) {
      v.up.do_wgrad = false;
      v.sk.C = v.Cin_act; v.sk.k = 1; v.sk.stride = 1; v.sk.p_w = b0; v.sk.p_b = b0 + 1;
      for (ConvOp* h : {&v.up_a, &v.up_b}) {
        h->C = 128; h->Ctot = 128 + CS; h->k = 3; h->stride = 1; h->rot = CS; h->p_w = b1 + 2; h->p_b = b1 + 3;
        h->do_fprop = false; h->dg_ld = 128 + CS;
      }
      v.up_b.coff = 128;
    }
    P->numel[b1 + 2] = (long long)v.nu * (v.cu + CS) * 9; P->numel[b1 + 3] = v.nu;
    bn_init(v.bn_u, v.nu, 0, v.H * v.W, b1 + 4, b1 + 3);
    // 1x1 nu -> nu
    v.c11.N = v.nu; v.c11.C = v.nu; v.c11.k = 1; v.c11.stride = 1; v.c11.p_w = b1 + 6; v.c11.p_b = b1 + 7;
    P->numel[b1 + 6] = (long long)v.nu * v.nu; P->numel[b1 + 7] = v.nu;
    bn_init(v.bn_v, v.nu, 0, v.H * v.W, b1 + 8, b1 + 7);
  }
  const int NH = NU[0];   // depth of the tensor the RGB head reads
  P->numel[P->p_head_w] = (long long)d.out_channels * NH;
  P->numel[P->p_head_b] = d.out_channels;
  // BN order (running-stat table) follows state_dict order: skip, d1, d2, <deeper>, cat, up, 1x1
  {
    std::vector<BnLayer*> a, b;
    for (int l = 0; l < L; ++l) { if (NS[l] > 0) a.push_back(&P->lv[l].bn_s); a.push_back(&P->lv[l].bn_d1); a.push_back(&P->lv[l].bn_d2); }
    for (int l = L - 1; l >= 0; --l) { a.push_back(&P->lv[l].bn_cat); a.push_back(&P->lv[l].bn_u); a.push_back(&P->lv[l].bn_v); }
    P->bns = a;
    for (size_t i = 0; i < P->bns.size(); ++i) P->bns[i]->idx = (int)i;
  }
  // ---- accumulators
  size_t skinny_acc = 0;
  for (int l = 0; l < L; ++l) skinny_acc += (size_t)NS[l] * P->lv[l].Cin;
  skinny_acc += (size_t)d.out_channels * NH + 8;
  P->acc_fwd_n = acc_f * kAccS;
  P->acc_bwd_n = (acc_b + skinny_acc) * kAccS;
  P->acc_fwd = A.get<double>(P->acc_fwd_n);
  P->acc_bwd = A.get<double>(P->acc_bwd_n);
  {
    double* f = P->acc_fwd; double* b = P->acc_bwd;
    for (BnLayer* bn : P->bns) {
      bn->fwd = f; f = f ? f + 2 * bn->C * kAccS : nullptr;
      bn->bwd = b; bn->dbias = b ? b + 2 * bn->C * kAccS : nullptr; b = b ? b + 3 * bn->C * kAccS : nullptr;
    }
    for (int l = 0; l < L; ++l) { P->lv[l].dw_s = b; b = b ? b + (size_t)NS[l] * P->lv[l].Cin * kAccS : nullptr; }
    P->dw_head = b; b = b ? b + (size_t)d.out_channels * NH * kAccS : nullptr;
    P->db_head = b;
    P->db_scratch = b ? b + 4 * kAccS : nullptr;
  }
  // ---- activations
  auto reg = [&](const std::string& name, void* p, int rows, int cols, int ld, int c) { P->bufs[name] = BufInfo{p, rows, cols, ld, c}; };
  for (int l = 0; l < L; ++l) {
    Level& v = P->lv[l];
    const std::string pf = "L" + std::to_string(l) + ".";
    const size_t HW = (size_t)v.H * v.W, hw = (size_t)v.h * v.w;
    const size_t HWp = (size_t)(v.H + 2) * (v.W + 2), hwp = (size_t)(v.h + 2) * (v.w + 2);
    const bool last = l == L - 1;
    const int CS = v.ns, nd = v.nd, nu = v.nu, CC = v.cu + v.ns;
    if (l == 0) v.Pin = A.get<float>(HWp * v.Cin); else v.Pin = P->lv[l - 1].P_d2;
    v.raw_s = A.get<float>(HW * CS);
    v.raw_d1 = A.get<float>(hw * nd);
    v.P_d1 = A.get<float>(hwp * nd);
    v.raw_d2 = A.get<float>(hw * nd);
    v.P_d2 = A.get<float>(last ? hw * nd : hwp * nd);
    v.P_cat = A.get<float>(HWp * CC);
    v.raw_u = A.get<float>(HW * nu);
    v.A_u = A.get<float>(HW * nu);
    v.raw_v = A.get<float>(HW * nu);
    v.U = (l > 0 || nu != 128) ? A.get<float>(HW * nu) : nullptr;  // a 128-deep level 0 feeds the fused RGB head instead
    v.dRaw_v = A.get<float>(HW * nu);
    v.dA_u = A.get<float>(HW * nu);
    v.dRaw_u = A.get<float>(HW * nu);
    v.dP_cat = A.get<float>(HWp * CC);
    v.dCat = A.get<float>(HW * CC);
    v.dRaw_s = A.get<float>(HW * CS);
    v.dUp = A.get<float>(hw * v.cu);
    v.dRaw_d2 = A.get<float>(hw * nd);
    v.dP_d1 = A.get<float>(hwp * nd);
    v.dRaw_d1 = A.get<float>(hw * nd);
    const bool in_grad = d.input_grad != 0;   // level 0 then also needs its input gradient
    if (avg) {
      v.rawF = A.get<float>(HW * nd);
      v.dRawF = A.get<float>(HW * nd);
      if (bf) v.dRawF16 = A.get<uint16_t>(HW * nd);
    }
    v.ZS = (l > 0 || in_grad) ? A.get<float>(HW * nd) : nullptr;
    v.dS = ((wide && l > 0) || (in_grad && l == 0)) ? A.get<float>(HW * v.Cin) : nullptr;
    v.dPin = (l > 0 || in_grad) ? A.get<float>(HWp * v.Cin) : nullptr;
    if (bf) {
      v.Pin_ld16 = round_up(v.Cin, 8); v.cat_ld16 = round_up(CC, 8);
      if (l == 0) v.Pin16 = A.get<uint16_t>(HWp * v.Pin_ld16); else v.Pin16 = P->lv[l - 1].P_d2_16;
      v.P_d1_16 = A.get<uint16_t>(hwp * nd);
      v.P_d2_16 = last ? nullptr : A.get<uint16_t>(hwp * nd);
      v.P_cat16 = A.get<uint16_t>(HWp * v.cat_ld16);
      v.A_u16 = A.get<uint16_t>(HW * nu);
      v.dRaw_v16 = A.get<uint16_t>(HW * nu);
      v.dRaw_u16 = A.get<uint16_t>(HW * nu);
      v.dRaw_d2_16 = A.get<uint16_t>(hw * nd);
      v.dRaw_d1_16 = A.get<uint16_t>(hw * nd);
      v.dRaw_s16 = wide ? A.get<uint16_t>(HW * 128) : nullptr;
    }
    reg(pf + "Pin", v.Pin, v.H + 2, v.W + 2, v.Cin, v.Cin);
    reg(pf + "raw_s", v.raw_s, v.H, v.W, CS, CS);
    reg(pf + "raw_d1", v.raw_d1, v.h, v.w, nd, nd);
    reg(pf + "P_d1", v.P_d1, v.h + 2, v.w + 2, nd, nd);
    reg(pf + "raw_d2", v.raw_d2, v.h, v.w, nd, nd);
    if (last) reg(pf + "P_d2", v.P_d2, v.h, v.w, nd, nd); else reg(pf + "P_d2", v.P_d2, v.h + 2, v.w + 2, nd, nd);
    reg(pf + "P_cat", v.P_cat, v.H + 2, v.W + 2, CC, CC);
    reg(pf + "raw_u", v.raw_u, v.H, v.W, nu, nu);
    reg(pf + "A_u", v.A_u, v.H, v.W, nu, nu);
    reg(pf + "raw_v", v.raw_v, v.H, v.W, nu, nu);
    if (v.U != nullptr) reg(pf + "U", v.U, v.H, v.W, nu, nu);
    reg(pf + "dRaw_v", v.dRaw_v, v.H, v.W, nu, nu);
    reg(pf + "dA_u", v.dA_u, v.H, v.W, nu, nu);
    reg(pf + "dRaw_u", v.dRaw_u, v.H, v.W, nu, nu);
    reg(pf + "dP_cat", v.dP_cat, v.H + 2, v.W + 2, CC, CC);
    reg(pf + "dCat", v.dCat, v.H, v.W, CC, CC);
    reg(pf + "dRaw_s", v.dRaw_s, v.H, v.W, CS, CS);
    reg(pf + "dRaw_d2", v.dRaw_d2, v.h, v.w, nd, nd);
    reg(pf + "dP_d1", v.dP_d1, v.h + 2, v.w + 2, nd, nd);
    reg(pf + "dRaw_d1", v.dRaw_d1, v.h, v.w, nd, nd);
    if (bf) {   // bf16 twins of the conv operands (names end in "16": the host side views them as bf16)
      reg(pf + "Pin16", v.Pin16, v.H + 2, v.W + 2, v.Pin_ld16, v.Cin);
      reg(pf + "P_d1_16", v.P_d1_16, v.h + 2, v.w + 2, nd, nd);
      if (!last) reg(pf + "P_d2_16", v.P_d2_16, v.h + 2, v.w + 2, nd, nd);
      reg(pf + "P_cat16", v.P_cat16, v.H + 2, v.W + 2, v.cat_ld16, CC);
      reg(pf + "A_u16", v.A_u16, v.H, v.W, nu, nu);
      reg(pf + "dRaw_v16", v.dRaw_v16, v.H, v.W, nu, nu);
      reg(pf + "dRaw_u16", v.dRaw_u16, v.H, v.W, nu, nu);
      reg(pf + "dRaw_d2_16", v.dRaw_d2_16, v.h, v.w, nd, nd);
      reg(pf + "dRaw_d1_16", v.dRaw_d1_16, v.h, v.w, nd, nd);
      if (wide) reg(pf + "dRaw_s16", v.dRaw_s16, v.H, v.W, 128, 128);
    }
    if (l > 0) {
      reg(pf + "ZS", v.ZS, v.H, v.W, nd, nd);
      reg(pf + "dPin", v.dPin, v.H + 2, v.W + 2, v.Cin, v.Cin);
    }
  }
  P->out_saved = A.get<float>((size_t)P->H * P->W * d.out_channels);
  P->zbuf = A.get<float>((size_t)P->H * P->W * d.in_channels);
  reg("zbuf", P->zbuf, d.in_channels, P->H, P->W, P->W);   // torch-layout planes [C][H][W]: the runner's perturbed input
  P->dout = A.get<float>((size_t)P->H * P->W * d.out_channels);
  P->dl4 = A.get<float>((size_t)P->H * P->W * 4);
  P->ds_kern = A.get<float>(dip_plan::kDownMaxK * dip_plan::kDownMaxK);
  P->ds_y = A.get<float>((size_t)P->H * P->W * d.out_channels);
  P->ds_dy = A.get<float>((size_t)P->H * P->W * d.out_channels);
  P->loss_ring = A.get<double>(dip_plan::kLossRing);
  P->it_dev = A.get<int>(4);
  // ---- conv ops
  size_t partial_max = 0;
  for (int l = 0; l < L; ++l) {
    Level& v = P->lv[l];
    // down1: Pin (padded, stride 2) -> raw_d1
    ConvOp& a = v.d1;
    a.set_shapes();
    a.in = v.Pin; a.in_rows = v.H + 2; a.in_cols = v.W + 2; a.in_ld = v.Cin; a.offx = a.offy = 0;
    a.out = v.raw_d1; a.out_h = v.h; a.out_w = v.w; a.stats = v.bn_d1.fwd;
    a.has_dgrad = l > 0 || d.input_grad != 0;
    a.dg_ld = v.Cin;   // level 0: stored depth (>= the conv's real input depth)
    a.dg_s2 = bf || (prec == DIP_PRECISION_TF32 && getenv("DIP_ZERO_STUFF") == nullptr);   // A/B switch: the old zero-stuffed stride-1 dgrad
    if (a.dg_s2) { a.dg_in = v.dRaw_d1; a.dg_in_h = v.h; a.dg_in_w = v.w; }
    else { a.dg_in = v.ZS; a.dg_in_h = v.H; a.dg_in_w = v.W; }
    a.dg_out = v.dPin; a.dg_out_h = v.H + 2; a.dg_out_w = v.W + 2; a.dg_off = -2;
    a.wg_dy = v.dRaw_d1; a.wg_h = v.h; a.wg_w = v.w;
    a.in16 = v.Pin16; a.in_ld16 = v.Pin_ld16; a.dg_in16 = v.dRaw_d1_16; a.wg_dy16 = v.dRaw_d1_16;
    if (avg) {   // stride-1 conv at the level's full resolution; pooling is a separate pass (fwd_level / bwd_level)
      a.stride = 1; a.out = v.rawF; a.out_h = v.H; a.out_w = v.W; a.stats = nullptr;
      a.dg_s2 = false; a.dg_in = v.dRawF; a.dg_in_h = v.H; a.dg_in_w = v.W;
      a.wg_dy = v.dRawF; a.wg_h = v.H; a.wg_w = v.W;
      a.dg_in16 = v.dRawF16; a.wg_dy16 = v.dRawF16;
    }
    // down2: P_d1 -> raw_d2
    ConvOp& b = v.d2;
    b.set_shapes();
    b.in = v.P_d1; b.in_rows = v.h + 2; b.in_cols = v.w + 2; b.in_ld = v.nd; b.offx = b.offy = 0;
    b.out = v.raw_d2; b.out_h = v.h; b.out_w = v.w; b.stats = v.bn_d2.fwd;
    b.has_dgrad = true;
    b.dg_in = v.dRaw_d2; b.dg_in_h = v.h; b.dg_in_w = v.w; b.dg_out = v.dP_d1; b.dg_out_h = v.h + 2; b.dg_out_w = v.w + 2; b.dg_off = -2;
    b.wg_dy = v.dRaw_d2; b.wg_h = v.h; b.wg_w = v.w;
    b.in16 = v.P_d1_16; b.in_ld16 = v.nd; b.dg_in16 = v.dRaw_d2_16; b.wg_dy16 = v.dRaw_d2_16;
    // up: P_cat -> raw_u
    ConvOp& c = v.up;
    c.set_shapes();
    c.in = v.P_cat; c.in_rows = v.H + 2; c.in_cols = v.W + 2; c.in_ld = v.cu + v.ns; c.offx = c.offy = 0;
    c.out = v.raw_u; c.out_h = v.H; c.out_w = v.W; c.stats = v.bn_u.fwd;
    c.has_dgrad = !wide;
    c.dg_in = v.dRaw_u; c.dg_in_h = v.H; c.dg_in_w = v.W; c.dg_out = v.dP_cat; c.dg_out_h = v.H + 2; c.dg_out_w = v.W + 2; c.dg_off = -2;
    c.wg_dy = v.dRaw_u; c.wg_h = v.H; c.wg_w = v.W;
    c.in16 = v.P_cat16; c.in_ld16 = v.cat_ld16; c.dg_in16 = v.dRaw_u16; c.wg_dy16 = v.dRaw_u16;
    // 1x1: A_u -> raw_v
    ConvOp& e = v.c11;
    e.set_shapes();
    e.in = v.A_u; e.in_rows = v.H; e.in_cols = v.W; e.in_ld = v.nu; e.offx = e.offy = 0;
    e.out = v.raw_v; e.out_h = v.H; e.out_w = v.W; e.stats = v.bn_v.fwd;
    e.has_dgrad = true;
    e.dg_in = v.dRaw_v; e.dg_in_h = v.H; e.dg_in_w = v.W; e.dg_out = v.dA_u; e.dg_out_h = v.H; e.dg_out_w = v.W; e.dg_off = 0;
    e.wg_dy = v.dRaw_v; e.wg_h = v.H; e.wg_w = v.W;
    e.in16 = v.A_u16; e.in_ld16 = v.nu; e.dg_in16 = v.dRaw_v16; e.wg_dy16 = v.dRaw_v16;
    std::vector<ConvOp*> ops = {&a, &b, &c, &e};
    if (wide) {
      // skip conv 1x1 on the interior of the padded level input
      ConvOp& k1 = v.sk;
      k1.set_shapes();
      k1.in = v.Pin; k1.in_rows = v.H + 2; k1.in_cols = v.W + 2; k1.in_ld = v.Cin; k1.offx = k1.offy = 1;
      k1.out = v.raw_s; k1.out_h = v.H; k1.out_w = v.W; k1.stats = v.bn_s.fwd;
      k1.has_dgrad = l > 0 || d.input_grad != 0;
      k1.dg_ld = v.Cin;
      k1.dg_in = v.dRaw_s; k1.dg_in_h = v.H; k1.dg_in_w = v.W; k1.dg_out = v.dS; k1.dg_out_h = v.H; k1.dg_out_w = v.W; k1.dg_off = 0;
      k1.wg_dy = v.dRaw_s; k1.wg_h = v.H; k1.wg_w = v.W;
      k1.in16 = v.Pin16; k1.in_ld16 = v.Pin_ld16; k1.dg_in16 = v.dRaw_s16; k1.wg_dy16 = v.dRaw_s16;
      ops.push_back(&k1);
      for (ConvOp* h : {&v.up_a, &v.up_b}) {
        h->set_shapes();
        h->in = v.P_cat + h->coff; h->in_rows = v.H + 2; h->in_cols = v.W + 2; h->in_ld = 128 + CS; h->offx = h->offy = 0;
        h->out = v.raw_u; h->out_h = v.H; h->out_w = v.W; h->stats = nullptr;
        h->has_dgrad = true;
        h->dg_in = v.dRaw_u; h->dg_in_h = v.H; h->dg_in_w = v.W;
        h->dg_out = v.dP_cat + h->coff; h->dg_out_h = v.H + 2; h->dg_out_w = v.W + 2; h->dg_off = -2;
        h->wg_dy = v.dRaw_u; h->wg_h = v.H; h->wg_w = v.W;
        h->in16 = v.P_cat16 != nullptr ? v.P_cat16 + h->coff : nullptr; h->in_ld16 = v.cat_ld16; h->dg_in16 = v.dRaw_u16; h->wg_dy16 = v.dRaw_u16;
        ops.push_back(h);
      }
    }
    for (ConvOp* op : ops) {
      op->wp_f = op->do_fprop ? A.get<float>(op->wp_f_elems()) : nullptr;
      op->wp_d = op->has_dgrad ? A.get<float>(op->wp_d_elems()) : nullptr;
      op->simt_ksplits = op->wg_h < 64 ? op->wg_h : 64;
      op->bf16 = bf;
      op->timer = &P->timer;
      const size_t pe = op->do_wgrad ? op->partial_elems(prec) : 0;
      if (pe > partial_max) partial_max = pe;
      P->convs.push_back(op);
    }
  }
  P->partial = A.get<float>(partial_max);
  // weight-gradient accumulators of the tensor-core path: one per conv, contiguous (a single memset per backward)
  P->n_unpack = 0;
  if (is_tc(prec)) {
    size_t tot = 0;
    for (ConvOp* op : P->convs) if (op->do_wgrad) { tot += (op->wacc_elems() + 63) & ~size_t(63); P->n_unpack++; }
    P->wacc_base = A.get<float>(tot);
    P->wacc_bytes = tot * sizeof(float);
    size_t off = 0;
    for (ConvOp* op : P->convs) if (op->do_wgrad) { op->wacc = P->wacc_base ? P->wacc_base + off : nullptr; off += (op->wacc_elems() + 63) & ~size_t(63); }
  }
  P->d_unpack = A.get<UnpackEntry>(P->n_unpack > 0 ? P->n_unpack : 1);
  P->d_deep_fwd = A.get<DeepOp>(dip_plan::kDeepMaxOps);
  P->d_deep_bwd = A.get<DeepOp>(2 * dip_plan::kDeepMaxOps);
  P->deep_bar = A.get<unsigned>(64);
  if (const char* e = getenv("DIP_DEEP_FROM")) P->deep_from = atoi(e);
  if (P->deep_from < 1) P->deep_from = 1;
  P->n_pack = (int)P->convs.size();
  P->n_cvt = (int)P->bns.size() * 3 + L + 2;
  P->n_run = (int)P->bns.size();
  P->d_pack = A.get<PackEntry>(P->n_pack);
  P->d_cvt = A.get<CvtEntry>(P->n_cvt);
  P->d_run = A.get<RunEntry>(P->n_run);
  if (P->dry) return 0;
  // ---- device-side setup
  {
    // side streams at the lowest priority, the graph stream (= main chain of the captured step) at the highest
    int lo = 0, hi = 0;
    DIP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (getenv("DIP_NO_PRIO") != nullptr) lo = hi = 0;
    DIP_CUDA(cudaStreamCreateWithPriority(&P->wstream, cudaStreamNonBlocking, lo));
    DIP_CUDA(cudaStreamCreateWithPriority(&P->sstream, cudaStreamNonBlocking, lo));
  }
  // The zero-stuffed buffers are written at even positions only: clear them once.
  for (int l = 0; l < L; ++l)
    if (P->lv[l].ZS != nullptr) DIP_CUDA(cudaMemset(P->lv[l].ZS, 0, (size_t)P->lv[l].H * P->lv[l].W * P->lv[l].nd * sizeof(float)));
  if (is_tc(prec))
    for (ConvOp* op : P->convs) DIP_CHECK(op->build_tc(P->partial));
  P->pack_max = 0;
  for (ConvOp* op : P->convs) {
    const long long n = (op->do_fprop ? (long long)op->wp_f_elems() : 0) + (op->has_dgrad ? (long long)op->wp_d_elems() : 0);
    if (n > P->pack_max) P->pack_max = n;
  }
  return 0;
}

static int upload_tables(dip_plan* P) {
  std::vector<PackEntry> pk;
  for (ConvOp* op : P->convs) {
    PackEntry e{};
    e.w = P->params[op->p_w]; e.dst_f = op->do_fprop ? op->wp_f : nullptr; e.dst_d = op->has_dgrad ? op->wp_d : nullptr;
    e.N = op->N; e.C = op->C; e.k = op->k; e.rot = op->rot; e.n_rows = op->Np; e.c_pad = op->c_pad; e.c_rows = op->crows;
    e.Ctot = op->Ctot; e.coff = op->coff; e.s2 = op->dg_s2 ? 1 : 0;
    e.bf16 = op->bf16 ? 1 : 0; e.c_pad16 = op->c_pad16; e.n_pad = op->bf16 ? op->n_pad16 : op->n_pad;
    pk.push_back(e);
  }
  DIP_CUDA(cudaMemcpy(P->d_pack, pk.data(), pk.size() * sizeof(PackEntry), cudaMemcpyHostToDevice));
  if (P->n_unpack > 0) {
    std::vector<UnpackEntry> up;
    for (ConvOp* op : P->convs) {
      if (!op->do_wgrad || op->wacc == nullptr) continue;
      up.push_back(UnpackEntry{op->wacc, P->grads[op->p_w], op->N, op->C, op->k * op->k, op->rot, op->c_pad, op->Ctot, op->coff});
    }
    if ((int)up.size() != P->n_unpack) return fail("internal: unpack table size mismatch");
    DIP_CUDA(cudaMemcpy(P->d_unpack, up.data(), up.size() * sizeof(UnpackEntry), cudaMemcpyHostToDevice));
  }
  std::vector<CvtEntry> cv;
  for (BnLayer* b : P->bns) {
    cv.push_back(CvtEntry{b->bwd + b->C * kAccS, P->grads[b->p_gamma], b->C, b->rot});  // dgamma = sum dz*xhat
    cv.push_back(CvtEntry{b->bwd, P->grads[b->p_beta], b->C, b->rot});          // dbeta  = sum dz
    if (b->p_bias >= 0) cv.push_back(CvtEntry{b->dbias, P->grads[b->p_bias], b->C, 0});
    else cv.push_back(CvtEntry{b->dbias, nullptr, 0, 0});
  }
  for (size_t l = 0; l < P->lv.size(); ++l) {
    // skip=128: the skip conv's weight gradient comes from the tensor-core wgrad, not from fp64 accumulators
    if (P->lv[l].ns == 128 || P->lv[l].ns == 0) cv.push_back(CvtEntry{P->lv[l].dw_s, nullptr, 0, 0});
    else cv.push_back(CvtEntry{P->lv[l].dw_s, P->grads[P->lv[l].p_skip_w], (int)P->numel[P->lv[l].p_skip_w], 0,
                               P->lv[l].Cin_act, P->lv[l].Cin});   // accumulator rows hold the stored depth
  }
  cv.push_back(CvtEntry{P->dw_head, P->grads[P->p_head_w], (int)P->numel[P->p_head_w], 0});
  cv.push_back(CvtEntry{P->db_head, P->grads[P->p_head_b], (int)P->numel[P->p_head_b], 0});
  if ((int)cv.size() != P->n_cvt) return fail("internal: cvt table size mismatch");
  DIP_CUDA(cudaMemcpy(P->d_cvt, cv.data(), cv.size() * sizeof(CvtEntry), cudaMemcpyHostToDevice));
  std::vector<RunEntry> rn;
  for (BnLayer* b : P->bns) {
    RunEntry e{};
    e.fwd = b->fwd; e.C = b->C; e.rot = b->rot; e.n = b->n;
    if (!P->running.empty()) {
      e.rm = (float*)P->running[3 * b->idx]; e.rv = (float*)P->running[3 * b->idx + 1]; e.nb = P->running[3 * b->idx + 2];
      e.nb_is_float = P->nbt_is_float;
    }
    rn.push_back(e);
  }
  DIP_CUDA(cudaMemcpy(P->d_run, rn.data(), rn.size() * sizeof(RunEntry), cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------------------ side streams
// Dependency edge between two streams (event record + wait; inside graph capture this becomes a graph edge).
static void stream_edge(dip_plan* P, cudaStream_t from, cudaStream_t to) {
  if (P->wev_used == P->wev.size()) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    P->wev.push_back(e);
  }
  cudaEvent_t e = P->wev[P->wev_used++];
  cudaEventRecord(e, from);
  cudaStreamWaitEvent(to, e, 0);
}
// Stream on which the weight-gradient work that depends on everything recorded so far on `s` may run concurrently.
static cudaStream_t fork_side(dip_plan* P, cudaStream_t s) {
  if (!P->side_on) return s;
  stream_edge(P, s, P->wstream);
  return P->wstream;
}
static void join_side(dip_plan* P, cudaStream_t s) {
  if (P->side_on) stream_edge(P, P->wstream, s);
}
// Same for the skip-branch stream.
static bool skip_stream_on(const dip_plan* P) {
  static const bool off = getenv("DIP_NO_SKIPSTREAM") != nullptr;   // experiment switch
  return P->side_on && !off;
}
static cudaStream_t fork_skip(dip_plan* P, cudaStream_t s) {
  if (!skip_stream_on(P)) return s;
  stream_edge(P, s, P->sstream);
  return P->sstream;
}
static void join_skip(dip_plan* P, cudaStream_t s) {
  if (skip_stream_on(P)) stream_edge(P, P->sstream, s);
}

// ------------------------------------------------------------------------------------------------ forward
static CatArgs cat_args(const dip_plan* P, const Level& v, const float* Usrc) {
  CatArgs a;
  a.U = Usrc; a.raw_s = v.raw_s;
  if (v.ns > 0) a.bn_s = bn_ref(P, v.bn_s);
  else a.bn_s = BnRef{nullptr, nullptr, nullptr, 0, 0, 0.f};   // num_channels_skip = 0: the "concat" is the upsampled tensor alone
  a.Cu = v.cu; a.Cs = v.ns; a.H = v.H; a.W = v.W; a.bilinear = v.bilinear;
  return a;
}
static const float* level_usrc(const dip_plan* P, int l) {
  const int L = (int)P->lv.size();
  return l == L - 1 ? P->lv[l].P_d2 : P->lv[l + 1].U;
}

// The persistent deep-level kernel (deep.cu) replaces the launches of levels >= deep_from (tensor-core precision, no
// per-launch timers).  OPT-IN (DIP_DEEP=1): parity-green (tests/test_engine_gpu.py::test_deep_kernel_matches_launches) but
// measured SLOWER than the launches it replaces on B200 -- forward 382 us vs 340 us, backward 745 us vs ~700 us of kernel time
// (ncu), 314 vs 378 it/s end to end: the small kernels are bounded by their own prologue + a few memory round trips, not by
// launch gaps, and one 256-thread CTA per SM (255 registers: the conv code lives in the same kernel) has an eighth of the
// loads in flight that the stand-alone kernels have.  See DESIGN.md section 10.
static bool deep_on(const dip_plan* P) {
  const char* e = getenv("DIP_DEEP");   // read per call: the test switches it inside one process
  return e != nullptr && e[0] == '1' && P->desc.precision == DIP_PRECISION_TF32 && !P->timer.on && P->n_deep_fwd > 0 &&
         (int)P->lv.size() > P->deep_from;
}
static int fwd_level(dip_plan* P, int l, cudaStream_t s, int& nl) {
  Level& v = P->lv[l];
  const int prec = P->desc.precision;
  const int CS = v.ns, nd = v.nd, nu = v.nu;
  const bool last = l == (int)P->lv.size() - 1;
  const float* pin_interior = v.Pin + ((size_t)(v.W + 2) + 1) * v.Cin;
  // precision mode bf16: conv inputs are written as bf16 twins; the fp32 tensor is dropped where only convolutions read it
  const bool bf = prec == DIP_PRECISION_BF16;
  // skip branch: 1x1 conv Cin -> CS (+ statistics); independent of the deeper branch until the concat -> skip stream
  if (CS > 0) {
    cudaStream_t ks = fork_skip(P, s);
    if (CS == 128) {
      DIP_CHECK(v.sk.run_fprop(prec, P->params[v.p_skip_b], ks));
      nl += prec == DIP_PRECISION_FP32 ? 2 : 1;
    } else {
      HBM_T(&P->timer, H_SKINNY_FWD, 0, (double)v.H * v.W * (v.Cin_act + CS) * sizeof(float), ks,
            launch_skinny_fwd(pin_interior, v.Cin, v.W + 2, P->params[v.p_skip_w], P->params[v.p_skip_b], v.Cin, CS, v.H, v.W, v.raw_s, 0,
                              v.bn_s.fwd, ks, v.Cin_act));
      nl += 1;
    }
  }
  // deeper branch
  DIP_CHECK(v.d1.run_fprop(prec, P->params[v.d1.p_b], s));
  if (v.rawF != nullptr) {   // downsample_mode 'avg': pool the stride-1 conv output, then the statistics of the pooled tensor
    launch_avgpool2(v.rawF, v.h, v.w, nd, v.raw_d1, s);
    launch_channel_stats(v.raw_d1, nd, nd, v.h * v.w, v.bn_d1.fwd, s);
    nl += 2;
  }
  HBM_T(&P->timer, H_BN_ACT_WRITE, 1, (double)nd * ((double)v.h * v.w + (double)(v.h + 2) * (v.w + 2)) * sizeof(float), s,
        launch_bn_act_write(v.raw_d1, nd, bn_ref(P, v.bn_d1), v.h, v.w, bf ? nullptr : v.P_d1, nd, 1, 1, s, Twin{v.P_d1_16, nd}));
  DIP_CHECK(v.d2.run_fprop(prec, P->params[v.d2.p_b], s));
  HBM_T(&P->timer, H_BN_ACT_WRITE, last ? 0 : 1,
        (double)nd * ((double)v.h * v.w + (last ? (double)v.h * v.w : (double)(v.h + 2) * (v.w + 2))) * sizeof(float), s,
        launch_bn_act_write(v.raw_d2, nd, bn_ref(P, v.bn_d2), v.h, v.w, (bf && !last && P->lv[l + 1].ns != 4) ? nullptr : v.P_d2, nd, last ? 0 : 1, 1, s,
                            Twin{last ? nullptr : v.P_d2_16, nd}));   // (the 4-channel skip conv of the next level reads the fp32 tensor)
  nl += 4 + (prec == DIP_PRECISION_FP32 ? 2 : 0);
  if (!last) {
    if (deep_on(P) && l + 1 == P->deep_from) {
      // every level below this one: ONE persistent launch (deep.cu) instead of ~13 launches per level
      DIP_CUDA(launch_deep(P->d_deep_fwd, P->n_deep_fwd, P->deep_bar, P->deep_grid, s));
      nl += 2;
    } else {
      DIP_CHECK(fwd_level(P, l + 1, s, nl));
    }
  }
  // upsample + concat + BN + pad
  if (CS > 0) join_skip(P, s);
  CatArgs ca = cat_args(P, v, level_usrc(P, l));
  const double cat_in = ((double)v.cu * v.h * v.w + (double)CS * v.H * v.W) * sizeof(float);
  HBM_T(&P->timer, H_CAT_STATS, v.bilinear, cat_in, s, launch_cat_stats(ca, v.bn_cat.fwd, s));
  HBM_T(&P->timer, H_CAT_WRITE, v.bilinear, cat_in + ((double)v.cu + CS) * (v.H + 2) * (v.W + 2) * sizeof(float), s,
        launch_cat_write(ca, bn_ref(P, v.bn_cat), v.P_cat, s, Twin{v.P_cat16, v.cat_ld16}));
  DIP_CHECK(v.up.run_fprop(prec, P->params[v.up.p_b], s));
  HBM_T(&P->timer, H_BN_ACT_WRITE, 0, 2.0 * nu * v.H * v.W * sizeof(float), s,
        launch_bn_act_write(v.raw_u, nu, bn_ref(P, v.bn_u), v.H, v.W, bf ? nullptr : v.A_u, nu, 0, 1, s, Twin{v.A_u16, nu}));
  DIP_CHECK(v.c11.run_fprop(prec, P->params[v.c11.p_b], s));
  if (l > 0 || nu != 128) {
    HBM_T(&P->timer, H_BN_ACT_WRITE, 0, 2.0 * nu * v.H * v.W * sizeof(float), s,
          launch_bn_act_write(v.raw_v, nu, bn_ref(P, v.bn_v), v.H, v.W, v.U, nu, 0, 1, s));
    if (l == 0) {
      // level 0 narrower than 128 channels: the RGB head (models/skip.py:95-98) as a skinny 1x1 conv over the materialised
      // activation (the fused BN + head kernel is specialised for 128 channels = one warp per pixel)
      HBM_T(&P->timer, H_SKINNY_FWD, 1, ((double)nu + P->desc.out_channels) * v.H * v.W * sizeof(float), s,
            launch_skinny_fwd(v.U, nu, v.W, P->params[P->p_head_w], P->params[P->p_head_b], nu, P->desc.out_channels, v.H, v.W,
                              P->out_saved, P->desc.need_sigmoid != 0 ? 1 : 2, nullptr, s));
      nl += 1;
    }
  } else {
    // top level: BN + LeakyReLU + RGB head + sigmoid in one pass; the 128-channel activation is never materialised
    HeadRef hd{P->params[P->p_head_w], P->params[P->p_head_b], P->desc.out_channels, P->out_saved, P->desc.need_sigmoid != 0};
    HBM_T(&P->timer, H_BN_ACT_HEAD, 0, (128.0 + P->desc.out_channels) * v.H * v.W * sizeof(float), s,
          launch_bn_act_head(v.raw_v, bn_ref(P, v.bn_v), v.H, v.W, hd, s));
  }
  nl += 6 + (prec == DIP_PRECISION_FP32 ? 2 : 0);
  DIP_CUDA(cudaGetLastError());
  return 0;
}

// one table-driven launch repacks the weights of all wide convs (OIHW -> per-tap K-major fprop / dgrad operands)
static void plan_pack(dip_plan* P, cudaStream_t s) {
  dim3 grid((unsigned)((P->pack_max + 255) / 256 < 64 ? (P->pack_max + 255) / 256 : 64), P->n_pack);
  double bytes = 0;
  for (ConvOp* op : P->convs)
    bytes += ((double)op->N * op->C * op->k * op->k + (op->do_fprop ? (double)op->wp_f_elems() : 0.0) + (op->has_dgrad ? (double)op->wp_d_elems() : 0.0)) * sizeof(float);
  HBM_T(&P->timer, H_PACK, 0, bytes, s, launch_k(k_pack_table, dim3(grid), dim3(256), 0, s, 1, P->d_pack));
}

static int plan_forward(dip_plan* P, const float* z, const float* noise, float sigma, float* out, cudaStream_t s) {
  if (!P->bound) return fail("dip_forward: parameters not bound (call dip_plan_bind)");
  // grid-wide reductions: fp64 atomics onto line-strided accumulators (default) or the deterministic last-block sum
  int nl = 0;
  P->side_on = getenv("DIP_NO_SIDE") == nullptr;
  if (!P->prepacked) P->wev_used = 0;
  DIP_CUDA(cudaMemsetAsync(P->acc_fwd, 0, P->acc_fwd_n * sizeof(double), s));
  if (!P->prepacked) plan_pack(P, fork_side(P, s));   // weight repack runs beside the input transform
  P->prepacked = false;
  Level& v0 = P->lv[0];
  if (P->fnoise.on)
    HBM_T(&P->timer, H_NOISE, 1, ((double)v0.Cin_act * v0.H * v0.W + (double)v0.Cin * (v0.H + 2) * (v0.W + 2)) * sizeof(float), s,
          launch_noise_pad(P->fnoise.z0, P->fnoise.sigma, P->fnoise.seed, P->fnoise.offset, P->fnoise.it_dev, v0.Pin, v0.Cin,
                           v0.H, v0.W, v0.Cin_act, s, Twin{v0.Pin16, v0.Pin_ld16}));
  else
    HBM_T(&P->timer, H_INPUT_PAD, noise != nullptr,
          ((noise != nullptr ? 2.0 : 1.0) * v0.Cin_act * v0.H * v0.W + (double)v0.Cin * (v0.H + 2) * (v0.W + 2)) * sizeof(float), s,
          launch_input_pad(z, noise, sigma, v0.Pin, v0.Cin, v0.H, v0.W, s, v0.Cin_act, Twin{v0.Pin16, v0.Pin_ld16}));
  join_side(P, s);
  nl += 3;
  DIP_CHECK(fwd_level(P, 0, s, nl));
  if (out != nullptr && out != P->out_saved)
    DIP_CUDA(cudaMemcpyAsync(out, P->out_saved, (size_t)v0.H * v0.W * P->desc.out_channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
  launch_k(k_running_table, dim3(P->n_run), dim3(160), 0, s, 1, P->d_run);
  nl += 3;
  DIP_CUDA(cudaGetLastError());
  P->launches_fwd = nl;
  return 0;
}

// ------------------------------------------------------------------------------------------------ backward
static int bn_bwd(dip_plan* P, const float* raw, int ld_raw, BnLayer& b, int act, GradSrc src, int H, int W, float* draw,
                  float* zs, cudaStream_t s, int& nl, uint16_t* draw16 = nullptr) {
  if (draw16 != nullptr && zs == nullptr) draw = nullptr;   // bf16 mode: only the tensor-core dgrad / wgrad read this gradient
  BnRef r = bn_ref(P, b);
  // algorithmic bytes: raw + the gradient source as the kernel's contract names it (plain [H][W][C]; fold: the padded
  // dgrad output (+ the 4-channel skip-branch gradient / the plain addend); upsample adjoint: the 2H x 2W gradient; head:
  // the 4 logit gradients per pixel), + the written input gradient (and its zero-stuffed copy) for the apply pass
  const double px = (double)H * W, C4 = b.C * sizeof(float);
  double gsrc = px * C4;
  if (src.kind == 1) gsrc = (double)(H + 2) * (W + 2) * C4 + (src.ds != nullptr ? px * 4 * sizeof(float) : 0.0) + (src.add != nullptr ? px * C4 : 0.0);
  else if (src.kind == 2) gsrc = 4.0 * px * C4;
  else if (src.kind == 3) gsrc = px * 4 * sizeof(float);
  HBM_T(&P->timer, H_BN_BWD_REDUCE, src.kind, px * C4 + gsrc, s, launch_bn_bwd_reduce(raw, ld_raw, r, act, src, H, W, b.bwd, s));
  HBM_T(&P->timer, H_BN_BWD_APPLY, src.kind + (zs != nullptr ? 4 : 0), px * C4 + gsrc + px * C4 * (zs != nullptr ? 2.0 : 1.0), s,
        launch_bn_bwd_apply(raw, ld_raw, r, act, src, H, W, b.bwd, draw, zs, b.dbias, s, Twin{draw16, b.C}));
  nl += 2;
  return 0;
}
static GradSrc src_plain(const float* g, int ld, int coff) { GradSrc s{}; s.kind = 0; s.g = g; s.ld = ld; s.coff = coff; return s; }
// fold of a padded dgrad output; optionally + the dgrad of the next level's 1x1 skip conv computed on the fly
static GradSrc src_fold(const float* gp, int ld, const float* ds, const float* w2, int n2) {
  GradSrc s{}; s.kind = 1; s.g = gp; s.ld = ld; s.coff = 0; s.ds = ds; s.w2 = w2; s.n2 = n2; return s;
}
static GradSrc src_upadj(const float* d, int ld, int bilinear) { GradSrc s{}; s.kind = 2; s.g = d; s.ld = ld; s.coff = 0; s.bilinear = bilinear; return s; }

// Weight gradient (side stream) and input gradient (main stream, on the critical path) of one conv.  Both are persistent
// kernels that fill every SM, so they run one after the other whichever way: the dgrad is enqueued first and the main
// stream has the higher priority, so that the wgrad overlaps the HBM-bound kernels that follow the dgrad instead of
// delaying it.  (DIP_WGRAD_FIRST=1: the old order, for A/B runs.)
static int defer_level() {
  // measured (profiles/r01_matrix_wgrad_schedule.txt): 347.5 it/s without deferral, 352.3 with the level-0/1 wgrads deferred
  static const int lv = getenv("DIP_DEFER_WGRAD") ? atoi(getenv("DIP_DEFER_WGRAD")) : 2;   // 0: no deferral
  return lv;
}
static int flush_deferred(dip_plan* P, int prec, cudaStream_t s) {
  if (P->deferred.empty()) return 0;
  cudaStream_t ws = fork_side(P, s);
  for (ConvOp* op : P->deferred) DIP_CHECK(op->run_wgrad(prec, P->partial, P->grads[op->p_w], ws));
  P->deferred.clear();
  return 0;
}
static int conv_backward(dip_plan* P, ConvOp& op, bool dgrad, int prec, cudaStream_t s, int level = 99) {
  static const bool wgrad_first = getenv("DIP_WGRAD_FIRST") != nullptr;
  if (P->side_on && level < defer_level() && (int)P->lv.size() > defer_level()) {
    if (dgrad) DIP_CHECK(op.run_dgrad(prec, s));
    P->deferred.push_back(&op);
    return 0;
  }
  static const bool after_dgrad = getenv("DIP_WGRAD_AFTER_DGRAD") != nullptr;
  if (after_dgrad) {
    // the wgrad becomes runnable only once its sibling dgrad has finished: it then starts beside the HBM-bound kernels
    // that follow the dgrad instead of fighting it for the SMs
    if (dgrad) DIP_CHECK(op.run_dgrad(prec, s));
    return op.run_wgrad(prec, P->partial, P->grads[op.p_w], fork_side(P, s));
  }
  cudaStream_t ws = fork_side(P, s);
  if (wgrad_first) DIP_CHECK(op.run_wgrad(prec, P->partial, P->grads[op.p_w], ws));
  if (dgrad) DIP_CHECK(op.run_dgrad(prec, s));
  if (!wgrad_first) DIP_CHECK(op.run_wgrad(prec, P->partial, P->grads[op.p_w], ws));
  return 0;
}

static int bwd_level(dip_plan* P, int l, GradSrc src_v, cudaStream_t s, int& nl) {
  Level& v = P->lv[l];
  const int prec = P->desc.precision;
  const int CS = v.ns, nd = v.nd, nu = v.nu;
  const int CC = v.cu + CS;
  const bool last = l == (int)P->lv.size() - 1;
  const int wl = is_tc(prec) ? 1 : 2;   // tensor-core wgrads accumulate in place (no per-conv reduction launch)
  // 1x1 conv + BN + LReLU
  DIP_CHECK(bn_bwd(P, v.raw_v, nu, v.bn_v, 1, src_v, v.H, v.W, v.dRaw_v, nullptr, s, nl, v.dRaw_v16));
  if (l == defer_level()) DIP_CHECK(flush_deferred(P, prec, s));
  DIP_CHECK(conv_backward(P, v.c11, true, prec, s, l));
  nl += wl + 1;
  // up conv + BN + LReLU
  DIP_CHECK(bn_bwd(P, v.raw_u, nu, v.bn_u, 1, src_plain(v.dA_u, nu, 0), v.H, v.W, v.dRaw_u, nullptr, s, nl, v.dRaw_u16));
  if (CS == 128) {
    cudaStream_t ws = fork_side(P, s);
    DIP_CHECK(v.up_a.run_dgrad(prec, s));
    DIP_CHECK(v.up_b.run_dgrad(prec, s));
    DIP_CHECK(v.up_a.run_wgrad(prec, P->partial, P->grads[v.up.p_w], ws));
    DIP_CHECK(v.up_b.run_wgrad(prec, P->partial, P->grads[v.up.p_w], ws));
    nl += 2 * (wl + 1);
  } else {
    DIP_CHECK(conv_backward(P, v.up, true, prec, s, l));
    nl += wl + 1;
  }
  // concat BN
  BnRef rc = bn_ref(P, v.bn_cat);
  const double catb = ((double)v.H * v.W + (double)(v.H + 2) * (v.W + 2)) * CC * sizeof(float);   // stored BN output + padded gradient
  HBM_T(&P->timer, H_CAT_BWD_REDUCE, 0, catb, s, launch_cat_bwd_reduce(v.P_cat, rc, v.dP_cat, CC, v.H, v.W, v.bn_cat.bwd, s));
  HBM_T(&P->timer, H_CAT_BWD_APPLY, 0, catb + (double)v.H * v.W * CC * sizeof(float), s,
        launch_cat_bwd_apply(v.P_cat, rc, v.dP_cat, CC, v.H, v.W, v.bn_cat.bwd, v.dCat, s));
  // skip branch (on the skip stream: independent of the deeper levels; the level above joins before it reads dRaw_s / dS)
  cudaStream_t ks = fork_skip(P, s);
  // gradient w.r.t. the low-resolution tensor that was upsampled into this concat (adjoint of x2 upsampling), once
  HBM_T(&P->timer, H_UPADJ, v.bilinear, (double)v.cu * ((double)v.H * v.W + (double)v.h * v.w) * sizeof(float), s,
        launch_upadj(v.dCat, CC, 0, v.h, v.w, v.cu, v.bilinear, v.dUp, s));
  nl += 3;
  if (CS > 0) DIP_CHECK(bn_bwd(P, v.raw_s, CS, v.bn_s, 1, src_plain(v.dCat, CC, v.cu), v.H, v.W, v.dRaw_s, nullptr, ks, nl, v.dRaw_s16));
  if (CS == 0) {
    // no skip branch (models/skip.py:50-53 with num_channels_skip = 0)
  } else if (CS == 128) {
    DIP_CHECK(v.sk.run_wgrad(prec, P->partial, P->grads[v.p_skip_w], fork_side(P, ks)));
    nl += wl;
    if (l > 0 || P->desc.input_grad) { DIP_CHECK(v.sk.run_dgrad(prec, ks)); nl += 1; }   // dS, added to the fold of dPin by the level above
  } else {
    const float* pin_interior = v.Pin + ((size_t)(v.W + 2) + 1) * v.Cin;
    // weight gradient only: the input gradient of this conv is folded into the BN backward of the level above
    HBM_T(&P->timer, H_SKINNY_BWD, 0, (double)v.H * v.W * (v.Cin_act + CS) * sizeof(float), ks,
          launch_skinny_bwd(pin_interior, v.Cin, v.W + 2, P->params[v.p_skip_w], v.Cin, CS, v.H, v.W, v.dRaw_s, nullptr, 0,
                            (l == 0 && P->desc.input_grad) ? v.dS : nullptr, v.dw_s, nullptr /*bias grad comes from the BN backward*/, ks, v.Cin_act));
    nl += 1;
  }
  // deeper branch
  GradSrc src_d2;
  if (!last) {
    if (deep_on(P) && P->n_deep_bwd > 0 && l + 1 == P->deep_from) {
      if (P->deep_from >= defer_level()) DIP_CHECK(flush_deferred(P, prec, s));   // outer-level wgrads run beside the deep kernel
      DIP_CUDA(launch_deep(P->d_deep_bwd, P->n_deep_bwd, P->deep_bar + 16, P->deep_grid, s));
      nl += 2;
    } else {
      DIP_CHECK(bwd_level(P, l + 1, src_plain(v.dUp, v.cu, 0), s, nl));
    }
    join_skip(P, s);   // the next level's skip-branch gradients (dRaw_s / dS) feed the BN backward below
    Level& n = P->lv[l + 1];   // (its input depth n.Cin == nd)
    if (n.ns == 128) { src_d2 = src_fold(n.dPin, nd, nullptr, nullptr, 0); src_d2.add = n.dS; src_d2.ld_add = nd; }
    else if (n.ns == 0) src_d2 = src_fold(n.dPin, nd, nullptr, nullptr, 0);
    else src_d2 = src_fold(n.dPin, nd, n.dRaw_s, P->params[n.p_skip_w], n.ns);
  } else {
    src_d2 = src_plain(v.dUp, v.cu, 0);
  }
  DIP_CHECK(bn_bwd(P, v.raw_d2, nd, v.bn_d2, 1, src_d2, v.h, v.w, v.dRaw_d2, nullptr, s, nl, v.dRaw_d2_16));
  DIP_CHECK(conv_backward(P, v.d2, true, prec, s));
  nl += wl + 1;
  const bool d1_dgrad = l > 0 || P->desc.input_grad != 0;
  const bool avg = v.rawF != nullptr;
  DIP_CHECK(bn_bwd(P, v.raw_d1, nd, v.bn_d1, 1, src_fold(v.dP_d1, nd, nullptr, nullptr, 0), v.h, v.w, v.dRaw_d1,
                   (d1_dgrad && !v.d1.dg_s2 && !avg) ? v.ZS : nullptr, s, nl, avg ? nullptr : v.dRaw_d1_16));
  if (avg) {   // adjoint of AvgPool2d(2, 2): the conv's dY at full resolution
    launch_avgpool2_bwd(v.dRaw_d1, v.h, v.w, nd, prec == DIP_PRECISION_BF16 ? nullptr : v.dRawF, s, Twin{v.dRawF16, nd});
    nl += 1;
  }
  DIP_CHECK(conv_backward(P, v.d1, d1_dgrad, prec, s));
  nl += wl + (d1_dgrad ? 1 : 0);
  DIP_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ deep-level op lists
static void vec_lanes(int C, int* VL, int* PPB) { *VL = C / 4; *PPB = 256 / *VL; if (*PPB < 1) *PPB = 1; }
static DeepOp deep_conv(dip_plan* P, const TcConvParams& src, const float* bias, double* stats, int sync = 1) {
  DeepOp o;
  o.type = DO_CONV; o.sync = sync;
  o.u.conv = src;
  o.u.conv.bias = bias;
  o.u.conv.stats = stats;
  fit_stages(o.u.conv, deep_dyn_smem());
  o.u.conv.vgrid = tc_conv_grid(o.u.conv, g_num_sms);
  if (o.u.conv.vgrid > P->deep_grid) P->deep_grid = o.u.conv.vgrid;
  return o;
}
static DeepOp deep_bn_act_write(dip_plan* P, const float* raw, const BnLayer& b, int H, int W, float* dst, int pad) {
  DeepOp o;
  o.type = DO_BN_ACT_WRITE;
  vec_lanes(b.C, &o.VL, &o.PPB);
  o.u.bnw = DeepBnActWrite{raw, b.C, bn_ref(P, b), H, W, dst, b.C, pad, 1};
  return o;
}
static void deep_fwd_level(dip_plan* P, int l, std::vector<DeepOp>& ops) {
  Level& v = P->lv[l];
  const int CS = P->desc.skip_channels;
  const bool last = l == (int)P->lv.size() - 1;
  const float* pin_interior = v.Pin + ((size_t)(v.W + 2) + 1) * v.Cin;
  // skip branch (independent of the deeper branch until the concat: no barrier behind it)
  if (CS == 128) {
    ops.push_back(deep_conv(P, v.sk.fp, P->params[v.p_skip_b], v.bn_s.fwd, 0));
  } else {
    DeepOp o;
    o.type = DO_SKINNY_FWD; o.sync = 0;
    o.u.skf = DeepSkinnyFwd{pin_interior, v.Cin, v.W + 2, P->params[v.p_skip_w], P->params[v.p_skip_b], v.Cin, CS, v.H, v.W, v.raw_s,
                            0, v.bn_s.fwd, v.Cin_act};
    ops.push_back(o);
  }
  ops.push_back(deep_conv(P, v.d1.fp, P->params[v.d1.p_b], v.bn_d1.fwd));
  ops.push_back(deep_bn_act_write(P, v.raw_d1, v.bn_d1, v.h, v.w, v.P_d1, 1));
  ops.push_back(deep_conv(P, v.d2.fp, P->params[v.d2.p_b], v.bn_d2.fwd));
  ops.push_back(deep_bn_act_write(P, v.raw_d2, v.bn_d2, v.h, v.w, v.P_d2, last ? 0 : 1));
  if (!last) deep_fwd_level(P, l + 1, ops);
  {
    DeepOp o;
    o.type = DO_CAT_STATS;
    vec_lanes(128 + CS, &o.VL, &o.PPB);
    o.u.cat = DeepCat{cat_args(P, v, level_usrc(P, l)), v.bn_cat.fwd, bn_ref(P, v.bn_cat), v.P_cat};
    ops.push_back(o);
    o.type = DO_CAT_WRITE;
    ops.push_back(o);
  }
  ops.push_back(deep_conv(P, v.up.fp, P->params[v.up.p_b], v.bn_u.fwd));
  {
    DeepOp o = deep_bn_act_write(P, v.raw_u, v.bn_u, v.H, v.W, v.A_u, 0);
    ops.push_back(o);
  }
  ops.push_back(deep_conv(P, v.c11.fp, P->params[v.c11.p_b], v.bn_v.fwd));
  ops.push_back(deep_bn_act_write(P, v.raw_v, v.bn_v, v.H, v.W, v.U, 0));
}
static void deep_bn_bwd(dip_plan* P, const float* raw, int ld_raw, BnLayer& b, GradSrc src, int H, int W, float* draw,
                        std::vector<DeepOp>& ops, int sync_last = 1) {
  DeepOp o;
  o.type = DO_BN_BWD_REDUCE;
  vec_lanes(b.C, &o.VL, &o.PPB);
  o.u.bnb = DeepBnBwd{raw, ld_raw, bn_ref(P, b), 1, src, H, W, b.bwd, draw, nullptr, b.dbias};
  ops.push_back(o);
  o.type = DO_BN_BWD_APPLY; o.sync = sync_last;
  ops.push_back(o);
}
static DeepOp deep_wgrad(dip_plan* P, const ConvOp& op, int sync = 1) {
  DeepOp o;
  o.type = DO_WGRAD; o.sync = sync;
  o.u.wg = op.wg;
  o.u.wg.atomic = 1;
  o.u.wg.partial = op.wacc;
  const int cap = P->deep_grid / op.k;   // kh * ksplits CTAs must fit the deep grid
  if (o.u.wg.ksplits > cap) o.u.wg.ksplits = cap;
  while (tc_wgrad_smem_bytes(o.u.wg) > deep_dyn_smem() && o.u.wg.stages > 1) o.u.wg.stages--;
  return o;
}
static void deep_bwd_level(dip_plan* P, int l, GradSrc src_v, std::vector<DeepOp>& ops) {
  Level& v = P->lv[l];
  const int CS = P->desc.skip_channels, CC = 128 + CS;
  const bool last = l == (int)P->lv.size() - 1;
  // 1x1 conv + BN + LReLU
  deep_bn_bwd(P, v.raw_v, 128, v.bn_v, src_v, v.H, v.W, v.dRaw_v, ops);
  ops.push_back(deep_conv(P, v.c11.dg, nullptr, nullptr, 0));
  ops.push_back(deep_wgrad(P, v.c11));
  // up conv + BN + LReLU
  deep_bn_bwd(P, v.raw_u, 128, v.bn_u, src_plain(v.dA_u, 128, 0), v.H, v.W, v.dRaw_u, ops);
  if (CS == 128) {
    ops.push_back(deep_conv(P, v.up_a.dg, nullptr, nullptr, 0));
    ops.push_back(deep_conv(P, v.up_b.dg, nullptr, nullptr, 0));
    ops.push_back(deep_wgrad(P, v.up_a, 0));
    ops.push_back(deep_wgrad(P, v.up_b));
  } else {
    ops.push_back(deep_conv(P, v.up.dg, nullptr, nullptr, 0));
    ops.push_back(deep_wgrad(P, v.up));
  }
  // concat BN
  {
    DeepOp o;
    o.type = DO_CAT_BWD_REDUCE;
    vec_lanes(CC, &o.VL, &o.PPB);
    o.u.catb = DeepCatBwd{v.P_cat, bn_ref(P, v.bn_cat), v.dP_cat, CC, v.H, v.W, v.bn_cat.bwd, v.dCat};
    ops.push_back(o);
    o.type = DO_CAT_BWD_APPLY;
    ops.push_back(o);
  }
  {
    DeepOp o;   // adjoint of the x2 upsampling; the skip-branch ops behind it only read dCat as well: no barrier
    o.type = DO_UPADJ; o.sync = 0;
    vec_lanes(128, &o.VL, &o.PPB);
    o.u.up = DeepUpadj{v.dCat, CC, 0, v.h, v.w, 128, v.bilinear, v.dUp};
    ops.push_back(o);
  }
  // skip branch
  deep_bn_bwd(P, v.raw_s, CS, v.bn_s, src_plain(v.dCat, CC, 128), v.H, v.W, v.dRaw_s, ops);
  if (CS == 128) {
    ops.push_back(deep_conv(P, v.sk.dg, nullptr, nullptr, 0));   // dS (levels > 0 always need it)
    ops.push_back(deep_wgrad(P, v.sk));
  } else {
    const float* pin_interior = v.Pin + ((size_t)(v.W + 2) + 1) * v.Cin;
    DeepOp o;
    o.type = DO_SKINNY_BWD; o.sync = 1;
    vec_lanes(v.Cin, &o.VL, &o.PPB);
    o.u.skb = DeepSkinnyBwd{pin_interior, v.Cin, v.W + 2, P->params[v.p_skip_w], v.Cin, CS, v.H, v.W, v.dRaw_s, nullptr, 0, nullptr,
                            v.dw_s, nullptr, v.Cin_act};
    ops.push_back(o);
  }
  // deeper branch
  GradSrc src_d2;
  if (!last) {
    deep_bwd_level(P, l + 1, src_plain(v.dUp, 128, 0), ops);
    Level& n = P->lv[l + 1];
    if (CS == 128) { src_d2 = src_fold(n.dPin, 128, nullptr, nullptr, 0); src_d2.add = n.dS; src_d2.ld_add = 128; }
    else src_d2 = src_fold(n.dPin, 128, n.dRaw_s, P->params[n.p_skip_w], CS);
  } else {
    src_d2 = src_plain(v.dUp, 128, 0);
  }
  deep_bn_bwd(P, v.raw_d2, 128, v.bn_d2, src_d2, v.h, v.w, v.dRaw_d2, ops);
  ops.push_back(deep_conv(P, v.d2.dg, nullptr, nullptr, 0));
  ops.push_back(deep_wgrad(P, v.d2));
  deep_bn_bwd(P, v.raw_d1, 128, v.bn_d1, src_fold(v.dP_d1, 128, nullptr, nullptr, 0), v.h, v.w, v.dRaw_d1, ops);
  ops.push_back(deep_conv(P, v.d1.dg, nullptr, nullptr, 0));     // 4-phase stride-2 input gradient -> dPin
  ops.push_back(deep_wgrad(P, v.d1));
}
static int build_deep_ops(dip_plan* P) {
  P->n_deep_fwd = P->n_deep_bwd = 0;
  P->deep_grid = 128 < g_num_sms ? 128 : g_num_sms;
  if (P->desc.precision != DIP_PRECISION_TF32 || (int)P->lv.size() <= P->deep_from) return 0;
  for (const Level& v : P->lv)   // the op lists below are written for the 128-wide network with a skip branch at every scale
    if (v.nd != 128 || v.nu != 128 || v.ns == 0 || v.ns != P->lv[0].ns) return 0;
  std::vector<DeepOp> fwd;
  deep_fwd_level(P, P->deep_from, fwd);
  if ((int)fwd.size() > dip_plan::kDeepMaxOps) return fail("internal: deep forward op list too long");
  DIP_CUDA(cudaMemcpy(P->d_deep_fwd, fwd.data(), fwd.size() * sizeof(DeepOp), cudaMemcpyHostToDevice));
  P->n_deep_fwd = (int)fwd.size();
  if (getenv("DIP_NO_DEEP_BWD") == nullptr && P->lv[P->deep_from].d1.dg_s2) {
    std::vector<DeepOp> bwd;
    deep_bwd_level(P, P->deep_from, src_plain(P->lv[P->deep_from - 1].dUp, 128, 0), bwd);
    if ((int)bwd.size() > 2 * dip_plan::kDeepMaxOps) return fail("internal: deep backward op list too long");
    DIP_CUDA(cudaMemcpy(P->d_deep_bwd, bwd.data(), bwd.size() * sizeof(DeepOp), cudaMemcpyHostToDevice));
    P->n_deep_bwd = (int)bwd.size();
  }
  if (P->deep_grid > g_num_sms) P->deep_grid = g_num_sms;
  return 0;
}

static int plan_backward(dip_plan* P, const float* dout, cudaStream_t s) {
  if (!P->bound) return fail("dip_backward: parameters not bound");
  int nl = 0;
  P->deferred.clear();
  DIP_CUDA(cudaMemsetAsync(P->acc_bwd, 0, P->acc_bwd_n * sizeof(double), s));
  if (P->n_unpack > 0) DIP_CUDA(cudaMemsetAsync(P->wacc_base, 0, P->wacc_bytes, s));
  P->side_on = getenv("DIP_NO_SIDE") == nullptr;
  Level& v0 = P->lv[0];
  // RGB head backward (sigmoid', dgrad 3->128, wgrad, bias grad) is fused into the BN backward of the last stage
  GradSrc sh{};
  HBM_T(&P->timer, H_HEAD_DLOGIT, 0, (2.0 * P->desc.out_channels + 4.0) * P->H * P->W * sizeof(float), s,
        launch_head_dlogit(dout, P->out_saved, P->desc.out_channels, P->H * P->W, P->dl4, s, P->desc.need_sigmoid != 0));
  sh.kind = 3; sh.dl4 = P->dl4; sh.wh = P->params[P->p_head_w]; sh.nh = P->desc.out_channels;
  sh.dwh = P->dw_head; sh.dbh = P->db_head;
  nl += 1;
  DIP_CHECK(bwd_level(P, 0, sh, s, nl));
  DIP_CHECK(flush_deferred(P, P->desc.precision, s));
  join_side(P, s);
  join_skip(P, s);
  launch_k(k_cvt_table, dim3(P->n_cvt), dim3(128), 0, s, 1, P->d_cvt);
  nl += 1;
  if (P->n_unpack > 0) {
    HBM_T(&P->timer, H_WGRAD_REDUCE, 1, 2.0 * (double)P->wacc_bytes, s,
          launch_k(k_wgrad_unpack_table, dim3(64, P->n_unpack), dim3(256), 0, s, 1, P->d_unpack));
    nl += 2;   // + the accumulator memset
  }
  DIP_CUDA(cudaGetLastError());
  P->launches_bwd = nl;
  return 0;
}

}  // namespace dip

// ================================================================================================ C ABI
struct dip_adam {
  int n = 0;
  std::vector<long long> numel;
  int nblocks = 0;
  float** d_p = nullptr; const float** d_g = nullptr; float** d_m = nullptr; float** d_v = nullptr;
  int* d_blk_tensor = nullptr; int* d_blk_start = nullptr; int* d_numel = nullptr;
  bool bound = false;
  unsigned long long id = 0;        // unique per dip_adam_create (graph cache key of dip_run_iterations)
  unsigned long long bind_gen = 0;  // bumped by dip_adam_bind (the captured k_adam reads the tables, not their addresses,
                                    // but a re-bind after capture must still invalidate nothing else; kept for clarity)
};
static unsigned long long g_adam_serial = 0;

static void drop_graphs(dip_plan* P) {
  for (cudaGraphExec_t* g : {&P->gexec, &P->gfwd, &P->gbwd})
    if (*g != nullptr) { cudaGraphExecDestroy(*g); *g = nullptr; }
}
static bool graphs_enabled(const dip_plan* P) { return !P->timer.on && getenv("DIP_NO_GRAPH") == nullptr; }
static int ensure_gstream(dip_plan* P) {
  // the legacy default stream cannot be captured: graphs replay on a private stream, ordered against the caller's stream
  if (P->gstream == nullptr) {
    int lo = 0, hi = 0;
    DIP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (getenv("DIP_NO_PRIO") != nullptr) hi = 0;
    DIP_CUDA(cudaStreamCreateWithPriority(&P->gstream, cudaStreamNonBlocking, hi));
    DIP_CUDA(cudaEventCreateWithFlags(&P->gev_in, cudaEventDisableTiming));
    DIP_CUDA(cudaEventCreateWithFlags(&P->gev_out, cudaEventDisableTiming));
  }
  return 0;
}
// Captures body(gstream) once into *exec, then replays it between two event edges to / from the caller's stream `s`.
template <class Body>
static int replay_graph(dip_plan* P, cudaGraphExec_t* exec, cudaStream_t s, Body body) {
  DIP_CHECK(ensure_gstream(P));
  cudaStream_t gs = P->gstream;
  DIP_CUDA(cudaEventRecord(P->gev_in, s));
  DIP_CUDA(cudaStreamWaitEvent(gs, P->gev_in, 0));
  if (*exec == nullptr) {
    cudaGraph_t graph = nullptr;
    DIP_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    const int rc = body(gs);
    cudaError_t ce = cudaStreamEndCapture(gs, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { *exec = nullptr; return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
  }
  DIP_CUDA(cudaGraphLaunch(*exec, gs));
  DIP_CUDA(cudaEventRecord(P->gev_out, gs));
  DIP_CUDA(cudaStreamWaitEvent(s, P->gev_out, 0));
  return 0;
}

extern "C" {

const char* dip_last_error(void) { return g_err.c_str(); }
int dip_version(void) { return 100; }

size_t dip_plan_workspace_bytes(const dip_net_desc* desc, int H, int W) {
  dip_plan P;
  P.desc = *desc; P.H = H; P.W = W; P.dry = true;
  Arena A{nullptr};
  if (build_plan(&P, A) != 0) return 0;
  return A.off + 4096;
}

int dip_plan_create(const dip_net_desc* desc, int H, int W, void* workspace, size_t workspace_bytes, dip_plan** out) {
  DIP_CHECK(engine_init());
  const size_t need = dip_plan_workspace_bytes(desc, H, W);
  if (need == 0) return -1;
  if (workspace == nullptr || workspace_bytes < need) return fail("dip_plan_create: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return fail("dip_plan_create: workspace must be 256-byte aligned");
  dip_plan* P = new dip_plan();
  P->desc = *desc; P->H = H; P->W = W; P->ws = (uint8_t*)workspace; P->ws_bytes = workspace_bytes;
  Arena A{(uint8_t*)workspace};
  if (build_plan(P, A) != 0) { delete P; return -1; }
  *out = P;
  return 0;
}
void dip_plan_destroy(dip_plan* plan) {
  if (plan == nullptr) return;
  if (plan->gexec) cudaGraphExecDestroy(plan->gexec);
  if (plan->gfwd) cudaGraphExecDestroy(plan->gfwd);
  if (plan->gbwd) cudaGraphExecDestroy(plan->gbwd);
  if (plan->gstream) cudaStreamDestroy(plan->gstream);
  if (plan->gev_in) cudaEventDestroy(plan->gev_in);
  if (plan->gev_out) cudaEventDestroy(plan->gev_out);
  for (cudaEvent_t e : plan->timer.pool) cudaEventDestroy(e);
  for (cudaEvent_t e : plan->wev) cudaEventDestroy(e);
  if (plan->wstream) cudaStreamDestroy(plan->wstream);
  if (plan->sstream) cudaStreamDestroy(plan->sstream);
  delete plan;
}
int dip_plan_num_params(const dip_plan* plan) { return (int)plan->numel.size(); }
int dip_plan_num_bn(const dip_plan* plan) { return (int)plan->bns.size(); }
long long dip_plan_param_numel(const dip_plan* plan, int index) {
  if (index < 0 || index >= (int)plan->numel.size()) return -1;
  return plan->numel[index];
}
int dip_plan_bind(dip_plan* P, void* const* params, void* const* grads, void* const* bn_running, int nbt_is_float) {
  drop_graphs(P);   // captured launches hold the old pointers
  P->nbt_is_float = nbt_is_float;
  const int n = (int)P->numel.size();
  P->params.resize(n); P->grads.resize(n);
  for (int i = 0; i < n; ++i) {
    if (params[i] == nullptr || grads[i] == nullptr) return fail("dip_plan_bind: null parameter/gradient pointer");
    P->params[i] = (float*)params[i]; P->grads[i] = (float*)grads[i];
  }
  P->running.clear();
  if (bn_running != nullptr) P->running.assign(bn_running, bn_running + 3 * P->bns.size());
  DIP_CHECK(upload_tables(P));
  DIP_CHECK(build_deep_ops(P));
  P->bound = true;
  return 0;
}
int dip_forward(dip_plan* P, const void* z, const void* noise, float sigma, void* out, dip_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!P->bound) return fail("dip_forward: parameters not bound (call dip_plan_bind)");
  if (noise != nullptr || !graphs_enabled(P))
    return plan_forward(P, (const float*)z, (const float*)noise, sigma, (float*)out, s);
  // stage the input in the plan's fixed buffer, replay the captured forward, hand the result out
  const size_t nz = (size_t)P->H * P->W * P->desc.in_channels * sizeof(float);
  DIP_CUDA(cudaMemcpyAsync(P->zbuf, z, nz, cudaMemcpyDeviceToDevice, s));
  DIP_CHECK(replay_graph(P, &P->gfwd, s, [&](cudaStream_t gs) { return plan_forward(P, P->zbuf, nullptr, 0.f, nullptr, gs); }));
  if (out != nullptr && out != P->out_saved)
    DIP_CUDA(cudaMemcpyAsync(out, P->out_saved, (size_t)P->H * P->W * P->desc.out_channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}
int dip_backward(dip_plan* P, const void* dout, dip_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!P->bound) return fail("dip_backward: parameters not bound");
  if (!graphs_enabled(P)) return plan_backward(P, (const float*)dout, s);
  if (dout != P->dout)
    DIP_CUDA(cudaMemcpyAsync(P->dout, dout, (size_t)P->H * P->W * P->desc.out_channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return replay_graph(P, &P->gbwd, s, [&](cudaStream_t gs) { return plan_backward(P, P->dout, gs); });
}
int dip_loss_mse(const void* out, const void* target, const void* mask, int channels, int hw, double* loss, void* dout,
                 dip_stream_t stream) {
  launch_mse((const float*)out, (const float*)target, (const float*)mask, channels, hw, loss, (float*)dout, nullptr, (cudaStream_t)stream);
  DIP_CUDA(cudaGetLastError());
  return 0;
}
int dip_noise_perturb(const void* z0, void* z, float sigma, uint64_t seed, uint64_t offset, size_t n, dip_stream_t stream) {
  if (n % 4 != 0) return fail("dip_noise_perturb: n must be a multiple of 4");
  launch_noise((const float*)z0, (float*)z, sigma, seed, offset, nullptr, n, (cudaStream_t)stream);
  DIP_CUDA(cudaGetLastError());
  return 0;
}

int dip_lanczos_down_fwd(const void* x, int C, int H, int W, const void* kern, int K, int factor, int pad, void* y,
                         dip_stream_t stream) {
  DIP_CHECK(engine_init());
  if (K < 1 || factor < 1 || pad < 0) return fail("dip_lanczos_down_fwd: bad K / factor / pad");
  if (down_out_size(H, K, factor, pad) < 1 || down_out_size(W, K, factor, pad) < 1)
    return fail("dip_lanczos_down_fwd: image smaller than the filter");
  DIP_CUDA(launch_down_fwd((const float*)x, C, H, W, (const float*)kern, K, factor, pad, (float*)y, (cudaStream_t)stream));
  return 0;
}
int dip_lanczos_down_bwd(const void* dy, int C, int H, int W, const void* kern, int K, int factor, int pad, void* dx,
                         dip_stream_t stream) {
  DIP_CHECK(engine_init());
  if (K < 1 || factor < 1 || pad < 0) return fail("dip_lanczos_down_bwd: bad K / factor / pad");
  if (down_out_size(H, K, factor, pad) < 1 || down_out_size(W, K, factor, pad) < 1)
    return fail("dip_lanczos_down_bwd: image smaller than the filter");
  DIP_CUDA(launch_down_bwd((const float*)dy, C, H, W, (const float*)kern, K, factor, pad, (float*)dx, (cudaStream_t)stream));
  return 0;
}
int dip_lanczos_down_out_size(int n, int K, int factor, int pad) { return down_out_size(n, K, factor, pad); }

int dip_plan_set_downsampler(dip_plan* P, const float* kern_host, int K, int factor, int pad) {
  if (P->gexec != nullptr) { cudaGraphExecDestroy(P->gexec); P->gexec = nullptr; }   // the captured step changes
  if (kern_host == nullptr || K == 0) { P->ds_K = 0; return 0; }
  if (K < 1 || K > dip_plan::kDownMaxK || factor < 1 || pad < 0) return fail("dip_plan_set_downsampler: bad K / factor / pad");
  const int Ho = down_out_size(P->H, K, factor, pad), Wo = down_out_size(P->W, K, factor, pad);
  if (Ho < 1 || Wo < 1 || Ho > P->H || Wo > P->W) return fail("dip_plan_set_downsampler: unsupported output size");
  DIP_CUDA(cudaMemcpy(P->ds_kern, kern_host, (size_t)K * K * sizeof(float), cudaMemcpyHostToDevice));
  P->ds_K = K; P->ds_f = factor; P->ds_pad = pad; P->ds_Ho = Ho; P->ds_Wo = Wo;
  return 0;
}

int dip_adam_create(int ntensors, const long long* numel, dip_adam** out) {
  dip_adam* a = new dip_adam();
  a->id = ++g_adam_serial;
  a->n = ntensors;
  a->numel.assign(numel, numel + ntensors);
  std::vector<int> bt, bs, ne;
  const int chunk = adam_chunk();
  for (int t = 0; t < ntensors; ++t) {
    ne.push_back((int)numel[t]);
    for (long long st = 0; st < numel[t]; st += chunk) { bt.push_back(t); bs.push_back((int)st); }
  }
  a->nblocks = (int)bt.size();
  DIP_CUDA(cudaMalloc(&a->d_p, ntensors * sizeof(void*)));
  DIP_CUDA(cudaMalloc(&a->d_g, ntensors * sizeof(void*)));
  DIP_CUDA(cudaMalloc(&a->d_m, ntensors * sizeof(void*)));
  DIP_CUDA(cudaMalloc(&a->d_v, ntensors * sizeof(void*)));
  DIP_CUDA(cudaMalloc(&a->d_blk_tensor, bt.size() * sizeof(int)));
  DIP_CUDA(cudaMalloc(&a->d_blk_start, bs.size() * sizeof(int)));
  DIP_CUDA(cudaMalloc(&a->d_numel, ne.size() * sizeof(int)));
  DIP_CUDA(cudaMemcpy(a->d_blk_tensor, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice));
  DIP_CUDA(cudaMemcpy(a->d_blk_start, bs.data(), bs.size() * sizeof(int), cudaMemcpyHostToDevice));
  DIP_CUDA(cudaMemcpy(a->d_numel, ne.data(), ne.size() * sizeof(int), cudaMemcpyHostToDevice));
  *out = a;
  return 0;
}
void dip_adam_destroy(dip_adam* a) {
  if (!a) return;
  cudaFree(a->d_p); cudaFree(a->d_g); cudaFree(a->d_m); cudaFree(a->d_v);
  cudaFree(a->d_blk_tensor); cudaFree(a->d_blk_start); cudaFree(a->d_numel);
  delete a;
}
int dip_adam_bind(dip_adam* a, void* const* p, void* const* g, void* const* m, void* const* v) {
  DIP_CUDA(cudaMemcpy(a->d_p, p, a->n * sizeof(void*), cudaMemcpyHostToDevice));
  DIP_CUDA(cudaMemcpy(a->d_g, g, a->n * sizeof(void*), cudaMemcpyHostToDevice));
  DIP_CUDA(cudaMemcpy(a->d_m, m, a->n * sizeof(void*), cudaMemcpyHostToDevice));
  DIP_CUDA(cudaMemcpy(a->d_v, v, a->n * sizeof(void*), cudaMemcpyHostToDevice));
  a->bound = true;
  a->bind_gen++;
  return 0;
}
int dip_adam_step(dip_adam* a, double lr, double beta1, double beta2, double eps, int step, dip_stream_t stream) {
  if (!a->bound) return fail("dip_adam_step: not bound");
  if (step < 1) return fail("dip_adam_step: step must be >= 1");
  AdamTable t{a->d_p, a->d_g, a->d_m, a->d_v, a->d_blk_tensor, a->d_blk_start, a->d_numel, a->nblocks};
  launch_adam(t, lr, beta1, beta2, eps, step, nullptr, (cudaStream_t)stream);
  DIP_CUDA(cudaGetLastError());
  return 0;
}

// One iteration of the lean closure: noise -> forward -> MSE -> backward -> Adam.  With it_dev != nullptr every
// per-iteration scalar (Philox stream, loss slot, Adam step) comes from device counters, so the launch sequence is
// identical for every iteration and can be replayed as a CUDA graph.
static int run_body(dip_plan* P, dip_adam* adam, const float* z0, const float* target, const float* mask, float sigma,
                    uint64_t seed, int step_base, double lr, float* out, double* loss_slot, int* it_dev, cudaStream_t s) {
  const size_t nz = (size_t)P->H * P->W * P->desc.in_channels;
  const int hw = P->H * P->W;
  const float* zin = z0;
  P->side_on = getenv("DIP_NO_SIDE") == nullptr;
  P->wev_used = 0;
  if (P->side_on && P->bound) {
    // weight repack on the side stream from the very start of the iteration (beside the noise kernel); plan_forward joins it
    plan_pack(P, fork_side(P, s));
    P->prepacked = true;
  }
  static const bool split_noise = getenv("DIP_SPLIT_NOISE") != nullptr;   // A/B switch: separate k_noise + k_input_pad launches
  if (sigma > 0.f && (split_noise || P->W % 4 != 0)) {
    HBM_T(&P->timer, H_NOISE, 0, 2.0 * nz * sizeof(float), s, launch_noise(z0, P->zbuf, sigma, seed, (uint64_t)step_base, it_dev, nz, s));
    zin = P->zbuf;
  } else if (sigma > 0.f) {
    P->fnoise.on = true; P->fnoise.z0 = z0; P->fnoise.sigma = sigma; P->fnoise.seed = seed; P->fnoise.offset = (uint64_t)step_base;
    P->fnoise.it_dev = it_dev;
  }
  const int frc = plan_forward(P, zin, nullptr, 0.f, out, s);
  P->fnoise.on = false;
  DIP_CHECK(frc);
  const int* slot_idx = it_dev != nullptr ? it_dev + 1 : nullptr;
  if (P->ds_K > 0) {
    // super-resolution: loss on the downsampled output (super-resolution.ipynb c10:8-11); the operator's adjoint
    // turns the low-resolution loss gradient into dL/d(out)
    const int co = P->desc.out_channels;
    HBM_T(&P->timer, H_DOWN_FWD, 0, (double)co * ((double)P->H * P->W + (double)P->ds_Ho * P->ds_Wo) * sizeof(float), s,
          DIP_CUDA(launch_down_fwd(P->out_saved, co, P->H, P->W, P->ds_kern, P->ds_K, P->ds_f, P->ds_pad, P->ds_y, s)));
    launch_mse(P->ds_y, target, mask, co, P->ds_Ho * P->ds_Wo, loss_slot, P->ds_dy, slot_idx, s);
    HBM_T(&P->timer, H_DOWN_BWD, 0, (double)co * ((double)P->H * P->W + (double)P->ds_Ho * P->ds_Wo) * sizeof(float), s,
          DIP_CUDA(launch_down_bwd(P->ds_dy, co, P->H, P->W, P->ds_kern, P->ds_K, P->ds_f, P->ds_pad, P->dout, s)));
  } else {
    HBM_T(&P->timer, H_MSE, mask != nullptr, (3.0 * P->desc.out_channels + (mask != nullptr ? 1.0 : 0.0)) * hw * sizeof(float), s,
          launch_mse(P->out_saved, target, mask, P->desc.out_channels, hw, loss_slot, P->dout, slot_idx, s));
  }
  DIP_CHECK(plan_backward(P, P->dout, s));
  if (!adam->bound) return fail("dip_run_iterations: adam not bound");
  AdamTable t{adam->d_p, adam->d_g, adam->d_m, adam->d_v, adam->d_blk_tensor, adam->d_blk_start, adam->d_numel, adam->nblocks};
  {
    double np_ = 0; for (long long n : adam->numel) np_ += (double)n;
    HBM_T(&P->timer, H_ADAM, 0, 7.0 * np_ * sizeof(float), s, launch_adam(t, lr, 0.9, 0.999, 1e-8, step_base + 1, it_dev, s));
  }
  if (it_dev != nullptr) launch_advance(it_dev, s);
  DIP_CUDA(cudaGetLastError());
  return 0;
}

int dip_run_iterations(dip_plan* P, dip_adam* adam, const void* z0, const void* target, const void* mask, float sigma,
                       uint64_t seed, int step0, int iters, double lr, void* out, double* loss_hist, dip_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (iters <= 0) return 0;
  const bool use_graph = !P->timer.on && getenv("DIP_NO_GRAPH") == nullptr;
  if (!use_graph) {
    for (int i = 0; i < iters; ++i) {
      double* lp = loss_hist != nullptr ? loss_hist + i : P->loss_ring;
      DIP_CUDA(cudaMemsetAsync(lp, 0, sizeof(double), s));
      DIP_CHECK(run_body(P, adam, (const float*)z0, (const float*)target, (const float*)mask, sigma, seed, step0 + i, lr,
                         (float*)out, lp, nullptr, s));
    }
    return 0;
  }
  if (loss_hist == nullptr && iters > dip_plan::kLossRing) {
    // internal ring too small: split the call
    for (int done = 0; done < iters; done += dip_plan::kLossRing) {
      const int n = iters - done < dip_plan::kLossRing ? iters - done : dip_plan::kLossRing;
      DIP_CHECK(dip_run_iterations(P, adam, z0, target, mask, sigma, seed, step0 + done, n, lr, out, nullptr, stream));
    }
    return 0;
  }
  // the legacy default stream cannot be captured: replay on a private stream, ordered against the caller's stream
  DIP_CHECK(ensure_gstream(P));
  cudaStream_t gs = P->gstream;
  DIP_CUDA(cudaEventRecord(P->gev_in, s));
  DIP_CUDA(cudaStreamWaitEvent(gs, P->gev_in, 0));
  double* slots = loss_hist != nullptr ? loss_hist : P->loss_ring;
  dip_plan::GraphKey key;
  key.z0 = z0; key.target = target; key.mask = mask; key.out = out; key.slots = slots;
  key.adam_id = adam->id; key.adam_bind = adam->bind_gen; key.sigma = sigma; key.seed = seed; key.lr = lr;
  if (P->gexec == nullptr || !(key == P->gkey)) {
    if (P->gexec != nullptr) { cudaGraphExecDestroy(P->gexec); P->gexec = nullptr; }
    cudaGraph_t graph = nullptr;
    DIP_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    const int rc = run_body(P, adam, (const float*)z0, (const float*)target, (const float*)mask, sigma, seed, 0, lr, (float*)out,
                            slots, P->it_dev, gs);
    cudaError_t ce = cudaStreamEndCapture(gs, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&P->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { P->gexec = nullptr; return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
    P->gkey = key;
  }
  const int init[2] = {step0, 0};
  DIP_CUDA(cudaMemcpyAsync(P->it_dev, init, sizeof init, cudaMemcpyHostToDevice, gs));
  DIP_CUDA(cudaMemsetAsync(slots, 0, (size_t)iters * sizeof(double), gs));
  for (int i = 0; i < iters; ++i) DIP_CUDA(cudaGraphLaunch(P->gexec, gs));
  DIP_CUDA(cudaEventRecord(P->gev_out, gs));
  DIP_CUDA(cudaStreamWaitEvent(s, P->gev_out, 0));
  return 0;
}

int dip_input_grad(dip_plan* P, void* dz, dip_stream_t stream) {
  if (!P->desc.input_grad) return fail("dip_input_grad: the plan was created without input_grad");
  Level& v = P->lv[0];
  launch_input_grad(v.dPin, v.ns > 0 ? v.dS : nullptr, v.Cin, v.Cin_act, v.H, v.W, (float*)dz, (cudaStream_t)stream);
  DIP_CUDA(cudaGetLastError());
  return 0;
}

int dip_plan_buffer(const dip_plan* plan, const char* name, void** ptr, int* dims4) {
  auto it = plan->bufs.find(name);
  if (it == plan->bufs.end()) return fail(std::string("dip_plan_buffer: unknown buffer ") + name);
  *ptr = it->second.ptr;
  dims4[0] = it->second.rows; dims4[1] = it->second.cols; dims4[2] = it->second.ld; dims4[3] = it->second.c;
  return 0;
}
int dip_plan_set_timing(dip_plan* plan, int enable) {
  plan->timer.on = enable != 0;
  plan->timer.reset();
  return 0;
}
int dip_plan_get_timing(dip_plan* plan, double* ms3, double* flops3, int* launches3) {
  for (int i = 0; i < 3; ++i) { ms3[i] = 0; flops3[i] = 0; launches3[i] = 0; }
  for (const Timer::Rec& r : plan->timer.recs) {
    DIP_CUDA(cudaEventSynchronize(plan->timer.pool[r.e1]));
    float ms = 0.f;
    DIP_CUDA(cudaEventElapsedTime(&ms, plan->timer.pool[r.e0], plan->timer.pool[r.e1]));
    if (r.cls > 2) continue;   // HBM-bound kernels (class >= 16) are reported through dip_plan_get_timing_records only
    ms3[r.cls] += ms; flops3[r.cls] += r.flops; launches3[r.cls] += 1;
  }
  plan->timer.reset();
  return 0;
}
int dip_plan_get_timing_records(dip_plan* plan, int max_records, int* cls, double* flops, double* ms) {
  int n = 0;
  for (const Timer::Rec& r : plan->timer.recs) {
    if (n >= max_records) break;
    DIP_CUDA(cudaEventSynchronize(plan->timer.pool[r.e1]));
    float t = 0.f;
    DIP_CUDA(cudaEventElapsedTime(&t, plan->timer.pool[r.e0], plan->timer.pool[r.e1]));
    cls[n] = r.cls; flops[n] = r.flops; ms[n] = t;
    ++n;
  }
  plan->timer.reset();
  return n;
}
int dip_plan_num_launches(const dip_plan* plan, int* fwd, int* bwd) {
  *fwd = plan->launches_fwd; *bwd = plan->launches_bwd;
  return 0;
}

// ---------------------------------------------------------------------------------------------- single-op entry points
size_t dip_op_scratch_bytes(void) { return (size_t)96 << 20; }

// scratch layout of the single-op entry points: [fprop pack | dgrad pack | one PackEntry (64 floats) | wgrad accumulator |
// bf16 copies of the operands (precision bf16)]
static float* op_partial(ConvOp& op, float* scratch) {
  return scratch + ((op.wp_f_elems() + 63) & ~size_t(63)) + ((op.wp_d_elems() + 63) & ~size_t(63)) + 64;
}
static int op_common(ConvOp& op, int N, int C, int k, int stride, int rot, float* scratch, const float* w, cudaStream_t s,
                     int max_c = 160, int prec = DIP_PRECISION_TF32) {
  if (N != 128) return fail("dip_op_conv_*: N must be 128");
  if (C % 4 != 0 || C > max_c) return fail("dip_op_conv_*: C must be a multiple of 4 and <= " + std::to_string(max_c));
  op.N = N; op.C = C; op.k = k; op.stride = stride; op.rot = rot;
  op.bf16 = prec == DIP_PRECISION_BF16;
  op.set_shapes();
  op.wp_f = scratch;
  op.wp_d = scratch + ((op.wp_f_elems() + 63) & ~size_t(63));
  if (op.bf16) {
    // bf16 packs come from the table kernel (one entry, staged in the scratch area)
    PackEntry e{};
    e.w = w; e.dst_f = op.wp_f; e.dst_d = op.wp_d; e.N = N; e.C = C; e.k = k; e.rot = rot; e.n_rows = N;
    e.c_pad = op.c_pad; e.c_rows = op.crows; e.Ctot = C; e.coff = 0; e.s2 = 0; e.bf16 = 1; e.c_pad16 = op.c_pad16;
    PackEntry* d_e = reinterpret_cast<PackEntry*>(op_partial(op, scratch) - 64);
    DIP_CUDA(cudaMemcpyAsync(d_e, &e, sizeof e, cudaMemcpyHostToDevice, s));
    launch_k(k_pack_table, dim3(64, 1), dim3(256), 0, s, 1, (const PackEntry*)d_e);
  } else {
    launch_pack_fprop(w, N, C, k, k, rot, op.wp_f, N, op.c_pad, s);
    launch_pack_dgrad(w, N, C, k, k, rot, op.wp_d, op.crows, 128, s);
  }
  DIP_CUDA(cudaGetLastError());
  return 0;
}
// precision bf16: bf16 copy of an NHWC fp32 operand [npix][ld] in the scratch area behind *area (advanced)
static const uint16_t* op_cast(const void* x, int ld, long long npix, uint16_t** area, int* ld16, cudaStream_t s) {
  *ld16 = round_up(ld, 8);
  uint16_t* dst = *area;
  *area += ((size_t)npix * *ld16 + 127) & ~size_t(127);
  launch_cast_bf16((const float*)x, ld, ld, npix, Twin{dst, *ld16}, s);
  return dst;
}
static uint16_t* op_cast_area(ConvOp& op, float* partial) {
  return reinterpret_cast<uint16_t*>(partial + ((op.wacc_elems() + 63) & ~size_t(63)));
}

int dip_op_conv_fprop(const void* a, int a_h, int a_w, int a_c, const void* w, const void* bias, int N, int C, int k, int stride,
                      int offx, int offy, int rot, void* d, int d_h, int d_w, double* stats, int precision, void* scratch,
                      dip_stream_t stream) {
  DIP_CHECK(engine_init());
  cudaStream_t s = (cudaStream_t)stream;
  ConvOp op;
  DIP_CHECK(op_common(op, N, C, k, stride, rot, (float*)scratch, (const float*)w, s, 256, precision));
  op.in = (const float*)a; op.in_rows = a_h; op.in_cols = a_w; op.in_ld = a_c; op.offx = offx; op.offy = offy;
  op.out = (float*)d; op.out_h = d_h; op.out_w = d_w; op.stats = stats;
  op.has_dgrad = false;
  op.wg_dy = (const float*)d; op.wg_h = d_h; op.wg_w = d_w;  // placeholders so that build_tc can encode maps
  if (op.bf16) {
    uint16_t* area = op_cast_area(op, op_partial(op, (float*)scratch));
    op.in16 = op_cast(a, a_c, (long long)a_h * a_w, &area, &op.in_ld16, s);
    op.do_wgrad = false;
  }
  if (is_tc(precision)) DIP_CHECK(op.build_tc(op_partial(op, (float*)scratch)));
  return op.run_fprop(precision, (const float*)bias, s);
}
int dip_op_conv_dgrad(const void* dy, int dy_h, int dy_w, const void* w, int N, int C, int k, int rot, void* dx, int dx_h, int dx_w,
                      int precision, void* scratch, dip_stream_t stream) {
  DIP_CHECK(engine_init());
  cudaStream_t s = (cudaStream_t)stream;
  ConvOp op;
  DIP_CHECK(op_common(op, N, C, k, 1, rot, (float*)scratch, (const float*)w, s, 160, precision));
  // fprop/wgrad placeholders (valid maps over the same buffers; not launched)
  op.in = (const float*)dx; op.in_rows = dx_h; op.in_cols = dx_w; op.in_ld = C; op.out = (float*)const_cast<void*>(dy);
  op.out_h = dy_h; op.out_w = dy_w; op.wg_dy = (const float*)dy; op.wg_h = dy_h; op.wg_w = dy_w;
  op.has_dgrad = true;
  op.dg_in = (const float*)dy; op.dg_in_h = dy_h; op.dg_in_w = dy_w;
  op.dg_out = (float*)dx; op.dg_out_h = dx_h; op.dg_out_w = dx_w; op.dg_off = -(k - 1);
  if (op.bf16) {
    uint16_t* area = op_cast_area(op, op_partial(op, (float*)scratch));
    int ld16 = 0;
    op.dg_in16 = op_cast(dy, 128, (long long)dy_h * dy_w, &area, &ld16, s);
    op.do_fprop = false; op.do_wgrad = false;
  }
  if (is_tc(precision)) DIP_CHECK(op.build_tc(op_partial(op, (float*)scratch)));
  return op.run_dgrad(precision, s);
}
int dip_op_conv_dgrad_s2(const void* dy, int dy_h, int dy_w, const void* w, int N, int C, int rot, void* dx, int precision,
                         void* scratch, dip_stream_t stream) {
  DIP_CHECK(engine_init());
  if (!is_tc(precision)) return fail("dip_op_conv_dgrad_s2: tensor-core path only (the exact-fp32 mode zero-stuffs)");
  cudaStream_t s = (cudaStream_t)stream;
  ConvOp op;
  if (N != 128) return fail("dip_op_conv_dgrad_s2: N must be 128");
  if (C % 4 != 0 || C > 160) return fail("dip_op_conv_dgrad_s2: C must be a multiple of 4 and <= 160");
  op.N = N; op.C = C; op.k = 3; op.stride = 2; op.rot = rot;
  op.bf16 = precision == DIP_PRECISION_BF16;
  op.set_shapes();
  op.wp_f = (float*)scratch;
  op.wp_d = (float*)scratch + ((op.wp_f_elems() + 63) & ~size_t(63));
  // one-entry pack table in the scratch area behind the packed weights
  PackEntry e{};
  e.w = (const float*)w; e.dst_f = nullptr; e.dst_d = op.wp_d; e.N = N; e.C = C; e.k = 3; e.rot = rot; e.n_rows = N;
  e.c_pad = op.c_pad; e.c_rows = op.crows; e.Ctot = C; e.coff = 0; e.s2 = 1; e.bf16 = op.bf16 ? 1 : 0; e.c_pad16 = op.c_pad16;
  PackEntry* d_e = reinterpret_cast<PackEntry*>(op_partial(op, (float*)scratch) - 64);
  DIP_CUDA(cudaMemcpyAsync(d_e, &e, sizeof e, cudaMemcpyHostToDevice, s));
  launch_k(k_pack_table, dim3(64, 1), dim3(256), 0, s, 1, (const PackEntry*)d_e);
  op.do_fprop = false; op.do_wgrad = false;
  op.has_dgrad = true; op.dg_s2 = true;
  op.dg_in = (const float*)dy; op.dg_in_h = dy_h; op.dg_in_w = dy_w;
  op.dg_out = (float*)dx; op.dg_out_h = 2 * dy_h + 2; op.dg_out_w = 2 * dy_w + 2; op.dg_off = -2;
  if (op.bf16) {
    uint16_t* area = op_cast_area(op, op_partial(op, (float*)scratch));
    int ld16 = 0;
    op.dg_in16 = op_cast(dy, 128, (long long)dy_h * dy_w, &area, &ld16, s);
  }
  DIP_CHECK(op.build_tc(nullptr));
  return op.run_dgrad(precision, s);
}
int dip_op_conv_wgrad(const void* dy, int dy_h, int dy_w, const void* a, int a_h, int a_w, int a_c, int N, int C, int k, int stride,
                      int offx, int offy, int rot, void* dw, int precision, void* scratch, dip_stream_t stream) {
  DIP_CHECK(engine_init());
  cudaStream_t s = (cudaStream_t)stream;
  ConvOp op;
  // weights are not needed for wgrad; pack from dw is skipped
  if (N != 128) return fail("dip_op_conv_wgrad: N must be 128");
  op.N = N; op.C = C; op.k = k; op.stride = stride; op.rot = rot;
  op.bf16 = precision == DIP_PRECISION_BF16;
  op.set_shapes();
  op.wp_f = (float*)scratch;
  op.wp_d = nullptr;
  op.in = (const float*)a; op.in_rows = a_h; op.in_cols = a_w; op.in_ld = a_c; op.offx = offx; op.offy = offy;
  op.out = (float*)const_cast<void*>(dy); op.out_h = dy_h; op.out_w = dy_w;
  op.has_dgrad = false;
  op.wg_dy = (const float*)dy; op.wg_h = dy_h; op.wg_w = dy_w;
  op.simt_ksplits = dy_h < 64 ? dy_h : 64;
  float* partial = (float*)scratch + ((op.wp_f_elems() + 63) & ~size_t(63));
  if (op.bf16) {
    uint16_t* area = op_cast_area(op, partial);
    int ld16 = 0;
    op.wg_dy16 = op_cast(dy, 128, (long long)dy_h * dy_w, &area, &ld16, s);
    op.in16 = op_cast(a, a_c, (long long)a_h * a_w, &area, &op.in_ld16, s);
    op.do_fprop = false;
    if ((uint8_t*)area > (uint8_t*)scratch + dip_op_scratch_bytes()) return fail("dip_op_conv_wgrad: operands too large for the scratch area (bf16 copies)");
  }
  if (is_tc(precision)) DIP_CHECK(op.build_tc(partial));
  return op.run_wgrad(precision, partial, (float*)dw, s);
}

}  // extern "C"
