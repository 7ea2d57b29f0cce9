// C ABI of the B200-native MSDeformAttn operator (declared in include/msda_b200.h).
// Torch-free: plain pointers, sizes and a cudaStream_t.  Compiled for sm_100a only.
//
// Replaces, in the reference tree (/root/reference/models/ops/src):
//   cuda/ms_deform_im2col_cuda.cuh:923-954   ms_deformable_im2col_cuda<T>   -> msda_fwd_*
//   cuda/ms_deform_im2col_cuda.cuh:956-1327  ms_deformable_col2im_cuda<T>   -> msda_bwd_*
//   cuda/ms_deform_attn_cuda.cu:50-52        chunking precondition          -> msda_b200_check_im2col_step
#include "../../include/msda_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "msda_d32.cuh"
#include "msda_d32_tiles.cuh"
#include "msda_d32_wide.cuh"
#include "msda_d32_slab.cuh"
#include "msda_generic.cuh"

namespace {

using namespace msda;

std::atomic<int> g_fwd_path{0};
std::atomic<int> g_bwd_path{0};
std::atomic<int> g_bf16_direct{0};
std::atomic<int> g_split_mode{0};   // 0 auto, 1 never, 2 always
std::atomic<int> g_fwd_slab{0};     // shared-memory slab forward for small pyramids (fp32, L = P = 4): 0 by shape, 1 never, 2 whenever eligible
std::atomic<int> g_fwd_wide{0};     // 8-channels-per-lane forward (L = P = 4): 0 by shape, 1 never, 2 whenever eligible
std::atomic<int> g_bwd_tiles{0};    // on-chip-combining backward (L = P = 4): 0 by shape, 1 never, 2 whenever eligible
std::atomic<long long> g_launches{0};
// Introspection only (bench.py / tests read it right after their own call); last writer wins.
std::atomic<const char*> g_last_path[2] = {{"none"}, {"none"}};

// SM count of the CURRENT device (one process may drive several devices; cached per device).
int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;
  }
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  int n = cached[slot].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;
  }
  cached[slot].store(n, std::memory_order_relaxed);
  return n;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int check_dims(int N, int S, int M, int D, int L, int Lq, int P) {
  if (N < 0 || S < 0 || M < 0 || D < 0 || L < 0 || Lq < 0 || P < 0) return MSDA_ERR_BAD_DIMS;
  // Flat element counts must fit in int64 comfortably; each factor is < 2^31 so two-by-two
  // products are safe, and we bound the totals at 2^46 elements (far beyond 180 GB).
  const long double lim = 70368744177664.0L;
  if ((long double)N * S * M * D > lim) return MSDA_ERR_BAD_DIMS;
  if ((long double)N * Lq * M * L * P * 2 > lim) return MSDA_ERR_BAD_DIMS;
  if ((long double)N * Lq * M * D > lim) return MSDA_ERR_BAD_DIMS;
  return MSDA_OK;
}

bool nothing_to_do(int N, int M, int D, int Lq) { return N == 0 || M == 0 || D == 0 || Lq == 0; }

// Checked build only: tell the kernels the extent of value / grad_value for the launch that follows
// (debug facility: one launch at a time per process).
void register_extents(const void* value, size_t value_bytes, const void* gv, size_t gv_bytes, cudaStream_t st) {
#if MSDA_BOUNDS_CHECK
  const char* lo[2] = {(const char*)value, (const char*)gv};
  const char* hi[2] = {(const char*)value + value_bytes, (const char*)gv + gv_bytes};
  cudaStreamSynchronize(st);
  cudaMemcpyToSymbol(g_chk_lo, lo, sizeof(lo));
  cudaMemcpyToSymbol(g_chk_hi, hi, sizeof(hi));
#else
  (void)value; (void)value_bytes; (void)gv; (void)gv_bytes; (void)st;
#endif
}

// A stale, non-sticky error left behind by the host framework or an earlier caller must not be reported as this
// call's: every entry point that launches clears the error state first (begin_call) and reports only what its own
// launches produced (finish_launch).
void begin_call() { cudaGetLastError(); }

int finish_launch(int n_launched) {
  g_launches.fetch_add(n_launched, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MSDA_OK : (int)e;
}

// Backward with no value tokens / levels / points: every sample contributes nothing, so the two gradients that the
// header promises to be "fully written" are written as zeros (grad_value stays as the caller zeroed it).
int zero_sample_grads(void* gl, void* ga, size_t elem, int N, int Lq, int M, int L, int P, cudaStream_t st) {
  const size_t n = (size_t)N * Lq * M * L * P;
  if (n == 0) return MSDA_OK;
  if (!gl || !ga) return MSDA_ERR_NULL_POINTER;
  cudaError_t e = cudaMemsetAsync(gl, 0, elem * n * 2, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ga, 0, elem * n, st);
  return e == cudaSuccess ? MSDA_OK : (int)e;
}

// The specialised kernels carry (token * M) row offsets in a signed 28-bit field.
bool d32_eligible(int S, int M, int D, int L, int P) {
  return D == 32 && L >= 1 && L <= kMaxLevelsD32 && L * P >= 1 && L * P <= kMaxLPD32 &&
         (long long)S * M < kMaxTokenHeadsD32;
}

// CTAs per launch: enough for every warp pass, capped at `waves` x what the device holds at once
// (resident CTAs per SM = the kernel's __launch_bounds__ minimum).  waves = 1 is a persistent grid
// whose warps each run many passes (what the forward's software pipeline needs); a large cap gives
// short-lived CTAs and lets the hardware scheduler balance the tail (better for the backward).
int d32_grid(long long n_items, int resident_per_sm, int waves) {
  const long long n_pass = (n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  long long ctas = (n_pass + kD32Warps - 1) / kD32Warps;
  const long long cap = (long long)sm_count() * resident_per_sm * waves;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}

size_t d32_smem(int LP) { return (size_t)kD32Warps * kItemsPerWarp * rec_stride(LP) * 16; }

int generic_grid(long long n_items) {
  long long ctas = (n_items + (kGenericThreads / kWarp) - 1) / (kGenericThreads / kWarp);
  const long long cap = (long long)sm_count() * 64;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}

// Small forward calls run the split kernel (one item per warp): when the 4-items-per-warp
// decomposition would leave the device with less than about one wave of warps, the call costs the
// latency of each warp's sample chain, not bandwidth (900-query DINO decoder: 24.5 -> 18.5 us).
// The backward does not gain (its chain is dominated by fire-and-forget reductions: 22.7 -> 26.6 us
// when split), so it only splits on request.  Measured in profiles/r01_tuning_matrix.md.
bool use_split(long long n_items, int resident_per_sm, bool backward) {
  const int mode = g_split_mode.load();
  if (mode == 1) return false;
  if (mode == 2) return true;
  if (backward) return false;
  return n_items / kItemsPerWarp < (long long)sm_count() * resident_per_sm * kD32Warps;
}

int split_grid(long long n_items, int resident_per_sm) {
  long long ctas = (n_items + kD32Warps - 1) / kD32Warps;
  const long long cap = (long long)sm_count() * resident_per_sm * 8;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}

constexpr size_t kSplitSmem = (size_t)kD32Warps * kMaxLPD32 * 16;

template <typename VT, bool kFuse = false>
int launch_fwd_d32(D32Params& p, cudaStream_t st, bool* split = nullptr, bool* wide = nullptr, bool* slab = nullptr) {
  begin_call();
  const int LP = p.L * p.P;
  p.n_items = (long long)p.N * p.Lq * p.M;
  const bool sp = use_split(p.n_items, MSDA_FWD_MINB, false);
  if (split) *split = sp;
  register_extents(p.value, (size_t)p.N * p.S * p.M * 32 * sizeof(VT), nullptr, 0, st);
  if (sp) {
    fwd_d32_split_kernel<VT, kFuse><<<split_grid(p.n_items, 4), kD32Threads, kSplitSmem, st>>>(p);
    return finish_launch(1);
  }
  // Small pyramid, fp32 rows: one head of one sample staged in shared memory, corner rows gathered with LDS.128
  // (msda_d32_slab.cuh).  By shape when the call has at least two CTA passes (64 queries of one head) per SM.
  if (sizeof(VT) == 4 && p.L == kSlabL && p.P == kSlabP && p.S <= kSlabMaxRows && p.n_items * kSlabLP < (1ll << 31)) {
    const int slab_mode = g_fwd_slab.load();
    const long long n_pass = (long long)p.N * p.M * ((p.Lq + kSlabPassQ - 1) / kSlabPassQ);
    if (slab_mode == 2 || (slab_mode == 0 && n_pass >= 2ll * sm_count())) {
      auto kernel = p.M == 8 ? fwd_d32_slab_kernel<8, kFuse> : fwd_d32_slab_kernel<0, kFuse>;
      static std::atomic<int> attr_done[64][2];
      int dev = 0;
      cudaGetDevice(&dev);
      const int slot = dev >= 0 && dev < 64 ? dev : 0;
      if (!attr_done[slot][p.M == 8].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)slab_smem_bytes(kSlabMaxRows));
        if (e != cudaSuccess) return (int)e;
        attr_done[slot][p.M == 8].store(1, std::memory_order_release);
      }
      const long long ctas = n_pass < sm_count() ? n_pass : sm_count();
      kernel<<<(int)ctas, kSlabThreads, slab_smem_bytes(p.S), st>>>(p);
      if (split) *split = false;
      if (slab) *slab = true;
      return finish_launch(1);
    }
  }
  // 4 lanes per item, 8 items per warp pass (msda_d32_wide.cuh).  Chosen for bf16 values (64 B rows: one 128-bit load
  // moves eight rows; C2 forward 135 -> 111 us).  For fp32 rows it is slower than 8 lanes per item (184 vs 152 us: a
  // load instruction that touches 8 cache lines costs the L1 twice the cycles of one that touches 4), so fp32 takes
  // it on request only (option "fwd_wide" = 2: tests).
  const int wide_mode = g_fwd_wide.load();
  if (p.L == 4 && p.P == 4 && (wide_mode == 2 || (wide_mode == 0 && sizeof(VT) == 2))) {
    const long long n_pass = (p.n_items + kWideItems - 1) / kWideItems;
    long long ctas = (n_pass + kD32Warps - 1) / kD32Warps;
    const long long cap = (long long)sm_count() * MSDA_FWD_WIDE_MINB * MSDA_FWD_WIDE_WAVES;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    if (split) *split = false;
    if (p.M == 8) fwd_d32_wide_kernel<VT, 8, kFuse><<<(int)ctas, kD32Threads, wide_smem_bytes(), st>>>(p);
    else fwd_d32_wide_kernel<VT, 0, kFuse><<<(int)ctas, kD32Threads, wide_smem_bytes(), st>>>(p);
    if (wide) *wide = true;
    return finish_launch(1);
  }
  const int grid = d32_grid(p.n_items, kFuse ? MSDA_FWD_FUSE_MINB : MSDA_FWD_MINB, sizeof(VT) == 4 ? MSDA_FWD_WAVES : MSDA_FWD_WAVES_BF16);
  if (p.L == 4 && p.P == 4 && p.M == 8) {   // the shipped geometry (util/settings.py:102-120)
    fwd_d32_kernel<VT, 4, 4, 8, kFuse><<<grid, kD32Threads, d32_smem(LP), st>>>(p);
  } else if (p.L == 4 && p.P == 4) {        // same sampling pattern, another head count
    fwd_d32_kernel<VT, 4, 4, 0, kFuse><<<grid, kD32Threads, d32_smem(LP), st>>>(p);
  } else {
    fwd_d32_kernel<VT, 0, 0, 0, kFuse><<<grid, kD32Threads, d32_smem(LP), st>>>(p);
  }
  return finish_launch(1);
}

template <typename VT, typename GT, bool kFuse = false>
int launch_bwd_d32(D32Params& p, cudaStream_t st, bool* split = nullptr) {
  begin_call();
  const int LP = p.L * p.P;
  p.n_items = (long long)p.N * p.Lq * p.M;
  const bool sp = !kFuse && use_split(p.n_items, MSDA_BWD_MINB, true);
  if (split) *split = sp;
  register_extents(p.value, (size_t)p.N * p.S * p.M * 32 * sizeof(VT), p.grad_value,
                   (size_t)p.N * p.S * p.M * 32 * sizeof(GT), st);
  if (sp) {
    bwd_d32_split_kernel<VT, GT><<<split_grid(p.n_items, 3), kD32Threads, kSplitSmem, st>>>(p);
    return finish_launch(1);
  }
  const int grid = d32_grid(p.n_items, MSDA_BWD_MINB, MSDA_BWD_WAVES);
  if (p.L == 4 && p.P == 4 && p.M == 8) {
    bwd_d32_kernel<VT, GT, 4, 4, 8, kFuse><<<grid, kD32Threads, d32_smem(LP), st>>>(p);
  } else if (p.L == 4 && p.P == 4) {
    bwd_d32_kernel<VT, GT, 4, 4, 0, kFuse><<<grid, kD32Threads, d32_smem(LP), st>>>(p);
  } else {
    bwd_d32_kernel<VT, GT, 0, 0, 0, kFuse><<<grid, kD32Threads, d32_smem(LP), st>>>(p);
  }
  return finish_launch(1);
}

// On-chip-combining backward (msda_d32_tiles.cuh).  Eligible: L = P = 4.  Chosen by shape when the unit windows can
// cover what the queries touch: encoder self-attention (Lq == S: query i is token i, tiles are 8x8 pixel blocks) or a
// pyramid small enough to be windowed whole (S <= 2048 tokens: ARCTIC's 1045, any query order), and the call is
// large enough to fill the device with units of 64 queries x 1 head.
bool tiles_eligible(const D32Params& p) {
  // 32-bit unit and row-chunk indices inside the kernel
  const long long units_ub = (long long)p.N * p.M * ((long long)p.Lq / kTileQ + (long long)p.S / 8 + 16);
  return p.L == kTileL && p.P == 4 && (long long)p.N * p.S * p.M * 8 < (1ll << 31) && units_ub < (1ll << 31) &&
         (long long)p.N * p.Lq * p.M * kTileLP < (1ll << 31);
}

bool use_tiles(const D32Params& p) {
  const int mode = g_bwd_tiles.load();
  if (mode == 1 || !tiles_eligible(p)) return false;
  if (mode == 2) return true;
  const long long units = (long long)p.N * p.M * ((p.Lq + kTileQ - 1) / kTileQ);
  return (p.Lq == p.S || p.S <= kWholeMaxRows) && units >= 2ll * sm_count();
}

template <typename VT, typename GT, bool kFuse = false>
int launch_bwd_tiles(D32Params& p, cudaStream_t st) {
  begin_call();
  // Tiles: 8x8 pixel blocks for encoder self-attention on a large pyramid (the unit's windows must follow its queries);
  // 64 consecutive queries otherwise -- also for self-attention on a pyramid small enough to be windowed whole, where
  // the block shape buys nothing and the partly empty blocks at the level borders cost units (ARCTIC: 22 blocks of 64
  // slots for 1045 queries against 17 linear tiles; C3 encoder backward 281 -> 252 us).
#ifndef MSDA_TILE_LINEAR_WHOLE
#define MSDA_TILE_LINEAR_WHOLE 1
#endif
  const int grid_mode = (p.Lq == p.S && !(MSDA_TILE_LINEAR_WHOLE && p.S <= kWholeMaxRows)) ? 1 : 0;
  auto kernel = bwd_d32_tile_kernel<VT, GT, kFuse>;
#if MSDA_TILE_EXPERIMENT == 4 || defined(MSDA_TILE_ONE_PER_SM)   // timing experiment: one unit per SM
  const size_t smem = tile_smem_bytes() + 100 * 1024;
#else
  const size_t smem = tile_smem_bytes();
#endif
  static std::atomic<int> attr_done[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  if (!attr_done[slot].load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done[slot].store(1, std::memory_order_release);
  }
  // units = N * tiles * M; the tile count of the pixel-block enumeration depends on the level shapes, which live on
  // the device, so the grid is a fixed number of CTAs that stride over the units (two are resident per SM)
  long long ctas = (long long)p.N * p.M * ((p.Lq + kTileQ - 1) / kTileQ + (grid_mode ? 4 * kTileL : 0));
  const long long cap = (long long)sm_count() * 2 * MSDA_TILE_WAVES;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  register_extents(p.value, (size_t)p.N * p.S * p.M * 32 * sizeof(VT), p.grad_value,
                   (size_t)p.N * p.S * p.M * 32 * sizeof(GT), st);
  kernel<<<(int)ctas, kTileThreads, smem, st>>>(p, grid_mode);
  return finish_launch(1);
}

template <typename T>
int fwd_any(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc, const T* aw,
            int N, int S, int M, int D, int L, int Lq, int P, T* out, cudaStream_t st) {
  begin_call();
  const Dims d{N, S, M, D, L, Lq, P};
  const long long n_items = (long long)N * Lq * M;
  register_extents(value, (size_t)N * S * M * D * sizeof(T), nullptr, 0, st);
  fwd_generic_kernel<T><<<generic_grid(n_items), kGenericThreads, 0, st>>>(value, shapes, lsi, loc, aw, d, out);
  return finish_launch(1);
}

template <typename T>
int bwd_any(const T* go, const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc,
            const T* aw, int N, int S, int M, int D, int L, int Lq, int P, T* gv, T* gl, T* ga,
            cudaStream_t st) {
  begin_call();
  const Dims d{N, S, M, D, L, Lq, P};
  const long long n_items = (long long)N * Lq * M;
  register_extents(value, (size_t)N * S * M * D * sizeof(T), gv, (size_t)N * S * M * D * sizeof(T), st);
  bwd_generic_kernel<T><<<generic_grid(n_items), kGenericThreads, 0, st>>>(go, value, shapes, lsi, loc, aw, d, gv, gl, ga);
  return finish_launch(1);
}

}  // namespace

extern "C" {

int msda_b200_abi_version(void) { return MSDA_B200_ABI_VERSION; }

const char* msda_b200_strerror(int rc) {
  switch (rc) {
    case MSDA_OK: return "ok";
    case MSDA_ERR_NULL_POINTER: return "msda: required pointer is NULL";
    case MSDA_ERR_BAD_DIMS: return "msda: negative dimension or size overflow";
    case MSDA_ERR_IM2COL_STEP: return "msda: batch must be divisible by min(batch, im2col_step)";
    case MSDA_ERR_UNSUPPORTED: return "msda: unsupported combination (bf16 value needs channels == 32, levels <= 8, levels*points <= 32)";
    case MSDA_ERR_MISALIGNED: return "msda: pointer not aligned to its element size";
    case MSDA_ERR_NO_DEVICE: return "msda: no usable CUDA device";
    case MSDA_ERR_BAD_OPTION: return "msda: unknown option key or value";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
  return "msda: unknown error";
}

int msda_b200_check_im2col_step(int batch, int im2col_step) {
  const int step = batch < im2col_step ? batch : im2col_step;
  if (batch == 0) return MSDA_OK;
  if (step <= 0 || batch % step != 0) return MSDA_ERR_IM2COL_STEP;
  return MSDA_OK;
}

int msda_b200_set_option(const char* key, int value) {
  if (!key) return MSDA_ERR_BAD_OPTION;
  if (!strcmp(key, "fwd_path")) {
    if (value < 0 || value > 2) return MSDA_ERR_BAD_OPTION;
    g_fwd_path.store(value);
    return MSDA_OK;
  }
  if (!strcmp(key, "bwd_path")) {
    if (value < 0 || value > 2) return MSDA_ERR_BAD_OPTION;
    g_bwd_path.store(value);
    return MSDA_OK;
  }
  if (!strcmp(key, "bf16_direct_red")) {
    if (value < 0 || value > 1) return MSDA_ERR_BAD_OPTION;
    g_bf16_direct.store(value);
    return MSDA_OK;
  }
  if (!strcmp(key, "split_mode")) {
    if (value < 0 || value > 2) return MSDA_ERR_BAD_OPTION;
    g_split_mode.store(value);
    return MSDA_OK;
  }
  if (!strcmp(key, "bwd_tiles")) {
    if (value < 0 || value > 2) return MSDA_ERR_BAD_OPTION;
    g_bwd_tiles.store(value);
    return MSDA_OK;
  }
  if (!strcmp(key, "fwd_wide")) {
    if (value < 0 || value > 2) return MSDA_ERR_BAD_OPTION;
    g_fwd_wide.store(value);
    return MSDA_OK;
  }
  if (!strcmp(key, "fwd_slab")) {
    if (value < 0 || value > 2) return MSDA_ERR_BAD_OPTION;
    g_fwd_slab.store(value);
    return MSDA_OK;
  }
  return MSDA_ERR_BAD_OPTION;
}

int msda_b200_get_option(const char* key, int* value) {
  if (!key || !value) return MSDA_ERR_BAD_OPTION;
  if (!strcmp(key, "fwd_path")) { *value = g_fwd_path.load(); return MSDA_OK; }
  if (!strcmp(key, "bwd_path")) { *value = g_bwd_path.load(); return MSDA_OK; }
  if (!strcmp(key, "bf16_direct_red")) { *value = g_bf16_direct.load(); return MSDA_OK; }
  if (!strcmp(key, "split_mode")) { *value = g_split_mode.load(); return MSDA_OK; }
  if (!strcmp(key, "bwd_tiles")) { *value = g_bwd_tiles.load(); return MSDA_OK; }
  if (!strcmp(key, "fwd_wide")) { *value = g_fwd_wide.load(); return MSDA_OK; }
  if (!strcmp(key, "fwd_slab")) { *value = g_fwd_slab.load(); return MSDA_OK; }
  return MSDA_ERR_BAD_OPTION;
}

int64_t msda_b200_launch_count(void) { return (int64_t)g_launches.load(); }

const char* msda_b200_last_path(int backward) { return g_last_path[backward ? 1 : 0].load(); }

// ---- forward ------------------------------------------------------------------------------

int msda_fwd_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                 const float* aw, int N, int S, int M, int D, int L, int Lq, int P, float* out,
                 void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq)) return MSDA_OK;
  if (!out) return MSDA_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  if (L == 0 || P == 0 || S == 0) {  // no samples: the sum is empty
    const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * Lq * M * D, st);
    return e == cudaSuccess ? MSDA_OK : (int)e;
  }
  if (!value || !shapes || !lsi || !loc || !aw) return MSDA_ERR_NULL_POINTER;
  if (!aligned(value, 4) || !aligned(loc, 4) || !aligned(aw, 4) || !aligned(out, 4) ||
      !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  const int path = g_fwd_path.load();
  const bool fast = path != 1 && d32_eligible(S, M, D, L, P) && aligned(value, 16) &&
                    aligned(out, 16) && aligned(loc, 8);
  if (fast) {
    D32Params p{};
    p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = loc; p.aw = aw; p.out = out;
    p.N = N; p.S = S; p.M = M; p.L = L; p.Lq = Lq; p.P = P;
    bool sp = false, wide = false, slab = false;
    rc = launch_fwd_d32<float>(p, st, &sp, &wide, &slab);
    g_last_path[0] = sp ? "fwd_d32_split_f32" : slab ? "fwd_d32_slab_f32" : wide ? "fwd_d32_wide_f32" : "fwd_d32_f32";
    return rc;
  }
  g_last_path[0] = "fwd_generic_f32";
  return fwd_any<float>(value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, out, st);
}

int msda_fwd_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                 const double* aw, int N, int S, int M, int D, int L, int Lq, int P, double* out,
                 void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq)) return MSDA_OK;
  if (!out) return MSDA_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  if (L == 0 || P == 0 || S == 0) {
    const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * (size_t)N * Lq * M * D, st);
    return e == cudaSuccess ? MSDA_OK : (int)e;
  }
  if (!value || !shapes || !lsi || !loc || !aw) return MSDA_ERR_NULL_POINTER;
  if (!aligned(value, 8) || !aligned(loc, 8) || !aligned(aw, 8) || !aligned(out, 8) ||
      !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  g_last_path[0] = "fwd_generic_f64";
  return fwd_any<double>(value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, out, st);
}

int msda_fwd_bf16(const uint16_t* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                  const float* aw, int N, int S, int M, int D, int L, int Lq, int P, float* out,
                  void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq)) return MSDA_OK;
  if (!out) return MSDA_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  if (L == 0 || P == 0 || S == 0) {
    const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * Lq * M * D, st);
    return e == cudaSuccess ? MSDA_OK : (int)e;
  }
  if (!value || !shapes || !lsi || !loc || !aw) return MSDA_ERR_NULL_POINTER;
  if (!d32_eligible(S, M, D, L, P)) return MSDA_ERR_UNSUPPORTED;
  if (!aligned(value, 8) || !aligned(out, 16) || !aligned(loc, 8) || !aligned(aw, 4) ||
      !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  D32Params p{};
  p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = loc; p.aw = aw; p.out = out;
  p.N = N; p.S = S; p.M = M; p.L = L; p.Lq = Lq; p.P = P;
  bool sp = false, wide = false;
  rc = launch_fwd_d32<__nv_bfloat16>(p, st, &sp, &wide);
  g_last_path[0] = sp ? "fwd_d32_split_bf16" : wide ? "fwd_d32_wide_bf16" : "fwd_d32_bf16";
  return rc;
}

// ---- backward -----------------------------------------------------------------------------

int msda_bwd_f32(const float* go, const float* value, const int64_t* shapes, const int64_t* lsi,
                 const float* loc, const float* aw, int N, int S, int M, int D, int L, int Lq, int P,
                 float* gv, float* gl, float* ga, void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq) || L == 0 || P == 0) return MSDA_OK;
  if (S == 0) return zero_sample_grads(gl, ga, sizeof(float), N, Lq, M, L, P, (cudaStream_t)stream);
  if (!go || !value || !shapes || !lsi || !loc || !aw || !gv || !gl || !ga) return MSDA_ERR_NULL_POINTER;
  if (!aligned(go, 4) || !aligned(value, 4) || !aligned(loc, 4) || !aligned(aw, 4) || !aligned(gv, 4) ||
      !aligned(gl, 4) || !aligned(ga, 4) || !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  const int path = g_bwd_path.load();
  const bool fast = path != 1 && d32_eligible(S, M, D, L, P) && aligned(value, 16) && aligned(gv, 16) &&
                    aligned(go, 16) && aligned(loc, 8) && aligned(gl, 8);
  if (fast) {
    D32Params p{};
    p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = loc; p.aw = aw; p.grad_out = go;
    p.grad_value = gv; p.grad_loc = gl; p.grad_aw = ga;
    p.N = N; p.S = S; p.M = M; p.L = L; p.Lq = Lq; p.P = P;
    if (use_tiles(p)) {
      g_last_path[1] = "bwd_d32_tiles_f32";
      return launch_bwd_tiles<float, float>(p, st);
    }
    bool sp = false;
    rc = launch_bwd_d32<float, float>(p, st, &sp);
    g_last_path[1] = sp ? "bwd_d32_split_f32" : "bwd_d32_f32";
    return rc;
  }
  g_last_path[1] = "bwd_generic_f32";
  return bwd_any<float>(go, value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, gv, gl, ga, st);
}

int msda_bwd_f64(const double* go, const double* value, const int64_t* shapes, const int64_t* lsi,
                 const double* loc, const double* aw, int N, int S, int M, int D, int L, int Lq, int P,
                 double* gv, double* gl, double* ga, void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq) || L == 0 || P == 0) return MSDA_OK;
  if (S == 0) return zero_sample_grads(gl, ga, sizeof(double), N, Lq, M, L, P, (cudaStream_t)stream);
  if (!go || !value || !shapes || !lsi || !loc || !aw || !gv || !gl || !ga) return MSDA_ERR_NULL_POINTER;
  if (!aligned(go, 8) || !aligned(value, 8) || !aligned(loc, 8) || !aligned(aw, 8) || !aligned(gv, 8) ||
      !aligned(gl, 8) || !aligned(ga, 8) || !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  g_last_path[1] = "bwd_generic_f64";
  return bwd_any<double>(go, value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, gv, gl, ga,
                         (cudaStream_t)stream);
}

// fp32 scratch -> bf16 grad_value (one rounding), `n4` groups of four elements
static int round_scratch_to_bf16(const float* scratch, uint16_t* gv, long long n4, cudaStream_t st) {
  if (n4 <= 0) return MSDA_OK;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  begin_call();
  f32_to_bf16_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(scratch), reinterpret_cast<uint2*>(gv), n4);
  return finish_launch(1);
}

int msda_bwd_bf16(const float* go, const uint16_t* value, const int64_t* shapes, const int64_t* lsi,
                  const float* loc, const float* aw, int N, int S, int M, int D, int L, int Lq, int P,
                  uint16_t* gv, float* gl, float* ga, float* gv_scratch, void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq) || L == 0 || P == 0) return MSDA_OK;
  if (S == 0) return zero_sample_grads(gl, ga, sizeof(float), N, Lq, M, L, P, (cudaStream_t)stream);
  if (!go || !value || !shapes || !lsi || !loc || !aw || !gv || !gl || !ga) return MSDA_ERR_NULL_POINTER;
  if (!d32_eligible(S, M, D, L, P)) return MSDA_ERR_UNSUPPORTED;
  if (!aligned(go, 16) || !aligned(value, 8) || !aligned(gv, 8) || !aligned(loc, 8) || !aligned(aw, 4) ||
      !aligned(gl, 8) || !aligned(ga, 4) || !aligned(shapes, 8) || !aligned(lsi, 8) ||
      (gv_scratch && !aligned(gv_scratch, 16)))
    return MSDA_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  D32Params p{};
  p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = loc; p.aw = aw; p.grad_out = go;
  p.grad_loc = gl; p.grad_aw = ga;
  p.N = N; p.S = S; p.M = M; p.L = L; p.Lq = Lq; p.P = P;
  if (gv_scratch) {
    // accurate mode: fp32 reductions into the scratch, one rounding to bf16 at the end
    p.grad_value = gv_scratch;
    if (use_tiles(p)) {
      g_last_path[1] = "bwd_d32_tiles_bf16_f32acc";
      rc = launch_bwd_tiles<__nv_bfloat16, float>(p, st);
    } else {
      g_last_path[1] = "bwd_d32_bf16_f32acc";
      rc = launch_bwd_d32<__nv_bfloat16, float>(p, st);
    }
    if (rc) return rc;
    if (gv == reinterpret_cast<uint16_t*>(gv_scratch)) return MSDA_OK;   // (host pipeline: rounding is done per sample later)
    return round_scratch_to_bf16(gv_scratch, gv, (long long)N * S * M * 8, st);
  }
  p.grad_value = gv;
  if (use_tiles(p)) {
    g_last_path[1] = "bwd_d32_tiles_bf16";
    return launch_bwd_tiles<__nv_bfloat16, __nv_bfloat16>(p, st);
  }
  g_last_path[1] = "bwd_d32_bf16";
  return launch_bwd_d32<__nv_bfloat16, __nv_bfloat16>(p, st);
}

// ---- fused prologue -----------------------------------------------------

int msda_fused_fwd_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* ref,
                       const float* offsets, const float* logits, int N, int S, int M, int D, int L,
                       int Lq, int P, float* out, void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq)) return MSDA_OK;
  if (!out) return MSDA_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  if (L == 0 || P == 0 || S == 0) {
    const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * Lq * M * D, st);
    return e == cudaSuccess ? MSDA_OK : (int)e;
  }
  if (!value || !shapes || !lsi || !ref || !offsets || !logits) return MSDA_ERR_NULL_POINTER;
  if (!d32_eligible(S, M, D, L, P)) return MSDA_ERR_UNSUPPORTED;
  if (!aligned(value, 16) || !aligned(out, 16) || !aligned(offsets, 8) || !aligned(ref, 8) || !aligned(logits, 4) ||
      !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  D32Params p{};
  p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = offsets; p.aw = logits; p.ref = ref; p.out = out;
  p.N = N; p.S = S; p.M = M; p.L = L; p.Lq = Lq; p.P = P;
  bool sp = false, wide = false, slab = false;
  rc = launch_fwd_d32<float, true>(p, st, &sp, &wide, &slab);
  g_last_path[0] = sp ? "fwd_d32_split_fused_f32" : slab ? "fwd_d32_slab_fused_f32" : wide ? "fwd_d32_wide_fused_f32" : "fwd_d32_fused_f32";
  return rc;
}

int msda_fused_bwd_f32(const float* go, const float* value, const int64_t* shapes, const int64_t* lsi,
                       const float* ref, const float* offsets, const float* logits, int N, int S, int M,
                       int D, int L, int Lq, int P, float* gv, float* g_offsets, float* g_logits,
                       float* g_ref, void* stream) {
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  if (nothing_to_do(N, M, D, Lq) || L == 0 || P == 0) return MSDA_OK;
  if (S == 0) return zero_sample_grads(g_offsets, g_logits, sizeof(float), N, Lq, M, L, P, (cudaStream_t)stream);
  if (!go || !value || !shapes || !lsi || !ref || !offsets || !logits || !gv || !g_offsets || !g_logits)
    return MSDA_ERR_NULL_POINTER;
  if (!d32_eligible(S, M, D, L, P)) return MSDA_ERR_UNSUPPORTED;
  if (!aligned(go, 16) || !aligned(value, 16) || !aligned(gv, 16) || !aligned(offsets, 8) || !aligned(ref, 8) ||
      !aligned(logits, 4) || !aligned(g_offsets, 8) || !aligned(g_logits, 4) || (g_ref && !aligned(g_ref, 4)) ||
      !aligned(shapes, 8) || !aligned(lsi, 8))
    return MSDA_ERR_MISALIGNED;
  D32Params p{};
  p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = offsets; p.aw = logits; p.ref = ref; p.grad_out = go;
  p.grad_value = gv; p.grad_loc = g_offsets; p.grad_aw = g_logits; p.grad_ref = g_ref;
  p.N = N; p.S = S; p.M = M; p.L = L; p.Lq = Lq; p.P = P;
  if (use_tiles(p)) {
    g_last_path[1] = "bwd_d32_tiles_fused_f32";
    return launch_bwd_tiles<float, float, true>(p, (cudaStream_t)stream);
  }
  g_last_path[1] = "bwd_d32_fused_f32";
  return launch_bwd_d32<float, float, true>(p, (cudaStream_t)stream);
}

// ---- value producer side ------------------------------------------------

static int value_pack_common(const void* src, const uint8_t* mask, void* dst, int64_t rows, int cols, bool pack,
                             void* stream) {
  if (rows < 0 || cols < 0) return MSDA_ERR_BAD_DIMS;
  if (rows == 0 || cols == 0) return MSDA_OK;
  if (!src || !dst) return MSDA_ERR_NULL_POINTER;
  if (cols % 8 != 0) return MSDA_ERR_UNSUPPORTED;
  if (!aligned(src, 16) || !aligned(dst, 16)) return MSDA_ERR_MISALIGNED;
  const long long n8 = (long long)rows * (cols / 8);
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  begin_call();
  if (pack)
    value_pack_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src), mask,
                                                                         reinterpret_cast<uint4*>(dst), n8, cols / 8);
  else
    value_unpack_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), mask,
                                                                           reinterpret_cast<float4*>(dst), n8, cols / 8);
  return finish_launch(1);
}

int msda_value_pack_bf16(const float* value_f32, const uint8_t* padding_mask, int64_t rows, int cols,
                         uint16_t* value_bf16, void* stream) {
  return value_pack_common(value_f32, padding_mask, value_bf16, rows, cols, true, stream);
}

int msda_value_unpack_grad_bf16(const uint16_t* grad_bf16, const uint8_t* padding_mask, int64_t rows, int cols,
                                float* grad_f32, void* stream) {
  return value_pack_common(grad_bf16, padding_mask, grad_f32, rows, cols, false, stream);
}

// ---- parity probe -------------------------------------------------------------------------

int msda_sample_index_f32(const int64_t* shapes, const float* loc, int N, int Lq, int M, int L, int P,
                          int32_t* h_low, int32_t* w_low, uint8_t* mask, void* stream) {
  if (N < 0 || Lq < 0 || M < 0 || L < 0 || P < 0) return MSDA_ERR_BAD_DIMS;
  const long long n = (long long)N * Lq * M * L * P;
  if (n == 0) return MSDA_OK;
  if (!shapes || !loc || !h_low || !w_low || !mask) return MSDA_ERR_NULL_POINTER;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  begin_call();
  sample_index_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(shapes, loc, n, L, P, h_low, w_low, mask);
  return finish_launch(1);
}

// ---- host-buffer entry points -------------------------------------------------------------

struct msda_host_plan {
  int device;
  int N, S, M, D, L, Lq, P;
  size_t n_value, n_loc, n_aw, n_out;
  float *d_value, *d_loc, *d_aw, *d_go, *d_out, *d_gv, *d_gl, *d_ga;
  int value_bf16;               // value and grad_value travel as bf16 (d_value holds bf16, d_gv is the fp32 accumulation
  uint16_t* d_gv16;             // buffer, d_gv16 the rounded grad_value that is downloaded)
  cudaEvent_t* ev_gv;           // [N] bf16 mode: grad_value[b] rounded
  int64_t *d_shapes, *d_lsi;
  // Three streams so that upload, kernels and download overlap (PCIe is full duplex and the copy
  // engines run beside the SMs): a call is cut into per-sample query chunks that flow through them.
  cudaStream_t s_in, s_run, s_out;
  int n_chunks;                 // query chunks per sample
  cudaEvent_t* ev_in;           // [N * n_chunks] chunk inputs on the device
  cudaEvent_t* ev_run;          // [N * n_chunks] chunk results computed
  cudaEvent_t* ev_value;        // [N] value[b] on the device
  cudaEvent_t ev_tables;
  cudaEvent_t* ev_out;          // [N * n_chunks + N] trace only: chunk results / grad_value[b] on the host
  int64_t h_tables[3 * 64];     // spatial_shapes + level_start_index as last uploaded: a step with the same level tables
  int tables_valid;             // (every step of a training run) skips the two small uploads
  int whole_batch;              // a pipe slot with MSDA_B200_PIPE_CHUNKS=0 (default): the step is ONE packet -- four
                                // uploads, one forward + backward over all samples, four downloads (steps overlap each
                                // other, so nothing inside a step needs to; large copies run at the best two-way rate)
  int owns_streams;             // 0: a slot of a msda_host_pipe, its streams belong to the pipe's first slot
  cudaEvent_t ev_done;          // asynchronous use (msda_host_pipe): everything of the last submitted step has landed
  int ramp;                     // MSDA_B200_HOST_RAMP=1: growing chunks in the first sample (measured: no gain)
  int trace;                    // MSDA_B200_HOST_TRACE=1: print the timeline of every call to stderr
};

#define MSDA_CUDA_TRY(expr)                     \
  do {                                          \
    const cudaError_t e__ = (expr);             \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

static int plan_create(msda_host_plan** out_plan, int device, int N, int S, int M, int D, int L, int Lq, int P,
                       int value_bf16, const msda_host_plan* share = nullptr, int chunks = 0);

int msda_host_plan_create(msda_host_plan** out_plan, int device, int N, int S, int M, int D, int L,
                          int Lq, int P) {
  return plan_create(out_plan, device, N, S, M, D, L, Lq, P, 0);
}

int msda_host_plan_create_bf16(msda_host_plan** out_plan, int device, int N, int S, int M, int D, int L,
                               int Lq, int P) {
  if (!d32_eligible(S > 0 ? S : 1, M > 0 ? M : 1, D, L > 0 ? L : 1, P > 0 ? P : 1)) return MSDA_ERR_UNSUPPORTED;
  return plan_create(out_plan, device, N, S, M, D, L, Lq, P, 1);
}

static int plan_create(msda_host_plan** out_plan, int device, int N, int S, int M, int D, int L, int Lq, int P,
                       int value_bf16, const msda_host_plan* share, int chunks) {
  if (!out_plan) return MSDA_ERR_NULL_POINTER;
  *out_plan = nullptr;
  int rc = check_dims(N, S, M, D, L, Lq, P);
  if (rc) return rc;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
    cudaGetLastError();
    return MSDA_ERR_NO_DEVICE;
  }
  MSDA_CUDA_TRY(cudaSetDevice(device));
  cudaGetLastError();  // a stale, non-sticky error of an earlier caller must not fail this call
  msda_host_plan* pl = new (std::nothrow) msda_host_plan();
  if (!pl) return (int)cudaErrorMemoryAllocation;
  memset(pl, 0, sizeof(*pl));
  pl->device = device;
  pl->value_bf16 = value_bf16;
  pl->N = N; pl->S = S; pl->M = M; pl->D = D; pl->L = L; pl->Lq = Lq; pl->P = P;
  pl->n_value = (size_t)N * S * M * D;
  pl->n_aw = (size_t)N * Lq * M * L * P;
  pl->n_loc = pl->n_aw * 2;
  pl->n_out = (size_t)N * Lq * M * D;
  // chunks of about 28 MB of per-query traffic (in + out), at most 16 per sample: measured at the
  // DETR-encoder shape 1/2/4/8/16/32 chunks per sample -> 5.10/4.79/4.72/4.86/5.16/5.90 ms per call
  // (each chunk costs ~25 us of copy / event / launch overhead; too few leave the download idle)
  {
    const double per_query = 4.0 * M * (2.0 * D + 6.0 * L * P);  // go + out + (loc, aw, gl, ga)
    long long c = (long long)(per_query * Lq / (28.0 * 1024 * 1024) + 0.5);
    if (c < 1) c = 1;
    if (c > 16) c = 16;
    if (const char* env = getenv("MSDA_B200_HOST_CHUNKS")) {   // tuning override
      const long long v = atoll(env);
      if (v >= 1 && v <= 256) c = v;
    }
    if (chunks > 0) c = chunks;                                // a pipe's slots: steps overlap each other instead
    if (c > Lq) c = Lq > 0 ? Lq : 1;
    pl->n_chunks = (int)c;
  }
  cudaError_t e = cudaSuccess;
  pl->owns_streams = share == nullptr;
  if (share) {
    pl->s_in = share->s_in; pl->s_run = share->s_run; pl->s_out = share->s_out;
  } else {
    e = cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->s_run, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_done, cudaEventDisableTiming);
  const int n_ev = (N > 0 ? N : 1) * pl->n_chunks;
  pl->trace = getenv("MSDA_B200_HOST_TRACE") != nullptr;
  pl->ramp = 0;
  if (const char* env = getenv("MSDA_B200_HOST_RAMP")) pl->ramp = atoi(env) != 0;
  const unsigned ev_flags = pl->trace ? cudaEventDefault : cudaEventDisableTiming;
  if (pl->trace) {
    pl->ev_out = new (std::nothrow) cudaEvent_t[n_ev + (N > 0 ? N : 1)]();
    for (int i = 0; pl->ev_out && i < n_ev + N; ++i) cudaEventCreate(&pl->ev_out[i]);
  }
  pl->ev_in = new (std::nothrow) cudaEvent_t[n_ev]();
  pl->ev_run = new (std::nothrow) cudaEvent_t[n_ev]();
  pl->ev_value = new (std::nothrow) cudaEvent_t[N > 0 ? N : 1]();
  pl->ev_gv = new (std::nothrow) cudaEvent_t[N > 0 ? N : 1]();
  if (!pl->ev_in || !pl->ev_run || !pl->ev_value || !pl->ev_gv) e = cudaErrorMemoryAllocation;
  for (int i = 0; i < n_ev && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&pl->ev_in[i], ev_flags);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_run[i], ev_flags);
  }
  for (int i = 0; i < N && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&pl->ev_value[i], ev_flags);
  for (int i = 0; i < N && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&pl->ev_gv[i], ev_flags);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_tables, ev_flags);
  auto alloc = [&](float** p, size_t n) {
    if (e == cudaSuccess) e = cudaMalloc((void**)p, sizeof(float) * (n ? n : 1));
  };
  alloc(&pl->d_value, pl->n_value);
  alloc(&pl->d_loc, pl->n_loc);
  alloc(&pl->d_aw, pl->n_aw);
  alloc(&pl->d_go, pl->n_out);
  alloc(&pl->d_out, pl->n_out);
  alloc(&pl->d_gv, pl->n_value);
  if (value_bf16 && e == cudaSuccess) e = cudaMalloc((void**)&pl->d_gv16, sizeof(uint16_t) * (pl->n_value ? pl->n_value : 1));
  alloc(&pl->d_gl, pl->n_loc);
  alloc(&pl->d_ga, pl->n_aw);
  if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_shapes, sizeof(int64_t) * 2 * (L ? L : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_lsi, sizeof(int64_t) * (L ? L : 1));
  if (e != cudaSuccess) {
    msda_host_plan_destroy(pl);
    return (int)e;
  }
  *out_plan = pl;
  return MSDA_OK;
}

void msda_host_plan_destroy(msda_host_plan* pl) {
  if (!pl) return;
  cudaSetDevice(pl->device);
  cudaFree(pl->d_value); cudaFree(pl->d_loc); cudaFree(pl->d_aw); cudaFree(pl->d_go);
  cudaFree(pl->d_out); cudaFree(pl->d_gv); cudaFree(pl->d_gl); cudaFree(pl->d_ga);
  cudaFree(pl->d_gv16);
  cudaFree(pl->d_shapes); cudaFree(pl->d_lsi);
  const int n_ev = (pl->N > 0 ? pl->N : 1) * pl->n_chunks;
  for (int i = 0; i < n_ev; ++i) {
    if (pl->ev_in && pl->ev_in[i]) cudaEventDestroy(pl->ev_in[i]);
    if (pl->ev_run && pl->ev_run[i]) cudaEventDestroy(pl->ev_run[i]);
  }
  for (int i = 0; i < pl->N; ++i) {
    if (pl->ev_value && pl->ev_value[i]) cudaEventDestroy(pl->ev_value[i]);
    if (pl->ev_gv && pl->ev_gv[i]) cudaEventDestroy(pl->ev_gv[i]);
  }
  delete[] pl->ev_gv;
  if (pl->ev_out) {
    for (int i = 0; i < n_ev + pl->N; ++i)
      if (pl->ev_out[i]) cudaEventDestroy(pl->ev_out[i]);
    delete[] pl->ev_out;
  }
  delete[] pl->ev_in; delete[] pl->ev_run; delete[] pl->ev_value;
  if (pl->ev_tables) cudaEventDestroy(pl->ev_tables);
  if (pl->ev_done) cudaEventDestroy(pl->ev_done);
  if (pl->owns_streams) {
    if (pl->s_in) cudaStreamDestroy(pl->s_in);
    if (pl->s_run) cudaStreamDestroy(pl->s_run);
    if (pl->s_out) cudaStreamDestroy(pl->s_out);
  }
  delete pl;
}

int msda_host_plan_traffic(const msda_host_plan* pl, int64_t* h2d, int64_t* d2h) {
  if (!pl) return MSDA_ERR_NULL_POINTER;
  const int64_t vb = pl->value_bf16 ? 2 : 4;
  const int64_t in = vb * (int64_t)pl->n_value + (int64_t)sizeof(float) * (pl->n_loc + pl->n_aw + pl->n_out) +
                     (int64_t)sizeof(int64_t) * 3 * pl->L;
  const int64_t out = vb * (int64_t)pl->n_value + (int64_t)sizeof(float) * (pl->n_out + pl->n_loc + pl->n_aw);
  if (h2d) *h2d = in;
  if (d2h) *d2h = out;
  return MSDA_OK;
}

int msda_host_register(void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return MSDA_ERR_NULL_POINTER;
  MSDA_CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return MSDA_OK;
}

int msda_host_unregister(void* ptr) {
  if (!ptr) return MSDA_ERR_NULL_POINTER;
  MSDA_CUDA_TRY(cudaHostUnregister(ptr));
  return MSDA_OK;
}

// The (up to three) copies of one chunk, issued back to back on one stream.
static cudaError_t copy_batch(void** dsts, void** srcs, size_t* sizes, int n, cudaMemcpyKind kind, cudaStream_t st) {
  for (int i = 0; i < n; ++i) {
    const cudaError_t e = cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], kind, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// One call = N * n_chunks work packets (sample b, queries [q0, q1)):
//   s_in : tables, then per sample: value[b], then per chunk: loc, aw (, grad_out)
//   s_run: memset(grad_value); per chunk: forward (+ backward) on that query range of sample b
//   s_out: per chunk: out (, grad_loc, grad_aw); after the last chunk of a sample: grad_value[b]
// The kernels see each packet as a problem with batch 1 and q1-q0 queries (the operator has no
// cross-query or cross-sample term), so nothing but pointer offsets is needed.
static int host_pipeline(msda_host_plan* pl, const void* value_any, const int64_t* shapes, const int64_t* lsi,
                         const float* loc, const float* aw, const float* go, float* out, void* gv_any,
                         float* gl, float* ga, bool blocking = true) {
  const bool b16 = pl->value_bf16 != 0;
  const size_t vsz = b16 ? 2 : 4;
  const char* value = static_cast<const char*>(value_any);
  char* gv = static_cast<char*>(gv_any);
  char* d_value = reinterpret_cast<char*>(pl->d_value);
  const bool bwd = go != nullptr;
  const int N = pl->N, S = pl->S, M = pl->M, D = pl->D, L = pl->L, Lq = pl->Lq, P = pl->P;
  if (N == 0) return MSDA_OK;
  const size_t v_per = (size_t)S * M * D;
  const size_t o_per_q = (size_t)M * D, a_per_q = (size_t)M * L * P;
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  {
    // the level tables are read from the CALLER's arrays right now (they may be pageable and change after we return)
    const bool cacheable = L <= 64;
    if (!(cacheable && pl->tables_valid && !memcmp(pl->h_tables, shapes, sizeof(int64_t) * 2 * L) &&
          !memcmp(pl->h_tables + 2 * L, lsi, sizeof(int64_t) * L))) {
      MSDA_CUDA_TRY(cudaMemcpyAsync(pl->d_shapes, shapes, sizeof(int64_t) * 2 * L, H2D, pl->s_in));
      MSDA_CUDA_TRY(cudaMemcpyAsync(pl->d_lsi, lsi, sizeof(int64_t) * L, H2D, pl->s_in));
      pl->tables_valid = 0;
      if (cacheable) {
        memcpy(pl->h_tables, shapes, sizeof(int64_t) * 2 * L);
        memcpy(pl->h_tables + 2 * L, lsi, sizeof(int64_t) * L);
        pl->tables_valid = 1;
      }
    }
  }
  MSDA_CUDA_TRY(cudaEventRecord(pl->ev_tables, pl->s_in));
  MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_run, pl->ev_tables, 0));
  if (bwd && pl->n_value) MSDA_CUDA_TRY(cudaMemsetAsync(pl->d_gv, 0, sizeof(float) * pl->n_value, pl->s_run));
  if (pl->whole_batch && !b16 && bwd) {
    if (pl->n_value) MSDA_CUDA_TRY(cudaMemcpyAsync(pl->d_value, value, sizeof(float) * pl->n_value, H2D, pl->s_in));
    if (pl->n_loc) MSDA_CUDA_TRY(cudaMemcpyAsync(pl->d_loc, loc, sizeof(float) * pl->n_loc, H2D, pl->s_in));
    if (pl->n_aw) MSDA_CUDA_TRY(cudaMemcpyAsync(pl->d_aw, aw, sizeof(float) * pl->n_aw, H2D, pl->s_in));
    if (pl->n_out) MSDA_CUDA_TRY(cudaMemcpyAsync(pl->d_go, go, sizeof(float) * pl->n_out, H2D, pl->s_in));
    MSDA_CUDA_TRY(cudaEventRecord(pl->ev_in[0], pl->s_in));
    MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_run, pl->ev_in[0], 0));
    if (Lq > 0) {
      int rc = msda_fwd_f32(pl->d_value, pl->d_shapes, pl->d_lsi, pl->d_loc, pl->d_aw, N, S, M, D, L, Lq, P, pl->d_out, pl->s_run);
      if (rc) return rc;
      rc = msda_bwd_f32(pl->d_go, pl->d_value, pl->d_shapes, pl->d_lsi, pl->d_loc, pl->d_aw, N, S, M, D, L, Lq, P, pl->d_gv,
                        pl->d_gl, pl->d_ga, pl->s_run);
      if (rc) return rc;
    }
    MSDA_CUDA_TRY(cudaEventRecord(pl->ev_run[0], pl->s_run));
    MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_out, pl->ev_run[0], 0));
    if (pl->n_out) MSDA_CUDA_TRY(cudaMemcpyAsync(out, pl->d_out, sizeof(float) * pl->n_out, D2H, pl->s_out));
    if (pl->n_loc) MSDA_CUDA_TRY(cudaMemcpyAsync(gl, pl->d_gl, sizeof(float) * pl->n_loc, D2H, pl->s_out));
    if (pl->n_aw) MSDA_CUDA_TRY(cudaMemcpyAsync(ga, pl->d_ga, sizeof(float) * pl->n_aw, D2H, pl->s_out));
    if (pl->n_value) MSDA_CUDA_TRY(cudaMemcpyAsync(gv, pl->d_gv, sizeof(float) * pl->n_value, D2H, pl->s_out));
    if (!blocking) {
      MSDA_CUDA_TRY(cudaEventRecord(pl->ev_done, pl->s_out));
      return MSDA_OK;
    }
    MSDA_CUDA_TRY(cudaStreamSynchronize(pl->s_out));
    return MSDA_OK;
  }
  for (int b = 0; b < N; ++b) {
    if (v_per) MSDA_CUDA_TRY(cudaMemcpyAsync(d_value + b * v_per * vsz, value + b * v_per * vsz, vsz * v_per, H2D, pl->s_in));
    MSDA_CUDA_TRY(cudaEventRecord(pl->ev_value[b], pl->s_in));
    MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_run, pl->ev_value[b], 0));
    for (int c = 0; c < pl->n_chunks; ++c) {
      // equal chunks, except in the first sample where they grow quadratically: a small first chunk
      // gets the download stream going early, large later chunks keep the copies efficient
      auto cut = [&](int i) {
        const long long C = pl->n_chunks;
        return (b == 0 && pl->ramp) ? (int)((long long)Lq * i * i / (C * C)) : (int)((long long)Lq * i / C);
      };
      const int q0 = cut(c), q1 = cut(c + 1);
      const int nq = q1 - q0;
      const int k = b * pl->n_chunks + c;
      const size_t qo = (size_t)b * Lq + q0;
      {
        void* dsts[3]; void* srcs[3]; size_t sizes[3]; int n = 0;
        if (nq > 0 && a_per_q) {
          dsts[n] = pl->d_loc + qo * a_per_q * 2; srcs[n] = const_cast<float*>(loc + qo * a_per_q * 2); sizes[n++] = sizeof(float) * nq * a_per_q * 2;
          dsts[n] = pl->d_aw + qo * a_per_q; srcs[n] = const_cast<float*>(aw + qo * a_per_q); sizes[n++] = sizeof(float) * nq * a_per_q;
        }
        if (bwd && nq > 0 && o_per_q) {
          dsts[n] = pl->d_go + qo * o_per_q; srcs[n] = const_cast<float*>(go + qo * o_per_q); sizes[n++] = sizeof(float) * nq * o_per_q;
        }
        if (n) MSDA_CUDA_TRY(copy_batch(dsts, srcs, sizes, n, H2D, pl->s_in));
      }
      MSDA_CUDA_TRY(cudaEventRecord(pl->ev_in[k], pl->s_in));
      MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_run, pl->ev_in[k], 0));
      if (nq > 0) {
        int rc;
        if (b16)
          rc = msda_fwd_bf16(reinterpret_cast<const uint16_t*>(d_value) + b * v_per, pl->d_shapes, pl->d_lsi,
                             pl->d_loc + qo * a_per_q * 2, pl->d_aw + qo * a_per_q, 1, S, M, D, L, nq, P,
                             pl->d_out + qo * o_per_q, pl->s_run);
        else
          rc = msda_fwd_f32(pl->d_value + b * v_per, pl->d_shapes, pl->d_lsi, pl->d_loc + qo * a_per_q * 2,
                            pl->d_aw + qo * a_per_q, 1, S, M, D, L, nq, P, pl->d_out + qo * o_per_q, pl->s_run);
        if (rc) return rc;
        if (bwd) {
          if (b16)   // fp32 accumulation into d_gv; rounded to bf16 once per sample below (gv == scratch: no rounding here)
            rc = msda_bwd_bf16(pl->d_go + qo * o_per_q, reinterpret_cast<const uint16_t*>(d_value) + b * v_per, pl->d_shapes,
                               pl->d_lsi, pl->d_loc + qo * a_per_q * 2, pl->d_aw + qo * a_per_q, 1, S, M, D, L, nq, P,
                               reinterpret_cast<uint16_t*>(pl->d_gv + b * v_per), pl->d_gl + qo * a_per_q * 2,
                               pl->d_ga + qo * a_per_q, pl->d_gv + b * v_per, pl->s_run);
          else
            rc = msda_bwd_f32(pl->d_go + qo * o_per_q, pl->d_value + b * v_per, pl->d_shapes, pl->d_lsi,
                              pl->d_loc + qo * a_per_q * 2, pl->d_aw + qo * a_per_q, 1, S, M, D, L, nq, P,
                              pl->d_gv + b * v_per, pl->d_gl + qo * a_per_q * 2, pl->d_ga + qo * a_per_q, pl->s_run);
          if (rc) return rc;
        }
      }
      MSDA_CUDA_TRY(cudaEventRecord(pl->ev_run[k], pl->s_run));
      MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_out, pl->ev_run[k], 0));
      {
        void* dsts[3]; void* srcs[3]; size_t sizes[3]; int n = 0;
        if (nq > 0 && o_per_q) {
          dsts[n] = out + qo * o_per_q; srcs[n] = pl->d_out + qo * o_per_q; sizes[n++] = sizeof(float) * nq * o_per_q;
        }
        if (bwd && nq > 0 && a_per_q) {
          dsts[n] = gl + qo * a_per_q * 2; srcs[n] = pl->d_gl + qo * a_per_q * 2; sizes[n++] = sizeof(float) * nq * a_per_q * 2;
          dsts[n] = ga + qo * a_per_q; srcs[n] = pl->d_ga + qo * a_per_q; sizes[n++] = sizeof(float) * nq * a_per_q;
        }
        if (n) MSDA_CUDA_TRY(copy_batch(dsts, srcs, sizes, n, D2H, pl->s_out));
      }
      if (pl->trace && pl->ev_out) cudaEventRecord(pl->ev_out[k], pl->s_out);
    }
    if (bwd && v_per) {
      if (b16) {
        const int rc = round_scratch_to_bf16(pl->d_gv + b * v_per, pl->d_gv16 + b * v_per, (long long)(v_per / 4), pl->s_run);
        if (rc) return rc;
        MSDA_CUDA_TRY(cudaEventRecord(pl->ev_gv[b], pl->s_run));
        MSDA_CUDA_TRY(cudaStreamWaitEvent(pl->s_out, pl->ev_gv[b], 0));
        MSDA_CUDA_TRY(cudaMemcpyAsync(gv + b * v_per * 2, pl->d_gv16 + b * v_per, 2 * v_per, D2H, pl->s_out));
      } else {
        MSDA_CUDA_TRY(cudaMemcpyAsync(gv + b * v_per * 4, pl->d_gv + b * v_per, sizeof(float) * v_per, D2H, pl->s_out));
      }
    }
    if (pl->trace && pl->ev_out) cudaEventRecord(pl->ev_out[N * pl->n_chunks + b], pl->s_out);
  }
  if (!blocking) {
    // the download stream has waited for every kernel, and every kernel for its uploads: its tail is the step's end
    MSDA_CUDA_TRY(cudaEventRecord(pl->ev_done, pl->s_out));
    return MSDA_OK;
  }
  MSDA_CUDA_TRY(cudaStreamSynchronize(pl->s_out));
  MSDA_CUDA_TRY(cudaStreamSynchronize(pl->s_run));
  MSDA_CUDA_TRY(cudaStreamSynchronize(pl->s_in));
  if (pl->trace && pl->ev_out) {
    auto at = [&](cudaEvent_t e) { float ms = 0.f; cudaEventElapsedTime(&ms, pl->ev_tables, e); return ms; };
    fprintf(stderr, "[msda host trace] ms after the level tables landed:  in = chunk inputs on device, run = kernels done, out = results on host\n");
    for (int b = 0; b < N; ++b) {
      fprintf(stderr, "  sample %d: value in %.3f\n", b, at(pl->ev_value[b]));
      for (int c = 0; c < pl->n_chunks; ++c) {
        const int k = b * pl->n_chunks + c;
        fprintf(stderr, "    chunk %d: in %.3f  run %.3f  out %.3f\n", c, at(pl->ev_in[k]), at(pl->ev_run[k]), at(pl->ev_out[k]));
      }
      if (bwd) fprintf(stderr, "    grad_value out %.3f\n", at(pl->ev_out[N * pl->n_chunks + b]));
    }
  }
  return MSDA_OK;
}

// On failure nothing may still be reading or writing the caller's buffers when we return.
static int host_finish(msda_host_plan* pl, int rc) {
  if (rc != MSDA_OK) {
    pl->tables_valid = 0;
    cudaStreamSynchronize(pl->s_in);
    cudaStreamSynchronize(pl->s_run);
    cudaStreamSynchronize(pl->s_out);
    cudaGetLastError();
  }
  return rc;
}

int msda_host_fwd_f32(msda_host_plan* pl, const float* value, const int64_t* shapes, const int64_t* lsi,
                      const float* loc, const float* aw, float* out) {
  if (!pl || !value || !shapes || !lsi || !loc || !aw || !out) return MSDA_ERR_NULL_POINTER;
  if (pl->value_bf16) return MSDA_ERR_UNSUPPORTED;
  MSDA_CUDA_TRY(cudaSetDevice(pl->device));
  return host_finish(pl, host_pipeline(pl, value, shapes, lsi, loc, aw, nullptr, out, nullptr, nullptr, nullptr));
}

int msda_host_fwd_bwd_f32(msda_host_plan* pl, const float* value, const int64_t* shapes,
                          const int64_t* lsi, const float* loc, const float* aw, const float* go,
                          float* out, float* gv, float* gl, float* ga) {
  if (!pl || !value || !shapes || !lsi || !loc || !aw || !go || !out || !gv || !gl || !ga)
    return MSDA_ERR_NULL_POINTER;
  if (pl->value_bf16) return MSDA_ERR_UNSUPPORTED;
  MSDA_CUDA_TRY(cudaSetDevice(pl->device));
  return host_finish(pl, host_pipeline(pl, value, shapes, lsi, loc, aw, go, out, gv, gl, ga));
}

int msda_host_fwd_bwd_bf16(msda_host_plan* pl, const uint16_t* value, const int64_t* shapes,
                           const int64_t* lsi, const float* loc, const float* aw, const float* go,
                           float* out, uint16_t* gv, float* gl, float* ga) {
  if (!pl || !value || !shapes || !lsi || !loc || !aw || !go || !out || !gv || !gl || !ga)
    return MSDA_ERR_NULL_POINTER;
  if (!pl->value_bf16) return MSDA_ERR_UNSUPPORTED;
  MSDA_CUDA_TRY(cudaSetDevice(pl->device));
  return host_finish(pl, host_pipeline(pl, value, shapes, lsi, loc, aw, go, out, gv, gl, ga));
}

// ---- several steps in flight -----------------------------------------------------------------------------------
// A pipe is `depth` plans (device buffers, events) on ONE set of three streams.  submit() enqueues a whole step and
// returns; because the streams are shared, the uploads of step i + 1 run while step i computes and downloads, so a
// stream of steps moves at the PCIe rate of large two-way copies instead of paying every step's fill and drain.
struct msda_host_pipe {
  int depth;
  msda_host_plan* slot[8];
  int pending[8];
};
void msda_host_pipe_destroy(msda_host_pipe* pp);

static int pipe_create(msda_host_pipe** out_pipe, int device, int depth, int N, int S, int M, int D, int L, int Lq, int P,
                       int value_bf16);

int msda_host_pipe_create(msda_host_pipe** out_pipe, int device, int depth, int N, int S, int M, int D, int L,
                          int Lq, int P) {
  return pipe_create(out_pipe, device, depth, N, S, M, D, L, Lq, P, 0);
}

int msda_host_pipe_create_bf16(msda_host_pipe** out_pipe, int device, int depth, int N, int S, int M, int D, int L,
                               int Lq, int P) {
  if (!d32_eligible(S > 0 ? S : 1, M > 0 ? M : 1, D, L > 0 ? L : 1, P > 0 ? P : 1)) return MSDA_ERR_UNSUPPORTED;
  return pipe_create(out_pipe, device, depth, N, S, M, D, L, Lq, P, 1);
}

static int pipe_create(msda_host_pipe** out_pipe, int device, int depth, int N, int S, int M, int D, int L, int Lq, int P,
                       int value_bf16) {
  if (!out_pipe) return MSDA_ERR_NULL_POINTER;
  *out_pipe = nullptr;
  if (depth < 1 || depth > 8) return MSDA_ERR_BAD_DIMS;
  msda_host_pipe* pp = new (std::nothrow) msda_host_pipe();
  if (!pp) return (int)cudaErrorMemoryAllocation;
  memset(pp, 0, sizeof(*pp));
  pp->depth = depth;
  int chunks = 0;      // 0: the whole step is one packet (steps overlap each other, so the copies can stay large)
  if (const char* env = getenv("MSDA_B200_PIPE_CHUNKS")) {   // >= 1: query chunks per sample inside a step, as in a plan
    const int v = atoi(env);
    if (v >= 0 && v <= 256) chunks = v;
  }
  for (int i = 0; i < depth; ++i) {
    const int rc = plan_create(&pp->slot[i], device, N, S, M, D, L, Lq, P, value_bf16, i ? pp->slot[0] : nullptr, chunks ? chunks : 1);
    if (rc) {
      msda_host_pipe_destroy(pp);
      return rc;
    }
    pp->slot[i]->whole_batch = chunks == 0;
  }
  *out_pipe = pp;
  return MSDA_OK;
}

void msda_host_pipe_destroy(msda_host_pipe* pp) {
  if (!pp) return;
  if (pp->slot[0]) {
    cudaSetDevice(pp->slot[0]->device);
    cudaStreamSynchronize(pp->slot[0]->s_in);
    cudaStreamSynchronize(pp->slot[0]->s_run);
    cudaStreamSynchronize(pp->slot[0]->s_out);
  }
  for (int i = pp->depth - 1; i >= 0; --i) msda_host_plan_destroy(pp->slot[i]);   // slot 0 (the streams' owner) last
  delete pp;
}

int msda_host_pipe_wait(msda_host_pipe* pp, int slot) {
  if (!pp) return MSDA_ERR_NULL_POINTER;
  if (slot >= pp->depth) return MSDA_ERR_BAD_DIMS;
  for (int i = slot < 0 ? 0 : slot; i < (slot < 0 ? pp->depth : slot + 1); ++i) {
    if (!pp->pending[i]) continue;
    MSDA_CUDA_TRY(cudaSetDevice(pp->slot[i]->device));
    MSDA_CUDA_TRY(cudaEventSynchronize(pp->slot[i]->ev_done));
    pp->pending[i] = 0;
  }
  return MSDA_OK;
}

int msda_host_pipe_submit_f32(msda_host_pipe* pp, int slot, const float* value, const int64_t* shapes, const int64_t* lsi,
                              const float* loc, const float* aw, const float* go, float* out, float* gv, float* gl,
                              float* ga) {
  if (!pp || !value || !shapes || !lsi || !loc || !aw || !go || !out || !gv || !gl || !ga) return MSDA_ERR_NULL_POINTER;
  if (slot < 0 || slot >= pp->depth) return MSDA_ERR_BAD_DIMS;
  if (pp->slot[slot]->value_bf16) return MSDA_ERR_UNSUPPORTED;
  int rc = msda_host_pipe_wait(pp, slot);     // the slot's device buffers are free once its previous step has landed
  if (rc) return rc;
  msda_host_plan* pl = pp->slot[slot];
  MSDA_CUDA_TRY(cudaSetDevice(pl->device));
  rc = host_finish(pl, host_pipeline(pl, value, shapes, lsi, loc, aw, go, out, gv, gl, ga, false));
  if (rc == MSDA_OK) pp->pending[slot] = 1;
  return rc;
}

int msda_host_pipe_submit_bf16(msda_host_pipe* pp, int slot, const uint16_t* value, const int64_t* shapes,
                               const int64_t* lsi, const float* loc, const float* aw, const float* go, float* out,
                               uint16_t* gv, float* gl, float* ga) {
  if (!pp || !value || !shapes || !lsi || !loc || !aw || !go || !out || !gv || !gl || !ga) return MSDA_ERR_NULL_POINTER;
  if (slot < 0 || slot >= pp->depth) return MSDA_ERR_BAD_DIMS;
  if (!pp->slot[slot]->value_bf16) return MSDA_ERR_UNSUPPORTED;
  int rc = msda_host_pipe_wait(pp, slot);
  if (rc) return rc;
  msda_host_plan* pl = pp->slot[slot];
  MSDA_CUDA_TRY(cudaSetDevice(pl->device));
  rc = host_finish(pl, host_pipeline(pl, value, shapes, lsi, loc, aw, go, out, gv, gl, ga, false));
  if (rc == MSDA_OK) pp->pending[slot] = 1;
  return rc;
}

int msda_host_pipe_traffic(const msda_host_pipe* pp, int64_t* h2d, int64_t* d2h) {
  if (!pp || !pp->slot[0]) return MSDA_ERR_NULL_POINTER;
  return msda_host_plan_traffic(pp->slot[0], h2d, d2h);
}

}  // extern "C"
