// D = 32 kernels -- the shipped head geometry (util/settings.py:102-120: d_model 256 / 8 heads)
// and the performance path of this library.
//
// Work decomposition (forward and backward):
//   * an ITEM is one (sample b, query q, head m); it owns 32 output channels = one 128 B row.
//   * 8 lanes serve one item, each lane owns 4 consecutive channels, so one warp instruction
//     moves FOUR complete value rows (fp32: 4 x 128 B with LDG.128; bf16: 4 x 64 B with LDG.64)
//     and one red.global.add.v4.f32 per lane replaces four scalar atomics in the backward.
//   * a warp therefore works on 4 consecutive items (same query, 4 neighbouring heads when
//     M % 4 == 0), whose sampling locations / weights / outputs are contiguous in memory:
//     every streaming access of the warp is one fully coalesced segment, and the four rows a
//     gather instruction fetches are one contiguous 512 B piece of `value`.
//
// Two phases per warp pass, separated only by __syncwarp (no block barrier in the loop):
//   A  "geometry": the 4*L*P samples of the pass are spread over the 32 lanes (one thread per
//      sample, no redundancy); each runs sample_geom() and parks a 16 B record
//          { row offset of corner (h_low, w_low) << 4 | corner mask,  f0, f1, f2 }
//      in the warp's private slice of shared memory.  Group slices are padded by one record so
//      that the four groups' broadcast reads fall into distinct banks.
//   B  "gather": every 8-lane group walks its item's L*P records (ONE broadcast LDS.128 per
//      sample), rebuilds the four bilinear weights (5 flops) and corner offsets (3 adds), issues
//      the 4 corner row loads and accumulates.
//
// Both kernels are bound by the SM's load/store data path (one 128 B row per wavefront) and,
// for the backward, by the SM->L2 reduction path; everything else is arranged to stay out of
// their way (profiles/ has the ncu evidence).
//
// The reference instead recomputes all of phase A in each of the 32 channel threads and issues
// 64 predicated scalar loads per thread (cuh:272-296); its backward adds 2*L*P block barriers
// and a serial thread-0 reduction per block (cuh:376-394).
#pragma once

#include "msda_device.cuh"
#include "msda_generic.cuh"

namespace msda {

// Minimum resident CTAs per SM the compiler must leave room for (register cap = 65536 / (256*n)).
#ifndef MSDA_FWD_MINB
#define MSDA_FWD_MINB 6
#endif
#ifndef MSDA_BWD_MINB
#define MSDA_BWD_MINB 4
#endif
// The fused-prologue forward carries the softmax / location arithmetic in phase A: at 40 registers it spills
// 112 B per thread; 4 CTAs per SM (64 registers, no spills) measured fastest (C2 194 -> 179 us, C3 encoder 136 -> 126;
// 5 CTAs 188 / 131, 3 CTAs 188 / 132).
#ifndef MSDA_FWD_FUSE_MINB
#define MSDA_FWD_FUSE_MINB 4
#endif
// Tuning switches, defaults = what measured fastest on B200 at the DETR-encoder shape
// (gpurun_out A/B logs summarised in profiles/r01_tuning_matrix.md):
//   *_PREFETCH  software-pipelined phase A (next pass's locations/weights requested before the gathers)
//   *_WAVES     grid cap in units of "CTAs the device holds at once": 1 = persistent grid
// Tried and not kept (same file): backward load + dot + red under one predicate, backward prefetch,
// persistent backward grid, 48 / 80 registers -- see the tuning matrix.
#ifndef MSDA_FWD_PREFETCH
#define MSDA_FWD_PREFETCH 1
#endif
#ifndef MSDA_BWD_PREFETCH
#define MSDA_BWD_PREFETCH 0
#endif
#ifndef MSDA_FWD_WAVES
#define MSDA_FWD_WAVES 1
#endif
// bf16 values: rows are half as long, the gathers finish sooner and the forward prefers short-lived
// CTAs without the pipelined phase A (C2 bf16: 150 -> 135 us)
#ifndef MSDA_FWD_PREFETCH_BF16
#define MSDA_FWD_PREFETCH_BF16 0
#endif
#ifndef MSDA_FWD_WAVES_BF16
#define MSDA_FWD_WAVES_BF16 2
#endif
#ifndef MSDA_BWD_WAVES
#define MSDA_BWD_WAVES 11
#endif

constexpr int kD32Warps = 8;                  // warps per CTA
constexpr int kD32Threads = kD32Warps * kWarp;
constexpr int kItemsPerWarp = 4;              // 8 lanes per item
constexpr int kMaxLevelsD32 = 8;
constexpr int kMaxLPD32 = 32;                 // L*P limit of the specialised kernels
// Row offsets are carried as (token*M) in a signed 28-bit field.
constexpr long long kMaxTokenHeadsD32 = 1ll << 27;

struct LevelTab {
  int H[kMaxLevelsD32], W[kMaxLevelsD32], start[kMaxLevelsD32];
};

struct D32Params {
  const void* value;
  const int64_t* shapes;
  const int64_t* lsi;
  const float* loc;
  const float* aw;
  const float* grad_out;  // backward only
  float* out;             // forward only
  void* grad_value;       // backward only
  float* grad_loc;
  float* grad_aw;
  int N, S, M, L, Lq, P;
  long long n_items;
  // Fused prologue (kFuse kernels): `loc` then holds the RAW sampling offsets [N,Lq,M,L,P,2] and
  // `aw` the RAW attention logits [N,Lq,M,L*P]; ref = reference points [N,Lq,L,2].  In the backward
  // grad_loc / grad_aw then receive d/d offsets and d/d logits, and grad_ref (optional, zeroed by
  // the caller) d/d reference points.
  const float* ref;
  float* grad_ref;
};

__host__ __device__ constexpr int rec_stride(int LP) { return LP | 1; }  // odd => conflict-free groups

__device__ __forceinline__ void load_level_tab(LevelTab& tab, const D32Params& p) {
  if (threadIdx.x < p.L) {
    tab.H[threadIdx.x] = (int)p.shapes[2 * threadIdx.x];
    tab.W[threadIdx.x] = (int)p.shapes[2 * threadIdx.x + 1];
    tab.start[threadIdx.x] = (int)p.lsi[threadIdx.x];
  }
  __syncthreads();
}

// Phase A for one sample.
//   forward : f = (a*hh, a*lh, lw)      backward : f = (lh, lw, a)
// A rejected sample gets mask 0 and zero payload: phase B then loads nothing and adds zero.
template <bool kBackward, bool kKeepWeight = false>
__device__ __forceinline__ int4 make_record(float x, float y, float a, int H, int W, int start, int M) {
  SampleGeom<float> g;
  // (fused backward: the softmax gradient needs the weight of rejected samples too)
  if (!sample_geom<float>(x, y, H, W, g)) return make_int4(0, 0, 0, kKeepWeight ? __float_as_int(a) : 0);
  const int row_m = (start + g.h_low * W + g.w_low) * M;  // may be negative at the top/left border
  int4 r;
  r.x = row_m * 16 + (int)g.mask;
  if (kBackward) {
    r.y = __float_as_int(g.lh);
    r.z = __float_as_int(g.lw);
    r.w = __float_as_int(a);
  } else {
    r.y = __float_as_int(a * (1.f - g.lh));
    r.z = __float_as_int(a * g.lh);
    r.w = __float_as_int(g.lw);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Fused prologue: what MSDeformAttn.forward does between its Linear layers and the operator
// (/root/reference/models/ops/modules/ms_deform_attn.py:100-108), done in phase A instead of in
// separate elementwise kernels that write locations / weights to HBM only to be read back here:
//   attention_weights = softmax(logits over L*P)                                   (:101-102)
//   sampling_locations = reference_points + sampling_offsets / (W_l, H_l)          (:104-108)
// Location arithmetic is torch's, operation for operation (one correctly rounded division, one
// correctly rounded addition, no contraction), so corner indices are bit-identical to the unfused path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 fused_location(const D32Params& p, long long item, int l, float2 off,
                                                 int H, int W) {
  const float2 r = __ldg(reinterpret_cast<const float2*>(p.ref) + (item / p.M) * p.L + l);
  return make_float2(__fadd_rn(r.x, __fdiv_rn(off.x, (float)W)), __fadd_rn(r.y, __fdiv_rn(off.y, (float)H)));
}
// the same with the reference point already on chip
__device__ __forceinline__ float2 fused_location(float2 r, float2 off, int H, int W) {
  return make_float2(__fadd_rn(r.x, __fdiv_rn(off.x, (float)W)), __fadd_rn(r.y, __fdiv_rn(off.y, (float)H)));
}

// softmax weight of this lane's logit within an aligned segment of kSeg lanes (kSeg a power of two);
// dead lanes pass -inf and get 0.
template <int kSeg>
__device__ __forceinline__ float segment_softmax(float logit) {
  float m = logit;
#pragma unroll
  for (int o = 1; o < kSeg; o <<= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
  const float e = (logit == -INFINITY) ? 0.f : expf(logit - m);
  float sum = e;
#pragma unroll
  for (int o = 1; o < kSeg; o <<= 1) sum += __shfl_xor_sync(kFull, sum, o);
  return __fdiv_rn(e, sum);
}

// any L*P: every sample thread walks its item's logits (L1-resident) -- not the shipped geometry
__device__ __forceinline__ float looped_softmax(const float* logits, int LP, float mine) {
  float m = -INFINITY;
  for (int j = 0; j < LP; ++j) m = fmaxf(m, __ldg(logits + j));
  float sum = 0.f;
  for (int j = 0; j < LP; ++j) sum += expf(__ldg(logits + j) - m);
  return __fdiv_rn(expf(mine - m), sum);
}

// Phase A for the whole warp pass: 4*LP samples over 32 lanes, coalesced loc / weight loads.
// kLPc: compile-time L*P (0 = runtime).  The fused softmax runs on shuffles when an item's logits
// occupy an aligned power-of-two run of lanes and every lane makes the same number of trips.
template <bool kBackward, bool kFuse, int kLPc>
__device__ __forceinline__ void geometry_phase(int4* warp_recs, const D32Params& p, const LevelTab& tab,
                                               long long item0, int LP, int lane) {
  constexpr bool kSegment = kFuse && kLPc > 0 && (kLPc & (kLPc - 1)) == 0 && (kItemsPerWarp * kLPc) % kWarp == 0;
  const int n_samples = kItemsPerWarp * LP;
  const int stride = rec_stride(LP);
  const long long s_base = item0 * LP;
  const long long s_end = p.n_items * LP;
  for (int s = lane; s < n_samples; s += kWarp) {
    const long long gs = s_base + s;
    const bool in = gs < s_end;
    if (!kSegment && !in) break;
    const int grp = s / LP;
    const int ls = s - grp * LP;
    const int l = ls / p.P;
    float2 xy = make_float2(0.f, 0.f);
    float a = -INFINITY;
    if (in) {
      xy = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
      a = ld_stream(p.aw + gs);
    }
    if (kFuse) {
      a = kSegment ? segment_softmax<(kSegment ? kLPc : 1)>(a) : looped_softmax(p.aw + (gs - ls), LP, a);
      if (in) xy = fused_location(p, item0 + grp, l, xy, tab.H[l], tab.W[l]);
    }
    if (in)
      warp_recs[grp * stride + ls] =
          make_record<kBackward, kFuse && kBackward>(xy.x, xy.y, a, tab.H[l], tab.W[l], tab.start[l], p.M);
  }
}

// Software-pipelined phase A (compile-time L*P with 4*L*P a multiple of 32): the raw locations
// and weights of the NEXT pass are requested before the current pass starts gathering, so their
// HBM latency (they are read exactly once, always cold) hides behind the gathers instead of
// heading every pass.
template <int kLP>
struct RawSamples {
  static constexpr int kPerLane = kItemsPerWarp * kLP / kWarp;
  float2 xy[kPerLane];
  float a[kPerLane];
  __device__ __forceinline__ void load(const D32Params& p, long long item0, int lane) {
    const long long s_base = item0 * kLP, s_end = p.n_items * kLP;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const long long gs = s_base + lane + i * kWarp;
      if (gs < s_end) {
        xy[i] = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
        a[i] = ld_stream(p.aw + gs);
      }
    }
  }
  template <bool kBackward, int kP, bool kFuse>
  __device__ __forceinline__ void park(int4* warp_recs, const D32Params& p, const LevelTab& tab,
                                       long long item0, int lane, int M) const {
    const long long s_base = item0 * kLP, s_end = p.n_items * kLP;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int s = lane + i * kWarp;
      const bool in = s_base + s < s_end;
      const int grp = s / kLP, ls = s % kLP, l = ls / kP;
      float2 loc = xy[i];
      float w = a[i];
      if (kFuse) {
        // the L*P logits of an item sit in L*P consecutive, aligned lanes
        static_assert(!kFuse || (kLP & (kLP - 1)) == 0, "fused pipeline needs a power-of-two L*P");
        w = segment_softmax<kLP>(in ? w : -INFINITY);
        if (in) loc = fused_location(p, item0 + grp, l, loc, tab.H[l], tab.W[l]);
      }
      if (in)
        warp_recs[grp * rec_stride(kLP) + ls] =
            make_record<kBackward, kFuse && kBackward>(loc.x, loc.y, w, tab.H[l], tab.W[l], tab.start[l], M);
    }
  }
};

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ float4 scale4(float s, const float4& v) {
  return make_float4(s * v.x, s * v.y, s * v.z, s * v.w);
}

// Per-level row stride W*M, in registers when the level count is a compile-time constant.
template <int kL>
struct LevelStride {
  int wm[kL > 0 ? kL : 1];
  __device__ __forceinline__ void init(const LevelTab& tab, int M) {
#pragma unroll
    for (int l = 0; l < (kL > 0 ? kL : 1); ++l) wm[l] = tab.W[l] * M;
  }
};

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
template <typename VT, int kL, int kP, int kM, bool kFuse = false>  // compile-time levels / points / heads (0: runtime)
__global__ void __launch_bounds__(kD32Threads, kFuse ? MSDA_FWD_FUSE_MINB : MSDA_FWD_MINB)
fwd_d32_kernel(const D32Params p) {
  using C = Chunk<VT>;
  using Vec = typename C::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);

  constexpr bool kStatic = kL > 0;
  const int P = kStatic ? kP : p.P;
  const int LP = kStatic ? kL * kP : p.L * p.P;
  const int stride = rec_stride(LP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * (kItemsPerWarp * stride);
  const int4* my_recs = warp_recs + group * stride;
  const int M = kM > 0 ? kM : p.M;   // compile-time M turns the x+1 corner into an immediate offset
  LevelStride<kL> lv;
  if (kStatic) lv.init(tab, M);

  const long long n_pass = (p.n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  const long long pass_stride = (long long)gridDim.x * kD32Warps;
  constexpr bool kPrefetch = sizeof(typename C::Vec) == 16 ? MSDA_FWD_PREFETCH : MSDA_FWD_PREFETCH_BF16;
  constexpr bool kPipe = kStatic && kPrefetch && (kItemsPerWarp * kL * kP) % kWarp == 0 &&
                         (!kFuse || ((kL * kP) & (kL * kP - 1)) == 0);
  RawSamples<kPipe ? kL * kP : kWarp> raw;
  const long long pass0 = (long long)blockIdx.x * kD32Warps + warp;
  if (kPipe && pass0 < n_pass) raw.load(p, pass0 * kItemsPerWarp, lane);
  for (long long pass = pass0; pass < n_pass; pass += pass_stride) {
    const long long item0 = pass * kItemsPerWarp;
    if (kPipe) {
      raw.template park<false, (kP > 0 ? kP : 1), kFuse>(warp_recs, p, tab, item0, lane, M);
      if (pass + pass_stride < n_pass) raw.load(p, (pass + pass_stride) * kItemsPerWarp, lane);
    } else {
      geometry_phase<false, kFuse, (kStatic ? kL * kP : 0)>(warp_recs, p, tab, item0, LP, lane);
    }
    __syncwarp();

    const long long item = item0 + group;
    if (item < p.n_items) {
      const int m = (int)(item % M);
      const int b = (int)(item / ((long long)M * p.Lq));
      // Chunk units: one token-head row is 8 chunks; row_m counts token*M, so row = row_m * 8.
      const Vec* vb = reinterpret_cast<const Vec*>(p.value) + ((long long)b * p.S * M + m) * 8 + sub;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int l = 0; l < (kStatic ? kL : 1); ++l) {
        const int s_lo = kStatic ? l * kP : 0;
        const int s_hi = kStatic ? (l + 1) * kP : LP;
#pragma unroll
        for (int s = s_lo; s < s_hi; ++s) {
          const int4 r = my_recs[s];
          const int wm = kStatic ? lv.wm[l] : tab.W[s / P] * M;
          const int row = r.x >> 4;
          const float ah = __int_as_float(r.y), al = __int_as_float(r.z), lw = __int_as_float(r.w);
          const float w2 = ah * lw, w4 = al * lw;
          const float w1 = ah - w2, w3 = al - w4;   // a*hh*(1-lw), a*lh*(1-lw)
          const Vec* r0 = vb + (long long)row * 8;
          const Vec* r1 = r0 + (long long)wm * 8;
          // load and accumulate under ONE predicate per corner: no zero-fill of the landing
          // registers, and an invalid corner is never read (zero padding = "do not load")
          if (r.x & 1) fma4(acc, w1, C::load(r0));
          if (r.x & 2) fma4(acc, w2, C::load(r0 + M * 8));
          if (r.x & 4) fma4(acc, w3, C::load(r1));
          if (r.x & 8) fma4(acc, w4, C::load(r1 + M * 8));
        }
      }
      st_stream(reinterpret_cast<float4*>(p.out) + item * 8 + sub, acc);
    }
    __syncwarp();  // records are rewritten by the next pass
  }
}

// ---------------------------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(kFull, v, 4);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 1);
  return v;
}

// Sum four per-lane values over the 8 lanes of a group with 4 shuffles instead of 12: after the
// call, lane `sub` holds the group total of v[sub >> 1] (both lanes of a pair hold the same one).
__device__ __forceinline__ float group8_sum4(const float (&v)[4], int sub) {
  const bool hi4 = sub & 4, hi2 = sub & 2;
  const float keep0 = hi4 ? v[2] : v[0], send0 = hi4 ? v[0] : v[2];
  const float keep1 = hi4 ? v[3] : v[1], send1 = hi4 ? v[1] : v[3];
  const float r0 = keep0 + __shfl_xor_sync(kFull, send0, 4);
  const float r1 = keep1 + __shfl_xor_sync(kFull, send1, 4);
  const float keep = hi2 ? r1 : r0, send = hi2 ? r0 : r1;
  float t = keep + __shfl_xor_sync(kFull, send, 2);
  t += __shfl_xor_sync(kFull, t, 1);
  return t;
}

// One sample of the backward for one lane: 4 corner loads, 4 vector reductions into grad_value,
// and this lane's partial sums for grad_attn_weight / grad_loc.
template <typename VT, typename GT>
__device__ __forceinline__ void bwd_sample(const int4 r, const typename Chunk<VT>::Vec* vb,
                                           typename Chunk<GT>::Vec* gvb, const float4& g, int M, int wm,
                                           float Wf, float Hf, bool live, float& pa, float& px, float& py) {
  using C = Chunk<VT>;
  using G = Chunk<GT>;
  const int row = r.x >> 4;
  const float lh = __int_as_float(r.y), lw = __int_as_float(r.z), a = __int_as_float(r.w);
  const float hh = 1.f - lh, hw = 1.f - lw;
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  const long long o0 = (long long)row * 8, o1 = o0 + (long long)wm * 8;
  const int dx = M * 8;
  const bool ok1 = live && (r.x & 1), ok2 = live && (r.x & 2), ok3 = live && (r.x & 4), ok4 = live && (r.x & 8);
  const float4 t = scale4(a, g);  // top_grad * attn_weight (cuh:115)
  // An invalid corner is never read and gets no atomic (cuh:119-151).
  float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
  if (ok1) v1 = C::load(vb + o0);
  if (ok2) v2 = C::load(vb + o0 + dx);
  if (ok3) v3 = C::load(vb + o1);
  if (ok4) v4 = C::load(vb + o1 + dx);
  if (ok1) G::red_add(gvb + o0, scale4(w1, t));
  if (ok2) G::red_add(gvb + o0 + dx, scale4(w2, t));
  if (ok3) G::red_add(gvb + o1, scale4(w3, t));
  if (ok4) G::red_add(gvb + o1 + dx, scale4(w4, t));
  const float d1 = dot4(g, v1), d2 = dot4(g, v2), d3 = dot4(g, v3), d4 = dot4(g, v4);
  pa = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;                    // cuh:156
  px = (hh * (d2 - d1) + lh * (d4 - d3)) * a * Wf;               // cuh:157
  py = (hw * (d3 - d1) + lw * (d4 - d2)) * a * Hf;               // cuh:158
}

// VT: storage type of value; GT: type grad_value is accumulated in (float: red.v4.f32 -- also used
// for bf16 values accumulating into an fp32 scratch; __nv_bfloat16: red.v2.bf16x2 straight into a
// bf16 grad_value, faster but every add rounds the running sum to 8 bits).
template <typename VT, typename GT, int kL, int kP, int kM, bool kFuse = false>
__global__ void __launch_bounds__(kD32Threads, MSDA_BWD_MINB)
bwd_d32_kernel(const D32Params p) {
  using C = Chunk<VT>;
  using Vec = typename C::Vec;
  using GVec = typename Chunk<GT>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);

  constexpr bool kStatic = kL > 0 && (kP % 4 == 0);
  const int P = kStatic ? kP : p.P;
  const int LP = kStatic ? kL * kP : p.L * p.P;
  const int stride = rec_stride(LP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * (kItemsPerWarp * stride);
  const int4* my_recs = warp_recs + group * stride;
  int4* my_recs_w = warp_recs + group * stride;
  const int M = kM > 0 ? kM : p.M;   // compile-time M turns the x+1 corner into an immediate offset

  const long long n_pass = (p.n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  const long long pass_stride = (long long)gridDim.x * kD32Warps;
  constexpr bool kPipe = kStatic && MSDA_BWD_PREFETCH && (kItemsPerWarp * kL * kP) % kWarp == 0 &&
                         (!kFuse || ((kL * kP) & (kL * kP - 1)) == 0);
  RawSamples<kPipe ? kL * kP : kWarp> raw;
  const long long pass0 = (long long)blockIdx.x * kD32Warps + warp;
  if (kPipe && pass0 < n_pass) raw.load(p, pass0 * kItemsPerWarp, lane);
  for (long long pass = pass0; pass < n_pass; pass += pass_stride) {
    const long long item0 = pass * kItemsPerWarp;
    if (kPipe) {
      raw.template park<true, (kP > 0 ? kP : 1), kFuse>(warp_recs, p, tab, item0, lane, M);
      if (pass + pass_stride < n_pass) raw.load(p, (pass + pass_stride) * kItemsPerWarp, lane);
    } else {
      geometry_phase<true, kFuse, (kStatic ? kL * kP : 0)>(warp_recs, p, tab, item0, LP, lane);
    }
    __syncwarp();

    // Groups past the end still take part in the shuffles (with zero data).
    const long long item = item0 + group;
    const bool live = item < p.n_items;
    const long long item_c = live ? item : (p.n_items - 1);
    const int m = (int)(item_c % M);
    const int b = (int)(item_c / ((long long)M * p.Lq));
    const long long head_off = ((long long)b * p.S * M + m) * 8 + sub;
    const Vec* vb = reinterpret_cast<const Vec*>(p.value) + head_off;
    GVec* gvb = reinterpret_cast<GVec*>(p.grad_value) + head_off;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g = ld_stream(reinterpret_cast<const float4*>(p.grad_out) + item * 8 + sub);
    float2* gl_out = reinterpret_cast<float2*>(p.grad_loc) + item_c * LP;
    float* ga_out = p.grad_aw + item_c * LP;

    if (kStatic) {
      // 4 samples at a time: 12 partial sums reduced with 12 shuffles; lane pair (sub>>1) ends up
      // with sample s0 + (sub>>1); the even lane stores grad_loc, the odd lane grad_attn_weight.
#pragma unroll
      for (int l = 0; l < (kStatic ? kL : 1); ++l) {
        const int wm = tab.W[l] * M;
        const float Wf = (float)tab.W[l], Hf = (float)tab.H[l];
#pragma unroll
        for (int q4 = 0; q4 < (kStatic ? kP / 4 : 1); ++q4) {
          const int s0 = l * kP + q4 * 4;
          // Transposed reduction of the 4 samples' (weight, x, y) partial sums over the 8 lanes, taken in two halves
          // (samples 0 + 2, then 1 + 3: the pairs the first exchange step combines) so that at most six partial sums
          // are alive next to a sample's four value chunks -- the all-at-once form spilled inside this loop.
          const bool hi4 = sub & 4, hi2 = sub & 2;
          float ra[2], rx[2], ry[2];
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            float pa0, px0, py0, pa2, px2, py2;
            bwd_sample<VT, GT>(my_recs[s0 + hlf], vb, gvb, g, M, wm, Wf, Hf, live, pa0, px0, py0);
            bwd_sample<VT, GT>(my_recs[s0 + hlf + 2], vb, gvb, g, M, wm, Wf, Hf, live, pa2, px2, py2);
            ra[hlf] = (hi4 ? pa2 : pa0) + __shfl_xor_sync(kFull, hi4 ? pa0 : pa2, 4);
            rx[hlf] = (hi4 ? px2 : px0) + __shfl_xor_sync(kFull, hi4 ? px0 : px2, 4);
            ry[hlf] = (hi4 ? py2 : py0) + __shfl_xor_sync(kFull, hi4 ? py0 : py2, 4);
          }
          float ta = (hi2 ? ra[1] : ra[0]) + __shfl_xor_sync(kFull, hi2 ? ra[0] : ra[1], 2);
          float tx = (hi2 ? rx[1] : rx[0]) + __shfl_xor_sync(kFull, hi2 ? rx[0] : rx[1], 2);
          float ty = (hi2 ? ry[1] : ry[0]) + __shfl_xor_sync(kFull, hi2 ? ry[0] : ry[1], 2);
          ta += __shfl_xor_sync(kFull, ta, 1);
          tx += __shfl_xor_sync(kFull, tx, 1);
          ty += __shfl_xor_sync(kFull, ty, 1);
          const int s = s0 + (sub >> 1);
          if (!kFuse) {
            if (live) {
              if (sub & 1) st_stream(ga_out + s, ta);
              else st_stream(gl_out + s, make_float2(tx, ty));
            }
          } else {
            // d loc / d offset = 1 / (W, H); d/d reference point = sum over this level's points (and over
            // heads: accumulated with scalar reductions, 2 per (query, head, level)); the gradient of
            // the post-softmax weight is parked in the record for the softmax backward below.
            if (live) {
              if (sub & 1) my_recs_w[s].y = __float_as_int(ta);
              else st_stream(gl_out + s, make_float2(__fdiv_rn(tx, Wf), __fdiv_rn(ty, Hf)));
            }
            if (p.grad_ref) {
              float rx = (sub & 1) ? 0.f : tx, ry = (sub & 1) ? 0.f : ty;
              rx = group8_sum(rx);
              ry = group8_sum(ry);
              if (live && sub == 0) {
                float* gr = p.grad_ref + ((item / M) * p.L + l) * 2;
                atomicAdd(gr, rx);
                atomicAdd(gr + 1, ry);
              }
            }
          }
        }
      }
    } else {
      for (int s = 0; s < LP; ++s) {
        const int l = s / P;
        float pa, px, py;
        bwd_sample<VT, GT>(my_recs[s], vb, gvb, g, M, tab.W[l] * M, (float)tab.W[l], (float)tab.H[l], live,
                       pa, px, py);
        pa = group8_sum(pa);
        px = group8_sum(px);
        py = group8_sum(py);
        if (live && sub == (s & 7)) {
          if (!kFuse) {
            st_stream(gl_out + s, make_float2(px, py));
            st_stream(ga_out + s, pa);
          } else {
            const float Wf = (float)tab.W[l], Hf = (float)tab.H[l];
            st_stream(gl_out + s, make_float2(__fdiv_rn(px, Wf), __fdiv_rn(py, Hf)));
            my_recs_w[s].y = __float_as_int(pa);
            if (p.grad_ref) {
              float* gr = p.grad_ref + ((item / M) * p.L + l) * 2;
              atomicAdd(gr, px);
              atomicAdd(gr + 1, py);
            }
          }
        }
      }
    }
    if (kFuse) {
      // softmax backward (modules/ms_deform_attn.py:102): d/d logit_s = a_s * (ga_s - sum_j a_j * ga_j),
      // a_s in record.w (kept for rejected samples too), ga_s parked in record.y above.
      __syncwarp();
      float part = 0.f;
      for (int s = sub; s < LP; s += 8) {
        const int4 r = my_recs[s];
        part = fmaf(__int_as_float(r.w), __int_as_float(r.y), part);
      }
      const float dot = group8_sum(part);
      if (live)
        for (int s = sub; s < LP; s += 8) {
          const int4 r = my_recs[s];
          st_stream(ga_out + s, __int_as_float(r.w) * (__int_as_float(r.y) - dot));
        }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Small problems (decoder queries: a few thousand (query, head) items): "split" kernels.
//
// With 4 items per warp a 900-query decoder call fills a quarter of the machine and every warp walks
// its 16 samples one after the other, each a cold L2/HBM round trip: the call costs the LATENCY of
// that chain (~25 us), not bandwidth.  Here a warp takes ONE item and its four 8-lane groups take
// every fourth sample, so the chain is 4x shorter and there are 4x as many warps in flight; the
// four partial rows are summed with two shuffles per register.  Any L, P with L*P <= 32; runtime M.
// ---------------------------------------------------------------------------------------------
template <bool kBackward, bool kFuse>
__device__ __forceinline__ void split_geometry(int4* warp_recs, const D32Params& p, const LevelTab& tab,
                                               long long item, int LP, int lane) {
  const bool in = lane < LP;
  const long long gs = item * LP + lane;
  const int l = in ? lane / p.P : 0;
  float2 xy = make_float2(0.f, 0.f);
  float a = -INFINITY;
  if (in) {
    xy = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
    a = ld_stream(p.aw + gs);
  }
  if (kFuse) {
    a = segment_softmax<kWarp>(a);     // the item's logits are lanes [0, LP); the rest pass -inf
    if (in) xy = fused_location(p, item, l, xy, tab.H[l], tab.W[l]);
  }
  if (in)
    warp_recs[lane] = make_record<kBackward, kFuse && kBackward>(xy.x, xy.y, a, tab.H[l], tab.W[l], tab.start[l], p.M);
}

template <typename VT, bool kFuse = false>
__global__ void __launch_bounds__(kD32Threads, 4)
fwd_d32_split_kernel(const D32Params p) {
  using C = Chunk<VT>;
  using Vec = typename C::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);
  const int P = p.P, LP = p.L * p.P, M = p.M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * kMaxLPD32;
  for (long long item = (long long)blockIdx.x * kD32Warps + warp; item < p.n_items;
       item += (long long)gridDim.x * kD32Warps) {
    split_geometry<false, kFuse>(warp_recs, p, tab, item, LP, lane);
    __syncwarp();
    const int m = (int)(item % M);
    const int b = (int)(item / ((long long)M * p.Lq));
    const Vec* vb = reinterpret_cast<const Vec*>(p.value) + ((long long)b * p.S * M + m) * 8 + sub;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = group; s < LP; s += kItemsPerWarp) {
      const int4 r = warp_recs[s];
      const int wm = tab.W[s / P] * M;
      const int row = r.x >> 4;
      const float ah = __int_as_float(r.y), al = __int_as_float(r.z), lw = __int_as_float(r.w);
      const float w2 = ah * lw, w4 = al * lw;
      const float w1 = ah - w2, w3 = al - w4;
      const Vec* r0 = vb + (long long)row * 8;
      const Vec* r1 = r0 + (long long)wm * 8;
      float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
      if (r.x & 1) v1 = C::load(r0);
      if (r.x & 2) v2 = C::load(r0 + M * 8);
      if (r.x & 4) v3 = C::load(r1);
      if (r.x & 8) v4 = C::load(r1 + M * 8);
      fma4(acc, w1, v1);
      fma4(acc, w2, v2);
      fma4(acc, w3, v3);
      fma4(acc, w4, v4);
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      acc.x += __shfl_xor_sync(kFull, acc.x, o);
      acc.y += __shfl_xor_sync(kFull, acc.y, o);
      acc.z += __shfl_xor_sync(kFull, acc.z, o);
      acc.w += __shfl_xor_sync(kFull, acc.w, o);
    }
    if (group == 0) st_stream(reinterpret_cast<float4*>(p.out) + item * 8 + sub, acc);
    __syncwarp();
  }
}

template <typename VT, typename GT>
__global__ void __launch_bounds__(kD32Threads, 3)
bwd_d32_split_kernel(const D32Params p) {
  using Vec = typename Chunk<VT>::Vec;
  using GVec = typename Chunk<GT>::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);
  const int P = p.P, LP = p.L * p.P, M = p.M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * kMaxLPD32;
  for (long long item = (long long)blockIdx.x * kD32Warps + warp; item < p.n_items;
       item += (long long)gridDim.x * kD32Warps) {
    split_geometry<true, false>(warp_recs, p, tab, item, LP, lane);
    __syncwarp();
    const int m = (int)(item % M);
    const int b = (int)(item / ((long long)M * p.Lq));
    const long long head_off = ((long long)b * p.S * M + m) * 8 + sub;
    const Vec* vb = reinterpret_cast<const Vec*>(p.value) + head_off;
    GVec* gvb = reinterpret_cast<GVec*>(p.grad_value) + head_off;
    const float4 g = ld_stream(reinterpret_cast<const float4*>(p.grad_out) + item * 8 + sub);
    float2* gl_out = reinterpret_cast<float2*>(p.grad_loc) + item * LP;
    float* ga_out = p.grad_aw + item * LP;
    // every group runs the same number of rounds: the shuffles below are warp-wide
    for (int s0 = 0; s0 < LP; s0 += kItemsPerWarp) {
      const int s = s0 + group;
      const bool live = s < LP;
      const int l = live ? s / P : 0;
      const int4 r = live ? warp_recs[s] : make_int4(0, 0, 0, 0);
      float pa, px, py;
      bwd_sample<VT, GT>(r, vb, gvb, g, M, tab.W[l] * M, (float)tab.W[l], (float)tab.H[l], live, pa, px, py);
      pa = group8_sum(pa);
      px = group8_sum(px);
      py = group8_sum(py);
      if (live && sub == 0) {
        st_stream(gl_out + s, make_float2(px, py));
        st_stream(ga_out + s, pa);
      }
    }
    __syncwarp();
  }
}

// fp32 scratch -> bf16 grad_value (the accurate bf16 mode: one rounding at the end).
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst,
                                                            long long n_vec4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld_stream(src + i);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned*>(&lo);
    o.y = *reinterpret_cast<const unsigned*>(&hi);
    dst[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Value producer side: the module's value path is
//     value = value_proj(x); value = value.masked_fill(padding_mask[..., None], 0)        (ms_deform_attn.py:96-98)
// followed, for bf16 kernels, by a cast.  These two kernels do mask + cast in ONE pass over the projection's fp32
// output (read 4 B, write 2 B per element; the torch sequence reads/writes 14 B), and the matching backward
// (bf16 grad_value -> fp32 gradient of the projection output, masked rows zero).
// rows = N*S tokens, `cols` channels per token (multiple of 8), mask[row] != 0 -> padded token.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) value_pack_bf16_kernel(const float4* __restrict__ src, const uint8_t* __restrict__ mask,
                                                              uint4* __restrict__ dst, long long n_vec8, int vec8_per_row) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec8; i += (long long)gridDim.x * blockDim.x) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (!mask || !mask[i / vec8_per_row]) {
      const float4 a = ld_stream(src + 2 * i), b = ld_stream(src + 2 * i + 1);
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      o.x = *reinterpret_cast<const unsigned*>(&p0); o.y = *reinterpret_cast<const unsigned*>(&p1);
      o.z = *reinterpret_cast<const unsigned*>(&p2); o.w = *reinterpret_cast<const unsigned*>(&p3);
    }
    dst[i] = o;
  }
}

__global__ void __launch_bounds__(256) value_unpack_grad_kernel(const uint4* __restrict__ src, const uint8_t* __restrict__ mask,
                                                                float4* __restrict__ dst, long long n_vec8, int vec8_per_row) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec8; i += (long long)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (!mask || !mask[i / vec8_per_row]) {
      const uint4 r = __ldg(src + i);
      a.x = __uint_as_float(r.x << 16); a.y = __uint_as_float(r.x & 0xffff0000u);
      a.z = __uint_as_float(r.y << 16); a.w = __uint_as_float(r.y & 0xffff0000u);
      b.x = __uint_as_float(r.z << 16); b.y = __uint_as_float(r.z & 0xffff0000u);
      b.z = __uint_as_float(r.w << 16); b.w = __uint_as_float(r.w & 0xffff0000u);
    }
    st_stream(dst + 2 * i, a);
    st_stream(dst + 2 * i + 1, b);
  }
}

}  // namespace msda
