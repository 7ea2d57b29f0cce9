// Forward for SMALL pyramids (L = P = 4, fp32 value, S <= 1600 tokens: ARCTIC's 28^2 + 14^2 + 7^2 + 4^2 = 1045,
// /root/reference/util/settings.py:102-120 with 224 x 224 windows): one head of one sample -- S rows of 128 B, 131 KB at
// ARCTIC's shape -- is staged in shared memory ONCE and every corner row is then gathered with LDS.128 instead of LDG.128.
//
// Why: a gather instruction that fetches 4 whole rows costs the SM 3.6 cycles from shared memory but 6.9 when any of the
// rows comes from L2 (experiments/l1_gather_bench.cu); the rows of a (sample, head) are re-read ~64 x by
// the sample's queries, and the staging itself goes through cp.async at the L2 -> SM fill rate once per slab.
//
// Same arithmetic as fwd_d32_kernel (msda_d32.cuh) item for item, in the same order: bit-identical output.
//
// Work: the call is N * M slabs x ceil(Lq / 96) CTA passes (96 queries of one head: 24 warps x one query per 8-lane
// group).  The grid is one persistent CTA per SM (the slab leaves room for one); CTA c takes the c-th equal share of the
// GLOBAL pass sequence, so the SMs finish together whatever N * M is, and re-stages a slab only when its share crosses
// into the next (sample, head) -- at most 1 + ceil(share / passes per slab) times.  Inside a slab the warps run
// free of each other (private records, no block barrier per pass).
#pragma once

#include "msda_d32.cuh"
#include "msda_d32_tiles.cuh"   // cp_async_16, cp_async_wait_all

namespace msda {

#ifndef MSDA_SLAB_WARPS
#define MSDA_SLAB_WARPS 24   // 16 / 24 / 32 measured 90 / 88 / 90 us at ARCTIC's encoder shape (fused prologue: 135 / 121 / 121)
#endif
constexpr int kSlabWarps = MSDA_SLAB_WARPS;
constexpr int kSlabThreads = kSlabWarps * kWarp;
constexpr int kSlabLP = 16, kSlabL = 4, kSlabP = 4;
constexpr int kSlabPassQ = kSlabWarps * kItemsPerWarp;   // 96 queries per CTA pass

struct SlabShared {
  LevelTab tab;
  int4 recs[kSlabWarps][kItemsPerWarp * rec_stride(kSlabLP)];
};
// fp32 rows that fit beside the records in the 227 KB a CTA may have (1600 with 16 warps)
constexpr int kSlabRoom = (227 * 1024 - (int)sizeof(SlabShared) - 128) / 128;
constexpr int kSlabMaxRows = kSlabRoom < 1600 ? kSlabRoom : 1600;

inline size_t slab_smem_bytes(int S) { return sizeof(SlabShared) + 128 + (size_t)S * 128; }

template <int kM, bool kFuse = false>
__global__ void __launch_bounds__(kSlabThreads, 1)
fwd_d32_slab_kernel(const D32Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlabShared& sh = *reinterpret_cast<SlabShared*>(smem_raw);
  // slab rows start on a 128 B boundary behind the records
  const unsigned past = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)sizeof(SlabShared);
  float4* slab = reinterpret_cast<float4*>(smem_raw + sizeof(SlabShared) + ((128u - (past & 127u)) & 127u));
  const int t = threadIdx.x;
  const int M = kM > 0 ? kM : p.M;
  if (t < kSlabL) {
    sh.tab.H[t] = (int)p.shapes[2 * t];
    sh.tab.W[t] = (int)p.shapes[2 * t + 1];
    sh.tab.start[t] = (int)p.lsi[t];
  }
  __syncthreads();

  const int warp = t >> 5, lane = t & 31;
  const int group = lane >> 3, sub = lane & 7;
  constexpr int stride = rec_stride(kSlabLP);
  int4* warp_recs = sh.recs[warp];
  const int4* my_recs = warp_recs + group * stride;
  int wrow[kSlabL];
#pragma unroll
  for (int l = 0; l < kSlabL; ++l) wrow[l] = sh.tab.W[l] * 8;   // one row down, in 16 B chunks

  const int ppq = (p.Lq + kSlabPassQ - 1) / kSlabPassQ;            // CTA passes per slab
  const long long n_pass = (long long)p.N * M * ppq;                // < 2^31: host-checked
  const int g0 = (int)(n_pass * blockIdx.x / gridDim.x), g1 = (int)(n_pass * (blockIdx.x + 1) / gridDim.x);

  // this lane's two samples of a pass: queries (lane >> 4) and (lane >> 4) + 2 of the warp's four, sample lane & 15
  const int ls = lane & 15, lvl = ls >> 2;
  const int Hl = sh.tab.H[lvl], Wl = sh.tab.W[lvl], startl = sh.tab.start[lvl];
  float2 xy_a, xy_b;
  float2 rf_a = make_float2(0.f, 0.f), rf_b = rf_a;   // fused prologue: the samples' reference points
  float a_a, a_b;
  int qa = -1, qb = -1;
  auto request = [&](int gp) {       // global pass gp -> this lane's raw samples
    const int bm = gp / ppq, q0 = (gp - bm * ppq) * kSlabPassQ + warp * kItemsPerWarp + (lane >> 4);
    qa = q0 < p.Lq ? q0 : -1;
    qb = q0 + 2 < p.Lq ? q0 + 2 : -1;
    const int b = bm / M, m = bm - b * M;
    const unsigned base = (((unsigned)b * (unsigned)p.Lq) * (unsigned)M + (unsigned)m) * kSlabLP + ls;
    if (qa >= 0) {
      const unsigned gs = base + (unsigned)qa * (unsigned)M * kSlabLP;
      xy_a = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
      a_a = ld_stream(p.aw + gs);
      if (kFuse) rf_a = __ldg(reinterpret_cast<const float2*>(p.ref) + ((size_t)b * p.Lq + qa) * kSlabL + lvl);
    }
    if (qb >= 0) {
      const unsigned gs = base + (unsigned)qb * (unsigned)M * kSlabLP;
      xy_b = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
      a_b = ld_stream(p.aw + gs);
      if (kFuse) rf_b = __ldg(reinterpret_cast<const float2*>(p.ref) + ((size_t)b * p.Lq + qb) * kSlabL + lvl);
    }
  };

  int staged = -1;
  if (g0 < g1) request(g0);
  for (int gp = g0; gp < g1; ++gp) {
    const int bm = gp / ppq;
    const int b = bm / M, m = bm - b * M;
    if (bm != staged) {
      // ---- stage the slab: row r of head m of sample b -> slab[r * 8 .. r * 8 + 7] ---------------------------------
      __syncthreads();                                   // every warp is done with the previous slab
      const float4* src = reinterpret_cast<const float4*>(p.value) + ((size_t)b * p.S * M + m) * 8;
      for (int i = t; i < p.S * 8; i += kSlabThreads) cp_async_16(slab + i, src + (size_t)(i >> 3) * M * 8 + (i & 7));
      cp_async_wait_all();
      __syncthreads();
      staged = bm;
    }
    // ---- phase A: park this pass's records (token rows: M = 1), request the next pass's samples --------------------
    {
      float2 la = xy_a, lb = xy_b;
      float wa = a_a, wb = a_b;
      if (kFuse) {
        wa = segment_softmax<kSlabLP>(qa >= 0 ? wa : -INFINITY);
        wb = segment_softmax<kSlabLP>(qb >= 0 ? wb : -INFINITY);
        if (qa >= 0) la = fused_location(rf_a, la, Hl, Wl);
        if (qb >= 0) lb = fused_location(rf_b, lb, Hl, Wl);
      }
      if (qa >= 0) warp_recs[(lane >> 4) * stride + ls] = make_record<false, false>(la.x, la.y, wa, Hl, Wl, startl, 1);
      if (qb >= 0) warp_recs[((lane >> 4) + 2) * stride + ls] = make_record<false, false>(lb.x, lb.y, wb, Hl, Wl, startl, 1);
    }
    const int q = (gp - bm * ppq) * kSlabPassQ + warp * kItemsPerWarp + group;
    if (gp + 1 < g1) request(gp + 1);
    __syncwarp();

    // ---- phase B: gather from the slab (fwd_d32_kernel's arithmetic, operation for operation) ----------------------
    if (q < p.Lq) {
      const float4* vb = slab + sub;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int l = 0; l < kSlabL; ++l) {
#pragma unroll
        for (int s = l * kSlabP; s < (l + 1) * kSlabP; ++s) {
          const int4 r = my_recs[s];
          const int row = r.x >> 4;                      // token of corner (h_low, w_low); may be negative at the border
          const float ah = __int_as_float(r.y), al = __int_as_float(r.z), lw = __int_as_float(r.w);
          const float w2 = ah * lw, w4 = al * lw;
          const float w1 = ah - w2, w3 = al - w4;
          const float4* r0 = vb + row * 8;
          const float4* r1 = r0 + wrow[l];
          if (r.x & 1) fma4(acc, w1, r0[0]);
          if (r.x & 2) fma4(acc, w2, r0[8]);
          if (r.x & 4) fma4(acc, w3, r1[0]);
          if (r.x & 8) fma4(acc, w4, r1[8]);
        }
      }
      const long long item = ((long long)b * p.Lq + q) * M + m;
      st_stream(reinterpret_cast<float4*>(p.out) + item * 8 + sub, acc);
    }
    __syncwarp();   // records are rewritten by the next pass
  }
}

}  // namespace msda
