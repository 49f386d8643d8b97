// Exact top-k threshold of a register-resident row (k-th largest value, duplicates counted), built for a
// bandwidth-bound kernel: no per-element shared-memory atomics, a handful of ALU ops per element per pass.
//
//   1. One statistics pass gives finite-count / mean / variance / min / max of the row.
//   2. A monotone 16-field classifier f(v) = clamp(round(v*scale + bias), 0, 15) splits the row into value-ordered
//      fields; every thread counts its elements per field in nibble-packed registers, the block reduces the 16
//      counts (REDUX + one shared-memory hop) and picks the field F holding the k-th largest element.
//      The first classifier brackets the moment estimate mean + z_k*sd with 14 narrow interior fields, so F
//      normally holds a few hundred candidates after ONE pass; otherwise the range is reset to the exact
//      [min, max] of field F and the pass repeats (any monotone classifier keeps the search exact).
//   3. The <= kListMax candidates of field F are compacted into shared memory and ranked exactly.
//   4. If the statistics are not finite or the refinement does not converge, an MSB radix select over the
//      ordered integer keys (shared-memory histograms) finishes the job; it is exact for every input.
#pragma once

#include "common.cuh"

namespace lantern {

constexpr int kListMax = 256;   // <= the smallest block size that calls select_kth_largest
constexpr int kSelMaxIters = 8;

struct SelectSmem {
  unsigned warp_cnt[32][8];   // per-warp field counts, two 16-bit fields per word
  unsigned total[16];
  float list[kListMax];
  int rank_gt[kListMax];
  int rank_ge[kListMax];
  float f4[4][33];            // float reductions
  unsigned hist[258];         // radix fallback
  int i_scr[8];
  float f_scr[8];
};

// MSB-first radix select on ordered keys (exact for any input); returns the k-th largest key.
template <int NE>
__device__ __noinline__ uint32_t radix_select_kth(const float (&s)[NE], int k, unsigned* hist /*[258]*/) {
  uint32_t prefix = 0, mask = 0;
  int krem = k;
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const uint32_t key = float_key(s[e]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      unsigned c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[tid * 8 + j]; tot += c[j]; }
      unsigned above = 0;
      for (int l = 31; l > 0; --l) {
        const unsigned t = __shfl_sync(0xffffffffu, tot, l);
        if (tid < l) above += t;
      }
      unsigned run = above;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (run < (unsigned)krem && run + c[j] >= (unsigned)krem) {
          hist[256] = tid * 8 + j;
          hist[257] = run;
        }
        run += c[j];
      }
    }
    __syncthreads();
    prefix |= hist[256] << shift;
    mask |= 0xffu << shift;
    krem -= (int)hist[257];
    __syncthreads();
  }
  return prefix;
}

struct Classifier {
  float scale, bias23;   // field = clamp(round(v * scale + bias), 0, 15), computed through the 2^23 trick
  __device__ __forceinline__ unsigned operator()(float v) const {
    float t = fmaf(v, scale, bias23);
    t = fminf(fmaxf(t, 8388608.0f), 8388623.0f);
    return __float_as_uint(t) & 15u;
  }
};

// Interior fields 1..14 tile [lo, hi]; field 0 collects v < lo, field 15 collects v > hi.
__device__ __forceinline__ Classifier make_classifier(float lo, float hi) {
  Classifier c;
  c.scale = 13.0f / (hi - lo);
  c.bias23 = fmaf(-lo, c.scale, 1.0f) + 8388608.0f;
  return c;
}

// Block-wide count of the 16 fields.  On return sm.i_scr[1..3] = {field F holding the k-th largest element,
// number of elements in higher fields, number of elements in F}.
template <int NE>
__device__ __forceinline__ void count_fields(const float (&s)[NE], const Classifier& cls, int k, SelectSmem& sm) {
  static_assert(NE <= 255, "byte accumulators");
  // nibble-packed counters, flushed into byte accumulators every 15 elements
  unsigned long long nib = 0ull, acc_e = 0ull, acc_o = 0ull;
  constexpr unsigned long long kEven = 0x0F0F0F0F0F0F0F0Full;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    nib += 1ull << (cls(s[e]) * 4u);
    if ((e % 15) == 14 || e == NE - 1) {
      acc_e += nib & kEven;            // fields 0,2,..,14 as bytes
      acc_o += (nib >> 4) & kEven;     // fields 1,3,..,15 as bytes
      nib = 0ull;
    }
  }
  // widen bytes to 16-bit pairs: word w holds (field a | field b << 16)
  const unsigned e_lo = (unsigned)acc_e, e_hi = (unsigned)(acc_e >> 32);
  const unsigned o_lo = (unsigned)acc_o, o_hi = (unsigned)(acc_o >> 32);
  unsigned w[8];
  // byte i of acc_e = field 2i ; byte i of acc_o = field 2i+1.  word j (j<8) = field 2j | field 2j+1 << 16
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned ev = ((j < 4 ? e_lo : e_hi) >> ((j & 3) * 8)) & 0xffu;
    const unsigned od = ((j < 4 ? o_lo : o_hi) >> ((j & 3) * 8)) & 0xffu;
    w[j] = ev | (od << 16);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = __reduce_add_sync(0xffffffffu, w[j]);   // 32 lanes * 255 < 65536: no carry
  __syncthreads();   // previous readers of sm.total / warp_cnt are done
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm.warp_cnt[warp][j] = w[j];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int f = threadIdx.x & 15;
    unsigned tot = 0;
    for (int wi = 0; wi < nwarp; ++wi) tot += (sm.warp_cnt[wi][f >> 1] >> ((f & 1) * 16)) & 0xffffu;
    // suffix sums over the 16 fields (lanes 16..31 mirror lanes 0..15): above = sum of counts of higher fields
    unsigned incl = tot;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const unsigned n = __shfl_down_sync(0xffffffffu, incl, o, 16);
      if (f + o < 16) incl += n;
    }
    const unsigned above = incl - tot;
    if (threadIdx.x < 16 && above < (unsigned)k && above + tot >= (unsigned)k) {
      sm.i_scr[1] = f; sm.i_scr[2] = (int)above; sm.i_scr[3] = (int)tot;
    }
  }
  __syncthreads();
}

// k-th largest of the row held in s[] (NE per thread, padded slots = -inf), 1 <= k <= number of slots.
// `z_guess` = inverse normal CDF of (1 - k/n), `win_sd` = half-width of the first bracket in standard deviations.
template <int NE>
__device__ __forceinline__ float select_kth_largest(const float (&s)[NE], int k, float row_min, float row_max,
                                                    float mean, float sd, float z_guess, float win_sd,
                                                    SelectSmem& sm) {
  const int tid = threadIdx.x;
  bool ok = isfinite(row_min) && isfinite(row_max) && isfinite(mean) && isfinite(sd);
  float lo = fmaxf(mean + (z_guess - win_sd) * sd, row_min);
  float hi = fminf(mean + (z_guess + win_sd) * sd, row_max);
  if (!(lo < hi)) { lo = row_min; hi = row_max; }
  float result = 0.f;
  bool done = false;
  if (ok && row_min == row_max) { result = row_max; done = true; }   // constant row (padding excluded by caller's k)
#pragma unroll 1
  for (int it = 0; ok && !done && it < kSelMaxIters; ++it) {
    const Classifier cls = make_classifier(lo, hi);
    if (!isfinite(cls.scale) || !isfinite(cls.bias23)) { ok = false; break; }
    count_fields<NE>(s, cls, k, sm);
    const int F = sm.i_scr[1];
    const unsigned above = (unsigned)sm.i_scr[2], cntF = (unsigned)sm.i_scr[3];
    const int krem = k - (int)above;   // rank inside field F (1-based from the top)
    if (cntF <= (unsigned)kListMax) {
      // ---- compact field F into shared memory and rank it exactly ----
      unsigned hit = 0u;
#pragma unroll
      for (int e = 0; e < NE; ++e) hit |= (cls(s[e]) == (unsigned)F ? 1u : 0u) << (e & 31);
      if (NE > 32) {   // (not instantiated today; keeps the bitmask honest)
        hit = 0xffffffffu;
      }
      __syncthreads();
      if (tid == 0) sm.i_scr[0] = 0;
      __syncthreads();
      if (hit) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (cls(s[e]) == (unsigned)F) sm.list[atomicAdd(&sm.i_scr[0], 1)] = s[e];
        }
      }
      const int m = (int)cntF;
      int mp = 32;
      while (mp < m) mp <<= 1;                 // power of two, <= kListMax <= blockDim.x
      const int nparts = (int)blockDim.x / mp;  // threads sharing one candidate
      const int ci = tid & (mp - 1), part = tid / mp;
      for (int j = tid; j < kListMax; j += blockDim.x) { sm.rank_gt[j] = 0; sm.rank_ge[j] = 0; }
      __syncthreads();
      if (ci < m) {
        const float vi = sm.list[ci];
        int gt = 0, ge = 0;
        for (int j = part; j < m; j += nparts) {   // warp-uniform j: shared-memory broadcast
          const float vj = sm.list[j];
          gt += vj > vi;
          ge += vj >= vi;
        }
        atomicAdd(&sm.rank_gt[ci], gt);
        atomicAdd(&sm.rank_ge[ci], ge);
      }
      __syncthreads();
      if (tid < m && sm.rank_gt[tid] < krem && krem <= sm.rank_ge[tid]) sm.f_scr[0] = sm.list[tid];  // same value from all writers
      __syncthreads();
      result = sm.f_scr[0];
      done = true;
    } else {
      // ---- too many candidates: shrink the range to the exact [min, max] of field F and classify again ----
      float mn = INFINITY, mxv = -INFINITY;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (cls(s[e]) == (unsigned)F) { mn = fminf(mn, s[e]); mxv = fmaxf(mxv, s[e]); }
      }
      mn = -block_reduce(-mn, OpMaxF(), -INFINITY, sm.f4[0]);
      mxv = block_reduce(mxv, OpMaxF(), -INFINITY, sm.f4[1]);
      if (mn == mxv) { result = mn; done = true; }
      else if (!isfinite(mn) || !isfinite(mxv)) { ok = false; }
      else { lo = mn; hi = mxv; }
    }
  }
  if (!done) result = key_float(radix_select_kth<NE>(s, k, sm.hist));
  return result;
}

}  // namespace lantern
