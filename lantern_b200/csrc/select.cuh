// Exact top-k threshold of a register-resident row (k-th largest value, duplicates counted), built for a
// bandwidth-bound kernel: a handful of ALU ops per element, no per-element shared-memory atomics.
//
// Tier 1 (bracket_select) - the common case, one pass over the registers:
//   * the caller's row statistics give the moment estimate mean + z_k*sd of the k-th largest value; a bracket
//     [lo, hi] of +-win_sd standard deviations is placed around it;
//   * every thread counts its elements above hi and parks its elements inside the bracket (a few per cent of
//     the row) in a thread-private, conflict-free shared-memory column;
//   * if the bracket really contains the k-th largest element (checked with the exact counts), the parked
//     elements are histogrammed into 64 value-ordered fields (a few hundred shared-memory atomics per row), the
//     field F holding the target is found with a warp scan, its handful of members is compacted and ranked.
// Tier 2 (select_kth_largest) - bracket missed (non-Gaussian rows, heavy ties): 16-field classification of all
//   register-resident elements with nibble-packed counters, range refinement on the exact [min, max] of the
//   selected field until it holds <= kListMax candidates.  Any monotone classifier keeps the search exact.
// Tier 3 - statistics not finite or no convergence: MSB radix select over ordered integer keys (caller).
#pragma once

#include "common.cuh"

namespace lantern {

constexpr int kListMax = 256;   // <= the smallest block size that calls select_kth_largest
constexpr int kSelMaxIters = 8;

struct SelectSmem {
  unsigned warp_cnt[32][8];   // per-warp field counts, two 16-bit fields per word
  unsigned total[16];
  float list[kListMax];
  float f4[4][33];            // float reductions
  unsigned hist[258];         // radix fallback
  unsigned hist64[64];        // tier-1 fine histogram
  int i_scr[8];
  float f_scr[8];
};

// MSB-first radix select on ordered keys (exact for any input); returns the k-th largest key.
template <int NE, class Bar = BlockBar>
__device__ __forceinline__ uint32_t radix_select_kth(const float (&s)[NE], int k, unsigned* hist /*[258]*/) {
  uint32_t prefix = 0, mask = 0;
  int krem = k;
  const int tid = Bar::tid();
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += Bar::size()) hist[i] = 0;
    Bar::sync();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const uint32_t key = float_key(s[e]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    Bar::sync();
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
    Bar::sync();
    prefix |= hist[256] << shift;
    mask |= 0xffu << shift;
    krem -= (int)hist[257];
    Bar::sync();
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
template <int NE, class Bar = BlockBar>
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
  const int lane = Bar::tid() & 31, warp = Bar::tid() >> 5, nwarp = Bar::size() >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = __reduce_add_sync(0xffffffffu, w[j]);   // 32 lanes * 255 < 65536: no carry
  Bar::sync();   // previous readers of sm.total / warp_cnt are done
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm.warp_cnt[warp][j] = w[j];
  }
  Bar::sync();
  if (Bar::tid() < 32) {
    const int f = Bar::tid() & 15;
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
    if (Bar::tid() < 16 && above < (unsigned)k && above + tot >= (unsigned)k) {
      sm.i_scr[1] = f; sm.i_scr[2] = (int)above; sm.i_scr[3] = (int)tot;
    }
  }
  Bar::sync();
}

// Exact krem-th largest (1-based) of sm.list[0..m): one warp per candidate, lanes split the comparisons.
template <class Bar = BlockBar>
__device__ __forceinline__ float rank_list(int m, int krem, SelectSmem& sm) {
  const int lane = Bar::tid() & 31, warp = Bar::tid() >> 5, nwarp = Bar::size() >> 5;
  for (int i = warp; i < m; i += nwarp) {
    const float vi = sm.list[i];
    int gt = 0, ge = 0;
    for (int j = lane; j < m; j += 32) {
      const float vj = sm.list[j];
      gt += vj > vi;
      ge += vj >= vi;
    }
    gt = __reduce_add_sync(0xffffffffu, gt);
    ge = __reduce_add_sync(0xffffffffu, ge);
    if (lane == 0 && gt < krem && krem <= ge) sm.f_scr[0] = vi;   // every qualifying candidate has the same value
  }
  Bar::sync();
  return sm.f_scr[0];
}

struct Classifier64 {
  float scale, bias23;   // field = clamp(round((v - lo) * 61 / (hi - lo)) + 1, 0, 63)
  __device__ __forceinline__ unsigned operator()(float v) const {
    float t = fmaf(v, scale, bias23);
    t = fminf(fmaxf(t, 8388608.0f), 8388671.0f);
    return __float_as_uint(t) & 63u;
  }
};
__device__ __forceinline__ Classifier64 make_classifier64(float lo, float hi) {
  Classifier64 c;
  c.scale = 61.0f / (hi - lo);
  c.bias23 = fmaf(-lo, c.scale, 1.0f) + 8388608.0f;
  return c;
}

// Tier 1, second half: everything after the parking pass.  Every thread of the group passes its count of elements
// above the bracket and the number of elements it parked in its column of `buf` ([slots][NT], column Bar::tid()).
// Returns true and sets *result when the k-th largest value was found.
template <int NT, class Bar = BlockBar>
__device__ __forceinline__ bool bracket_finish(int above, int slot, int k, float lo, float hi, float* buf,
                                               SelectSmem& sm, float* result) {
  const int tid = Bar::tid(), lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  // group totals of `above` and `slot`
  const int wa = __reduce_add_sync(0xffffffffu, above), wi = __reduce_add_sync(0xffffffffu, slot);
  Bar::sync();
  if (lane == 0) { sm.warp_cnt[warp][0] = (unsigned)wa; sm.warp_cnt[warp][1] = (unsigned)wi; }
  if (tid < 64) sm.hist64[tid] = 0u;
  Bar::sync();
  int tot_above = 0, tot_in = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) { tot_above += (int)sm.warp_cnt[w][0]; tot_in += (int)sm.warp_cnt[w][1]; }
  if (!(tot_above < k && k <= tot_above + tot_in)) return false;   // the bracket missed the target
  const int krem = k - tot_above;   // rank among the parked elements
#pragma unroll 1
  for (int it = 0; it < kSelMaxIters; ++it) {
    const Classifier64 cls = make_classifier64(lo, hi);
    if (!isfinite(cls.scale) || !isfinite(cls.bias23)) return false;
    for (int i = 0; i < slot; ++i) atomicAdd(&sm.hist64[cls(buf[i * NT + tid])], 1u);
    Bar::sync();
    if (warp == 0) {
      // lane l owns fields 2l, 2l+1; suffix sums from the top field down
      const unsigned c0 = sm.hist64[2 * lane], c1 = sm.hist64[2 * lane + 1];
      unsigned incl = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += n;
      }
      const unsigned above_pair = incl - (c0 + c1);     // elements in fields > 2l+1
      if (above_pair < (unsigned)krem && above_pair + c1 >= (unsigned)krem) {
        sm.i_scr[1] = 2 * lane + 1; sm.i_scr[2] = (int)above_pair; sm.i_scr[3] = (int)c1;
      } else if (above_pair + c1 < (unsigned)krem && above_pair + c1 + c0 >= (unsigned)krem) {
        sm.i_scr[1] = 2 * lane; sm.i_scr[2] = (int)(above_pair + c1); sm.i_scr[3] = (int)c0;
      }
      if (lane == 0) sm.i_scr[0] = 0;
    }
    Bar::sync();
    const unsigned F = (unsigned)sm.i_scr[1];
    const int above2 = sm.i_scr[2], cntF = sm.i_scr[3];
    if (cntF <= kListMax) {
      for (int i = 0; i < slot; ++i) {
        const float v = buf[i * NT + tid];
        if (cls(v) == F) sm.list[atomicAdd(&sm.i_scr[0], 1)] = v;
      }
      Bar::sync();
      *result = rank_list<Bar>(cntF, krem - above2, sm);
      return true;
    }
    // too many candidates (ties / dense bracket): shrink to the exact [min, max] of field F and repeat
    float mn = INFINITY, mxv = -INFINITY;
    for (int i = 0; i < slot; ++i) {
      const float v = buf[i * NT + tid];
      if (cls(v) == F) { mn = fminf(mn, v); mxv = fmaxf(mxv, v); }
    }
    mn = -group_reduce<Bar>(-mn, OpMaxF(), -INFINITY, sm.f4[0]);
    mxv = group_reduce<Bar>(mxv, OpMaxF(), -INFINITY, sm.f4[1]);
    if (mn == mxv) { *result = mn; return true; }
    lo = mn; hi = mxv;
    if (tid < 64) sm.hist64[tid] = 0u;
    Bar::sync();
  }
  return false;
}

// Tier 1.  `buf` is a [NE][NT] float array in shared memory (column tid is private to the thread).
// Returns true and sets *result when the k-th largest value was found.
template <int NE, int NT, class Bar = BlockBar>
__device__ __forceinline__ bool bracket_select(const float (&s)[NE], int k, float lo, float hi, float* buf,
                                               SelectSmem& sm, float* result) {
  const int tid = Bar::tid();
  int above = 0, slot = 0;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const float v = s[e];
    const bool ab = v > hi;
    above += ab;
    if (!ab && v >= lo) {
      buf[slot * NT + tid] = v;
      ++slot;
    }
  }
  return bracket_finish<NT, Bar>(above, slot, k, lo, hi, buf, sm, result);
}

// k-th largest of the row held in s[] (NE per thread, padded slots = -inf), 1 <= k <= number of slots.
// Tier 2: starts from the full [row_min, row_max] range.  Returns false when it cannot converge (tier 3 needed).
template <int NE, class Bar = BlockBar>
__device__ __forceinline__ bool select_kth_largest(const float (&s)[NE], int k, float row_min, float row_max,
                                                SelectSmem& sm, float* out) {
  const int tid = Bar::tid();
  bool ok = isfinite(row_min) && isfinite(row_max);
  float lo = row_min, hi = row_max;
  float result = 0.f;
  bool done = false;
  if (ok && row_min == row_max) { result = row_max; done = true; }   // constant row (padding excluded by caller's k)
#pragma unroll 1
  for (int it = 0; ok && !done && it < kSelMaxIters; ++it) {
    const Classifier cls = make_classifier(lo, hi);
    if (!isfinite(cls.scale) || !isfinite(cls.bias23)) { ok = false; break; }
    count_fields<NE, Bar>(s, cls, k, sm);
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
      Bar::sync();
      if (tid == 0) sm.i_scr[0] = 0;
      Bar::sync();
      if (hit) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (cls(s[e]) == (unsigned)F) sm.list[atomicAdd(&sm.i_scr[0], 1)] = s[e];
        }
      }
      Bar::sync();
      result = rank_list<Bar>((int)cntF, krem, sm);
      done = true;
    } else {
      // ---- too many candidates: shrink the range to the exact [min, max] of field F and classify again ----
      float mn = INFINITY, mxv = -INFINITY;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (cls(s[e]) == (unsigned)F) { mn = fminf(mn, s[e]); mxv = fmaxf(mxv, s[e]); }
      }
      mn = -group_reduce<Bar>(-mn, OpMaxF(), -INFINITY, sm.f4[0]);
      mxv = group_reduce<Bar>(mxv, OpMaxF(), -INFINITY, sm.f4[1]);
      if (mn == mxv) { result = mn; done = true; }
      else if (!isfinite(mn) || !isfinite(mxv)) { ok = false; }
      else { lo = mn; hi = mxv; }
    }
  }
  *out = result;
  return done;
}

// Tiers 2 + 3 behind one non-inlined call (rare path; works on the caller's copy of the row).
template <int NE, class Bar = BlockBar>
__device__ __noinline__ float select_slow(const float (&s)[NE], int k, float row_min, float row_max, SelectSmem& sm) {
  float r;
  if (select_kth_largest<NE, Bar>(s, k, row_min, row_max, sm, &r)) return r;
  Bar::sync();
  return key_float(radix_select_kth<NE, Bar>(s, k, sm.hist));
}

}  // namespace lantern
