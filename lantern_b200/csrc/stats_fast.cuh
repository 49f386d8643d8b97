// Row-statistics kernel, fast path (the HBM-bound kernel of the fused verify step).
//
// Persistent CTAs, one logits row per iteration:
//   * thread 0 issues TMA bulk copies (cp.async.bulk + mbarrier) of the cond / uncond windows of the CTA's NEXT
//     row as soon as the current row has been lifted into registers, so the DMA engine streams HBM while the SM
//     does the arithmetic of the current row;
//   * the row lives in registers (NE values per thread); CFG mix + temperature are applied while lifting;
//   * every block-wide exchange is "warp partials -> shared memory -> one barrier -> every warp combines the
//     partials itself": five barriers per row (statistics, bracket counts + histogram, compaction, rank, sum);
//   * the exact top-k threshold comes from the tier-1 bracket select of select.cuh restated in that style; if the
//     bracket misses (or the row is not finite) the tier-2/3 selectors of select.cuh finish the row.
//
// Preconditions (checked by the launcher): ncols == 4*NT*NQ, 4-element alignment of the window, room for a
// 16-byte aligned copy of the window.
#pragma once

#include "accept_types.cuh"
#include "select.cuh"

#ifndef LANTERN_EXP
#define LANTERN_EXP 0
#endif

namespace lantern {

template <int NW>
struct FastSmem {
  float4 st_part[NW];        // per-warp sum, sum of squares, min, max
  unsigned cnt_part[NW];     // per-warp (elements above the bracket) | (elements inside it) << 16
  float sum_part[NW];        // per-warp softmax partial sums
  unsigned hist[64];
  float list[kListMax];
  int list_n;
  float kth;
  alignas(8) uint64_t mbar;
};

template <int DT, int NT, int NQ, int MODE>   // MODE 1: cond + uncond, MODE 2: cond only
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) row_stats_fast_kernel(const AcceptParams P) {
  constexpr int NE = NQ * 4, NW = NT / 32, EB = Elem<DT>::kBytes;
  __shared__ FastSmem<NW> fs;
  __shared__ SelectSmem sm;   // slow path only
  __shared__ float slow_scr[33];
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  const lantern_accept_cfg& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_rows_total = cfg.n_items * cfg.n_rows;
  const int stage_bytes = (cfg.ncols * EB + 32 + 127) & ~127;
  unsigned char* buf_c = dyn_smem;
  unsigned char* buf_u = dyn_smem + stage_bytes;
  float* park = reinterpret_cast<float*>(dyn_smem + (MODE == 1 ? 2 : 1) * stage_bytes);
  MixParams mix = P.mix;
  mix.has_uncond = MODE == 1;

  auto base_of = [&](int item, int t) -> int64_t {
    return (int64_t)item * cfg.item_stride + (int64_t)t * cfg.row_stride + cfg.col0;
  };
  auto issue_row = [&](int64_t rb) {
    const uintptr_t gc = reinterpret_cast<uintptr_t>(P.in.logits_cond) + (uintptr_t)rb * EB;
    const uintptr_t ac = gc & ~uintptr_t(15);
    const uint32_t bc = (uint32_t)((gc - ac) + (uintptr_t)cfg.ncols * EB + 15) & ~15u;
    uint32_t bu = 0;
    uintptr_t au = 0;
    if (MODE == 1) {
      const uintptr_t gu = reinterpret_cast<uintptr_t>(P.in.logits_uncond) + (uintptr_t)rb * EB;
      au = gu & ~uintptr_t(15);
      bu = (uint32_t)((gu - au) + (uintptr_t)cfg.ncols * EB + 15) & ~15u;
    }
    mbar_expect_tx(&fs.mbar, bc + bu);
    bulk_g2s(buf_c, reinterpret_cast<const void*>(ac), bc, &fs.mbar);
    if (MODE == 1) bulk_g2s(buf_u, reinterpret_cast<const void*>(au), bu, &fs.mbar);
  };

  // Bracket centre in standard deviations: starts at the Gaussian quantile and then tracks the quantile observed on
  // the CTA's previous row (rows of one model share their shape), so a narrow bracket keeps hitting on
  // non-Gaussian logits; any miss is still resolved exactly by tiers 2/3.
  float z_run = P.z_guess, win_run = P.win_sd_first;
  uint32_t parity = 0;
  // (item, t) of the current row and of the next row of this CTA, advanced without divisions
  int item = (int)blockIdx.x / cfg.n_rows, trow = (int)blockIdx.x % cfg.n_rows;
  int n_item = item, n_trow = trow;
  const int step_item = (int)gridDim.x / cfg.n_rows, step_row = (int)gridDim.x % cfg.n_rows;
  auto advance = [&](int& it, int& tr) {
    tr += step_row;
    it += step_item;
    if (tr >= cfg.n_rows) { tr -= cfg.n_rows; ++it; }
  };
  advance(n_item, n_trow);
  if (tid == 0) {
    mbar_init(&fs.mbar, 1);
    if ((int)blockIdx.x < n_rows_total) issue_row(base_of(item, trow));
  }
  __syncthreads();

  for (int row = blockIdx.x; row < n_rows_total; row += gridDim.x) {
    RowStats st;
    st.thr = -INFINITY; st.mx = 0.f; st.sum = 1.f; st.vcut = -INFINITY; st.icut = -1;
    st.kind = P.in.row_kinds ? (int)P.in.row_kinds[row] : LANTERN_ROW_IMAGE;
    st.pad0 = st.pad1 = 0;
    const int64_t base = base_of(item, trow);
    const int lead_c = (int)((reinterpret_cast<uintptr_t>(P.in.logits_cond) + (uintptr_t)base * EB) & 15);
    const int lead_u = MODE == 1 ? (int)((reinterpret_cast<uintptr_t>(P.in.logits_uncond) + (uintptr_t)base * EB) & 15) : 0;

    // ---- lift the staged row into registers: CFG mix + temperature + per-thread statistics ----
    mbar_wait(&fs.mbar, parity);
    parity ^= 1;
    float s[NE];
    float fsum = 0.f, fsq = 0.f, fmx = -INFINITY;   // the minimum is only needed by the rare slow path: computed there
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e0 = (q * NT + tid) * 4;
      float c4[4], u4[4] = {0.f, 0.f, 0.f, 0.f};
      lds4<DT>(buf_c + lead_c, e0, c4);
      if (MODE == 1) lds4<DT>(buf_u + lead_u, e0, u4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = mix_temper(c4[j], u4[j], mix);
        s[q * 4 + j] = v;
        fsum += v;
        fsq = fmaf(v, v, fsq);
        fmx = fmaxf(fmx, v);
      }
    }
    fsum = warp_reduce(fsum, OpSum());
    fsq = warp_reduce(fsq, OpSum());
    fmx = warp_reduce(fmx, OpMaxF());
    if (lane == 0) fs.st_part[warp] = make_float4(fsum, fsq, 0.f, fmx);
    if (tid < 64) fs.hist[tid] = 0u;
    if (tid == 0) fs.list_n = 0;
    __syncthreads();   // B1: statistics partials visible; every thread has consumed the staged row
    if (tid == 0 && row + (int)gridDim.x < n_rows_total) issue_row(base_of(n_item, n_trow));
    item = n_item; trow = n_trow;
    advance(n_item, n_trow);
    if (st.kind != LANTERN_ROW_IMAGE) {   // one-hot rows (Lumina newline / end-of-image): no statistics needed
      if (tid == 0) P.stats[row] = st;
      continue;
    }
    {   // lane l takes warp (l mod NW)'s partial; a butterfly over NW lanes leaves the totals in every lane
      const float4 pw = fs.st_part[lane & (NW - 1)];
      fsum = pw.x; fsq = pw.y; fmx = pw.w;
#pragma unroll
      for (int o = 1; o < NW; o <<= 1) {
        fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
        fsq += __shfl_xor_sync(0xffffffffu, fsq, o);
        fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
      }
    }
    const float m = fmx;

    // ---- exact top-k threshold ----
    float thr = -INFINITY;
    if (P.do_topk) {
      bool found = false, bracket_hit = false;
      // a -inf / +inf / NaN element makes the sum of squares non-finite, a constant row has an empty bracket
      // (sd == 0): both are left to the slow path below
      const bool finite = isfinite(fmx) && isfinite(fsq);
      // moments once per row; they only steer the bracket (exactness comes from the counts), so the reciprocal
      // square root may be the approximate one.  A constant row gives 0 * inf = NaN: empty bracket -> slow path.
      const float mean = fsum * P.inv_ncols;
      const float var = fmaxf(fmaf(fsq, P.inv_ncols, -mean * mean), 0.f);
      const float inv_sd = rsqrtf(var);
      const float sd = var * inv_sd;
      if (finite) {
        const float lo = mean + (z_run - win_run) * sd, hi = mean + (z_run + win_run) * sd;
        if (lo < hi) {
          // bracket pass: count elements above hi, park the elements inside [lo, hi] in the thread's column
          // Branch-free, five instructions per element: two compares, a predicated store through a running shared
          // address (the thread's parking column), its predicated advance, and the predicated count of elements above.
          int above = 0;
          const uint32_t park0 = smem_u32(park + tid);
          uint32_t paddr = park0;
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            asm volatile(
                "{\n"
                ".reg .pred pa, pin;\n"
                "setp.gt.f32 pa, %2, %4;\n"
                "setp.ge.and.f32 pin, %2, %3, !pa;\n"
                "@pin st.shared.f32 [%0], %2;\n"
                "@pin add.u32 %0, %0, %5;\n"
                "@pa add.s32 %1, %1, 1;\n"
                "}\n"
                : "+r"(paddr), "+r"(above)
                : "f"(s[e]), "f"(lo), "f"(hi), "n"(NT * 4)
                : "memory");
          }
          const int slot = (int)((paddr - park0) / (NT * 4));
          const int wa = __reduce_add_sync(0xffffffffu, above), wi = __reduce_add_sync(0xffffffffu, slot);
          if (lane == 0) fs.cnt_part[warp] = (unsigned)wa | ((unsigned)wi << 16);
          // the histogram of the parked elements is filled before the counts are known (it is only wasted on the rare
          // rows whose bracket misses): one barrier covers both exchanges
          Classifier64 cls;   // any monotone classifier keeps the select exact: approximate reciprocal
          cls.scale = __fdividef(61.0f, hi - lo);
          cls.bias23 = fmaf(-lo, cls.scale, 1.0f) + 8388608.0f;
          unsigned fl0 = 0u, fl1 = 0u;   // fields of the first ten parked elements, six bits each
          const int wmax = __reduce_max_sync(0xffffffffu, slot);   // warp-uniform trip count of the unrolled loops
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            if (i >= wmax) break;
            if (i < slot) {
              const unsigned f = cls(park[i * NT + tid]);
              if (i < 5) fl0 |= f << (6 * i); else fl1 |= f << (6 * (i - 5));
              atomicAdd(&fs.hist[f], 1u);
            }
          }
          for (int i = 10; i < slot; ++i) atomicAdd(&fs.hist[cls(park[i * NT + tid])], 1u);
          __syncthreads();   // B2 + B3
          // both 16-bit fields of the packed counts sum without carry (at most ncols <= 32768 elements per row)
          const unsigned csum = __reduce_add_sync(0xffffffffu, lane < NW ? fs.cnt_part[lane] : 0u);
          const int tot_above = (int)(csum & 0xffffu), tot_in = (int)(csum >> 16);
          const int k = cfg.top_k;
          if (tot_above < k && k <= tot_above + tot_in) {
            const int krem = k - tot_above;
            // every warp scans the 64 fields itself: lane l owns fields 2l, 2l+1
            const unsigned c0 = fs.hist[2 * lane], c1 = fs.hist[2 * lane + 1];
            unsigned incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const unsigned n = __shfl_down_sync(0xffffffffu, incl, o);
              if (lane + o < 32) incl += n;
            }
            const unsigned above_pair = incl - (c0 + c1);
            int pick = -1;
            if (above_pair < (unsigned)krem && above_pair + c1 >= (unsigned)krem) pick = 1;
            else if (above_pair + c1 < (unsigned)krem && above_pair + c1 + c0 >= (unsigned)krem) pick = 0;
            const unsigned owner = __ballot_sync(0xffffffffu, pick >= 0);
            const int src = __ffs(owner) - 1;
            const unsigned F = (unsigned)__shfl_sync(0xffffffffu, 2 * lane + (pick > 0 ? 1 : 0), src);
            const int above2 = __shfl_sync(0xffffffffu, (int)(pick == 1 ? above_pair : above_pair + c1), src);
            const int cntF = __shfl_sync(0xffffffffu, (int)(pick == 1 ? c1 : c0), src);
            if (cntF <= kListMax) {
#pragma unroll
              for (int i = 0; i < 10; ++i) {
                if (i >= wmax) break;
                const unsigned f = ((i < 5 ? fl0 >> (6 * i) : fl1 >> (6 * (i - 5))) & 63u);
                if (i < slot && f == F) fs.list[atomicAdd(&fs.list_n, 1)] = park[i * NT + tid];
              }
              for (int i = 10; i < slot; ++i) {
                const float v = park[i * NT + tid];
                if (cls(v) == F) fs.list[atomicAdd(&fs.list_n, 1)] = v;
              }
              __syncthreads();   // B4
              const int kr = krem - above2;
              for (int i = warp; i < cntF; i += NW) {
                const float vi = fs.list[i];
                int gt = 0, ge = 0;
                for (int j = lane; j < cntF; j += 32) {
                  const float vj = fs.list[j];
                  gt += vj > vi;
                  ge += vj >= vi;
                }
                gt = __reduce_add_sync(0xffffffffu, gt);
                ge = __reduce_add_sync(0xffffffffu, ge);
                if (lane == 0 && gt < kr && kr <= ge) fs.kth = vi;   // all qualifying candidates carry the same value
              }
              __syncthreads();   // B5
              thr = fs.kth;
              found = true;
              bracket_hit = true;
            }
          }
        }
      }
      if (!found) {   // tiers 2 and 3 (rare): work on a copy so that s[] stays in registers
        __syncthreads();
        float tmp[NE];
        float fmn = INFINITY;
#pragma unroll
        for (int e = 0; e < NE; ++e) { tmp[e] = s[e]; fmn = fminf(fmn, s[e]); }
        fmn = -block_reduce(-fmn, OpMaxF(), -INFINITY, slow_scr);
        thr = (fmn == fmx) ? fmx : select_slow<NE>(tmp, cfg.top_k, fmn, fmx, sm);
        __syncthreads();
      }
      {   // remember where the quantile really was (in standard deviations) for the next row
        const float z_obs = (thr - mean) * inv_sd;
        // bracket width: the width the sampling noise of the quantile calls for (P.win_sd) after a hit, doubled
        // after every consecutive miss (rows whose shape the Gaussian prior describes badly)
        if (isfinite(z_obs)) {
          z_run = z_obs;
          win_run = bracket_hit ? P.win_sd : fminf(P.win_sd_first, 2.0f * fmaxf(win_run, P.win_sd));
        }
      }
    }

    // ---- softmax sum over the kept columns ----
    const ExpShift ex(m);
    float part = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const float ev = ex(s[e]);
      // predicated add (two instructions) instead of compare + select + add; adding 0.f is the identity, so the sum
      // is bit-identical
      asm("{\n.reg .pred pk;\nsetp.ge.f32 pk, %1, %2;\n@pk add.f32 %0, %0, %3;\n}\n" : "+f"(part) : "f"(s[e]), "f"(thr), "f"(ev));
    }
    part = warp_reduce(part, OpSum());
    if (lane == 0) fs.sum_part[warp] = part;
    __syncthreads();   // B6
    if (tid == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) tot += fs.sum_part[w];
      st.thr = thr; st.mx = m; st.sum = tot;
      P.stats[row] = st;
    }
  }
}

}  // namespace lantern
