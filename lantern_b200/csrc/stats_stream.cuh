// Row-statistics kernel, streaming form (the HBM-bound kernel of the fused verify step; replaces stats_fast.cuh for
// the 2048 ... 16384-column windows).
//
// Persistent CTAs of NT "main" threads plus one "select" warp, one logits row per iteration:
//   main threads
//     * thread 0 issues the TMA bulk copies (cp.async.bulk + mbarrier) of the CTA's NEXT row as soon as the current row
//       has been lifted into registers: the DMA engine streams HBM while the SM works;
//     * pass 1 lifts the staged row into registers (NE values per thread) with the CFG mix (+ temperature) and the
//       moments / maximum of the row; one named barrier exchanges the per-warp partials;
//     * pass 2 (one sweep over the registers, eight instructions per element): exp of every element, sum of the
//       exps above the bracket [lo, hi] placed around the predicted k-th largest value, count of the elements above it,
//       and the few per cent of elements inside the bracket parked in thread-private shared-memory columns; the warp
//       then compacts its parked elements into one contiguous segment and hands the row over (mbarrier arrive);
//   two select warps, one for the CTA's even image rows and one for the odd ones (every hand-over buffer exists
//   twice, so a select warp has two row periods for its row and the main threads practically never wait for it)
//     * exact k-th largest from the compact segments (64-field histogram, warp scan, rank of the handful of
//       candidates), sum of the parked elements that are kept, the 32-byte RowStats record, and the observed quantile
//       that steers the bracket of the row after next (row j uses what row j-2 observed, so the main threads
//       practically never wait for the select warp);
//   rows whose bracket misses (or that are not finite, tie-heavy beyond the candidate list, ...) are flagged and
//   finished after the stream by the main threads with the exact tier-2/3 selectors of select.cuh.
//
// Every choice that affects the result is a function of the CTA's row sequence only (static row striding, the bracket
// of row j comes from row j-2 of the same CTA), so thr / max are exact and the sum is reproducible bit for bit.
//
// Preconditions (checked by the launcher): ncols == 4*NT*NQ <= 16384, 4-element alignment of the window, room for a
// 16-byte aligned copy of the window.
#pragma once

#include "accept_types.cuh"
#include "select.cuh"

namespace lantern {

constexpr int kSegCap = 256;      // parked elements one main warp can hand over per row
constexpr int kCandCap = 256;     // candidates ranked exactly by the select warp

enum { kRowRedo = 1 };
constexpr unsigned kWarpOverflow = 0xffffffffu;   // cnt_part value of a warp whose parked elements did not fit

struct RowDesc {             // main -> select, per row
  float lo, hi, mean, inv_sd, mx;
  int flags;
};

template <int NW>
struct StreamSmem {
  float4 st_part[2][NW];     // pass-1 partials per warp: sum, sum of squares, -, max (double-buffered by row parity)
  // hand-over buffers, indexed by the parity of the CTA's image-row counter
  unsigned cnt_part[2][NW];  // per-warp (elements above the bracket) | (elements inside it) << 16
  float sab_part[2][NW];     // per-warp sum of exp over the elements above the bracket
  unsigned hist[2][64];      // 64-field histogram of the parked elements (filled by the main warps)
  RowDesc desc[2];
  float2 zbuf[2];            // select -> main: (quantile in standard deviations, bracket half-width) for row j + 2
  float cand[2][kCandCap];   // per select warp
  int redo_n;
  alignas(8) uint64_t mbar_tma, mbar_ready[2], mbar_done[2];
};

// mbarrier wait that backs off between polls: a spinning warp would otherwise take issue slots from the warps it
// is waiting for (they share the SM's schedulers)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned sleep_ns = 40) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  while (!done) {
    if (sleep_ns) __nanosleep(sleep_ns);
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

template <int DT, int NT, int NQ, int MODE, bool TEMP>   // MODE 1: cond + uncond, MODE 2: cond only
__global__ void __launch_bounds__(NT + 64, (NT <= 256 ? 2 : 1)) row_stats_stream_kernel(const AcceptParams P) {
  constexpr int NE = NQ * 4, NW = NT / 32, EB = Elem<DT>::kBytes;
  constexpr int PS = NE < 20 ? NE : 20;          // private parking slots per thread before the row is sent to the redo path
  constexpr int LPS = 32 / NW;                   // select-warp lanes per segment
  static_assert(NW <= 32 && (32 % NW) == 0, "one select lane group per main warp");
  constexpr bool CLAMP = NE > PS;                // a thread may park more elements than its column holds: clamp + redo
  using MainBar = NamedBar<1, NT>;
  __shared__ StreamSmem<NW> fs;
  __shared__ SelectSmem sm;     // redo path only
  __shared__ float slow_scr[33];
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  const lantern_accept_cfg& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_rows_total = cfg.n_items * cfg.n_rows;
  const int stage_bytes = (cfg.ncols * EB + 32 + 127) & ~127;
  unsigned char* buf_c = dyn_smem;
  unsigned char* buf_u = dyn_smem + stage_bytes;
  float* park = reinterpret_cast<float*>(dyn_smem + (MODE == 1 ? 2 : 1) * stage_bytes);   // [PS][NT]
  float* seg = park + PS * NT;                                                             // [2][NW][kSegCap]
  MixParams mix = P.mix;
  mix.has_uncond = MODE == 1;
  mix.do_temp = TEMP;

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
    mbar_expect_tx(&fs.mbar_tma, bc + bu);
    bulk_g2s(buf_c, reinterpret_cast<const void*>(ac), bc, &fs.mbar_tma);
    if (MODE == 1) bulk_g2s(buf_u, reinterpret_cast<const void*>(au), bu, &fs.mbar_tma);
  };
  // (item, t) of a row of this CTA, advanced without divisions
  const int step_item = (int)gridDim.x / cfg.n_rows, step_row = (int)gridDim.x % cfg.n_rows;
  auto advance = [&](int& it, int& tr) {
    tr += step_row;
    it += step_item;
    if (tr >= cfg.n_rows) { tr -= cfg.n_rows; ++it; }
  };

  if (tid == 0) {
    mbar_init(&fs.mbar_tma, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&fs.mbar_ready[b], NW);
      mbar_init(&fs.mbar_done[b], 1);
      fs.zbuf[b] = make_float2(P.z_guess, P.win_sd_first);
      for (int i = 0; i < 64; ++i) fs.hist[b][i] = 0u;
    }
    fs.redo_n = 0;
    if ((int)blockIdx.x < n_rows_total)
      issue_row(base_of((int)blockIdx.x / cfg.n_rows, (int)blockIdx.x % cfg.n_rows));
  }
  __syncthreads();

  const int rows_of_cta = (int)blockIdx.x < n_rows_total ? (n_rows_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid >= NT) {
    // ================================================================================================== select warps
    const int sl = lane;
    const uint32_t my_parity = (uint32_t)((tid - NT) >> 5);   // select warp 0: even image rows, warp 1: odd ones
    float* cand = fs.cand[my_parity];
    const int w_seg = sl / LPS, sub = sl % LPS;
    float z_run = P.z_guess, win_run = P.win_sd_first;
    uint32_t jj = 0;
    for (int row = blockIdx.x; row < n_rows_total; row += gridDim.x) {
      const int kind = P.in.row_kinds ? (int)P.in.row_kinds[row] : LANTERN_ROW_IMAGE;
      if (kind != LANTERN_ROW_IMAGE) continue;     // one-hot rows never reach the select warps
      if ((jj & 1) != my_parity) { ++jj; continue; }
      const int hb = (int)(jj & 1);                       // hand-over buffer of this row
      mbar_wait_backoff(&fs.mbar_ready[hb], (jj >> 1) & 1, 200);   // two row periods of slack: poll rarely
      const float* my_seg = seg + (hb * NW + w_seg) * kSegCap;
      unsigned* hist = fs.hist[hb];
      const RowDesc& dsc = fs.desc[hb];
      const float lo0 = dsc.lo, hi0 = dsc.hi, mean = dsc.mean, inv_sd = dsc.inv_sd, m = dsc.mx;
      const int flags = dsc.flags;
      const unsigned cw = sl < NW ? fs.cnt_part[hb][sl] : 0u;
      const bool overflow = __any_sync(0xffffffffu, cw == kWarpOverflow);
      const unsigned csum = __reduce_add_sync(0xffffffffu, cw);
      const int tot_above = (int)(csum & 0xffffu), tot_in = (int)(csum >> 16);
      const int n_w = (int)(__shfl_sync(0xffffffffu, cw, w_seg) >> 16);
      const int k = cfg.top_k;
      bool ok = !(flags & kRowRedo) && !overflow;
      float thr = -INFINITY;
      bool hit = false;
      float part = sl < NW ? fs.sab_part[hb][sl] : 0.f;   // kept mass above the bracket, summed by the main warps
      const ExpShift ex(m);
      if (ok && P.do_topk) {
        ok = tot_above < k && k <= tot_above + tot_in;
        if (ok) {
          const int krem = k - tot_above;      // rank among the parked elements (1-based from the top)
          const int trips = __reduce_max_sync(0xffffffffu, (n_w + 4 * LPS - 1) / (4 * LPS));
          float lo = lo0, hi = hi0;
          ok = false;
#pragma unroll 1
          for (int it = 0; it < kSelMaxIters; ++it) {
            Classifier64 cls;                  // any monotone classifier keeps the select exact
            cls.scale = __fdividef(61.0f, hi - lo);
            cls.bias23 = fmaf(-lo, cls.scale, 1.0f) + 8388608.0f;
            if (!isfinite(cls.scale) || !isfinite(cls.bias23)) break;
            if (it > 0) {   // the first histogram was filled by the main warps while they compacted their segments
              hist[2 * sl] = 0u;
              hist[2 * sl + 1] = 0u;
              __syncwarp();
              for (int i = sub; i < n_w; i += LPS) atomicAdd(&hist[cls(my_seg[i])], 1u);
              __syncwarp();
            }
            // lane l owns fields 2l, 2l+1; suffix sums from the top field down
            const unsigned c0 = hist[2 * sl], c1 = hist[2 * sl + 1];
            unsigned incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const unsigned n = __shfl_down_sync(0xffffffffu, incl, o);
              if (sl + o < 32) incl += n;
            }
            const unsigned above_pair = incl - (c0 + c1);
            int pick = -1;
            if (above_pair < (unsigned)krem && above_pair + c1 >= (unsigned)krem) pick = 1;
            else if (above_pair + c1 < (unsigned)krem && above_pair + c1 + c0 >= (unsigned)krem) pick = 0;
            const unsigned owner = __ballot_sync(0xffffffffu, pick >= 0);
            const int src = __ffs(owner) - 1;
            const unsigned F = (unsigned)__shfl_sync(0xffffffffu, 2 * sl + (pick > 0 ? 1 : 0), src);
            const int above2 = __shfl_sync(0xffffffffu, (int)(pick == 1 ? above_pair : above_pair + c1), src);
            const int cntF = __shfl_sync(0xffffffffu, (int)(pick == 1 ? c1 : c0), src);
            if (cntF <= kCandCap) {
              // one sweep over the segments: elements in fields above F are kept for sure (their exp is summed in a
              // fixed per-lane order), the members of F are appended to the candidate list in a fixed order too
              // (ballot prefix, not atomics): the sum stays reproducible bit for bit
              int n_cand = 0;
              float acc = 0.f;
#pragma unroll 1
              for (int t = 0; t < trips; ++t) {      // four consecutive elements per lane and trip (one 128-bit read)
                const int i = (t * LPS + sub) * 4;
                const float4 q = *reinterpret_cast<const float4*>(my_seg + i);
                const float v4[4] = {q.x, q.y, q.z, q.w};
                unsigned f4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) f4[j] = cls(v4[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const bool valid = i + j < n_w;
                  if (valid && f4[j] > F) acc += ex(v4[j]);
                  const bool is_f = valid && f4[j] == F;
                  const unsigned mk = __ballot_sync(0xffffffffu, is_f);
                  if (is_f) cand[n_cand + __popc(mk & ((1u << sl) - 1u))] = v4[j];
                  n_cand += __popc(mk);
                }
              }
              __syncwarp();
              const int kr = krem - above2;
              float found = -INFINITY;     // every qualifying candidate carries the same value
              for (int i0 = 0; i0 < cntF; i0 += 32) {
                const int i = i0 + sl;
                const float vi = i < cntF ? cand[i] : INFINITY;
                int gt = 0, ge = 0;
                for (int j = 0; j < cntF; ++j) {
                  const float vj = cand[j];     // broadcast read
                  gt += vj > vi;
                  ge += vj >= vi;
                }
                if (i < cntF && gt < kr && kr <= ge) found = vi;
              }
              thr = warp_reduce(found, OpMaxF());
              for (int i0 = 0; i0 < cntF; i0 += 32) {
                const int i = i0 + sl;
                if (i < cntF) {
                  const float vi = cand[i];
                  if (vi >= thr) acc += ex(vi);
                }
              }
              part += acc;
              ok = true;
              break;
            }
            // too many candidates (ties / dense bracket): shrink to the exact [min, max] of field F and repeat
            float mn = INFINITY, mxv = -INFINITY;
            for (int i = sub; i < n_w; i += LPS) {
              const float v = my_seg[i];
              if (cls(v) == F) { mn = fminf(mn, v); mxv = fmaxf(mxv, v); }
            }
            mn = -warp_reduce(-mn, OpMaxF());
            mxv = warp_reduce(mxv, OpMaxF());
            if (mn == mxv) {     // one value fills the field: it is the threshold; everything >= it is kept
              thr = mn;
              float acc = 0.f;
              for (int i = sub; i < n_w; i += LPS) {
                const float v = my_seg[i];
                if (v >= thr) acc += ex(v);
              }
              part += acc;
              ok = true;
              break;
            }
            lo = mn; hi = mxv;
            __syncwarp();
          }
          hit = ok;
        }
      }
      // the histogram is refilled by the main warps when they come back to this buffer
      hist[2 * sl] = 0u;
      hist[2 * sl + 1] = 0u;
      if (ok) {
        part = warp_reduce(part, OpSum());
        if (sl == 0) {
          RowStats st;
          st.thr = thr; st.mx = m; st.sum = part; st.vcut = -INFINITY; st.icut = -1;
          st.kind = LANTERN_ROW_IMAGE; st.pad0 = st.pad1 = 0;
          P.stats[row] = st;
        }
      } else if (sl == 0) {
        RowStats st;      // finished after the stream by the main threads (pad0 marks the row)
        st.thr = -INFINITY; st.mx = m; st.sum = 1.f; st.vcut = -INFINITY; st.icut = -1;
        st.kind = LANTERN_ROW_IMAGE; st.pad0 = 1; st.pad1 = 0;
        P.stats[row] = st;
        atomicAdd(&fs.redo_n, 1);
      }
      // where the quantile really was (in standard deviations) steers the next row's bracket; the width is the one
      // the sampling noise of the quantile calls for after a hit, doubled after every consecutive miss
      if (P.do_topk) {
        const float z_obs = (thr - mean) * inv_sd;
        if (hit && isfinite(z_obs)) { z_run = z_obs; win_run = P.win_sd; }
        else win_run = fminf(0.25f, 2.0f * fmaxf(win_run, P.win_sd));
      }
      if (sl == 0) fs.zbuf[hb] = make_float2(z_run, win_run);
      __syncwarp();
      if (sl == 0) mbar_arrive(&fs.mbar_done[hb]);
      ++jj;
    }
  } else {
  // ======================================================================================================= main threads
  uint32_t parity = 0, jj = 0;
  int item = (int)blockIdx.x / cfg.n_rows, trow = (int)blockIdx.x % cfg.n_rows;
  int n_item = item, n_trow = trow;
  advance(n_item, n_trow);
  unsigned it_row = 0;
  for (int row = blockIdx.x; row < n_rows_total; row += gridDim.x, ++it_row) {
    const int kind = P.in.row_kinds ? (int)P.in.row_kinds[row] : LANTERN_ROW_IMAGE;
    const int64_t base = base_of(item, trow);
    const int lead_c = (int)((reinterpret_cast<uintptr_t>(P.in.logits_cond) + (uintptr_t)base * EB) & 15);
    const int lead_u = MODE == 1 ? (int)((reinterpret_cast<uintptr_t>(P.in.logits_uncond) + (uintptr_t)base * EB) & 15) : 0;

    // ---- pass 1: lift the staged row into registers: CFG mix (+ temperature) + per-thread statistics ----
    mbar_wait(&fs.mbar_tma, parity);

    parity ^= 1;
    float s[NE];
    float fsum, fsq, fmx = -INFINITY;
    {
      uint64_t sum2 = pack2(0.f, 0.f), sq2 = pack2(0.f, 0.f);   // even / odd elements: the moments only steer the bracket
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e0 = (q * NT + tid) * 4;
        float c4[4], u4[4] = {0.f, 0.f, 0.f, 0.f};
        lds4<DT>(buf_c + lead_c, e0, c4);
        if (MODE == 1) lds4<DT>(buf_u + lead_u, e0, u4);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const uint64_t v2 = mix_temper2(c4[j], c4[j + 1], u4[j], u4[j + 1], mix);
          unpack2(v2, s[q * 4 + j], s[q * 4 + j + 1]);
          sum2 = add2(sum2, v2);
          sq2 = fma2(v2, v2, sq2);
          fmx = fmaxf(fmx, fmaxf(s[q * 4 + j], s[q * 4 + j + 1]));
        }
      }
      float a, b;
      unpack2(sum2, a, b);
      fsum = a + b;
      unpack2(sq2, a, b);
      fsq = a + b;
    }
    fsum = warp_reduce(fsum, OpSum());
    fsq = warp_reduce(fsq, OpSum());
    fmx = warp_reduce(fmx, OpMaxF());
    float4* stp = fs.st_part[it_row & 1];
    if (lane == 0) stp[warp] = make_float4(fsum, fsq, 0.f, fmx);
    MainBar::sync();   // statistics partials visible; every main thread has consumed the staged row
    if (tid == 0 && row + (int)gridDim.x < n_rows_total) issue_row(base_of(n_item, n_trow));
    item = n_item; trow = n_trow;
    advance(n_item, n_trow);
    if (kind != LANTERN_ROW_IMAGE) {   // one-hot rows (Lumina newline / end-of-image): no statistics needed
      if (tid == 0) {
        RowStats st;
        st.thr = -INFINITY; st.mx = 0.f; st.sum = 1.f; st.vcut = -INFINITY; st.icut = -1;
        st.kind = kind; st.pad0 = st.pad1 = 0;
        P.stats[row] = st;
      }
      continue;
    }
    {   // lane l takes warp (l mod NW)'s partial; a butterfly over NW lanes leaves the totals in every lane
      const float4 pw = stp[lane & (NW - 1)];
      fsum = pw.x; fsq = pw.y; fmx = pw.w;
#pragma unroll
      for (int o = 1; o < NW; o <<= 1) {
        fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
        fsq += __shfl_xor_sync(0xffffffffu, fsq, o);
        fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
      }
    }
    const float m = fmx;
    // moments once per row; they only steer the bracket (exactness comes from the counts), so the reciprocal square
    // root may be the approximate one.  A constant row gives 0 * inf = NaN: empty bracket -> redo path.
    const float mean = fsum * P.inv_ncols;
    const float var = fmaxf(fmaf(fsq, P.inv_ncols, -mean * mean), 0.f);
    const float inv_sd = rsqrtf(var);
    const float sd = var * inv_sd;
    // the hand-over buffers of this row were last used two rows ago: the select warp has normally long finished that
    // row; what it observed there steers this row's bracket
    const int hb = (int)(jj & 1);
    mbar_wait_backoff(&fs.mbar_done[hb], ((jj >> 1) & 1) ^ 1, 40);

    ++jj;
    float lo, hi;
    bool good = isfinite(fmx) && isfinite(fsq);     // -inf / +inf / NaN elements: the redo path handles the row
    if (P.do_topk) {
      const float2 zw = fs.zbuf[hb];
      lo = mean + (zw.x - zw.y) * sd;
      hi = mean + (zw.x + zw.y) * sd;
      good = good && lo < hi;
    } else {
      lo = hi = -INFINITY;                          // keep all: every finite element is "above"
    }
    unsigned wcnt = 0u;
    float sab = 0.f;
    if (good) {
      // ---- pass 2: exp of every element, sum / count above the bracket, elements inside it parked ----
      const float nml2 = -m * 1.4426950408889634f;
      int above = 0;
      const uint32_t park0 = smem_u32(park + tid);
      const uint32_t plimit = park0 + (uint32_t)((PS - 4) * NT * 4);   // CLAMP: at most PS slots are ever written
      uint32_t paddr = park0;
      // groups of eight: the exps of a group are issued back to back (MUFU latency overlaps), then the predicated
      // sum / count / park instructions of the group consume them
#pragma unroll
      for (int g = 0; g < NE; g += 8) {
        float ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(s[g + j], 1.4426950408889634f, nml2);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ev[j]) : "f"(t));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (CLAMP && (j & 3) == 0 && (g + j) > 0) paddr = min(paddr, plimit);
          asm volatile(
              "{\n"
              ".reg .pred pa, pin;\n"
              "setp.gt.f32 pa, %3, %5;\n"
              "setp.ge.and.f32 pin, %3, %4, !pa;\n"
              "@pa add.f32 %2, %2, %6;\n"
              "@pa add.s32 %1, %1, 1;\n"
              "@pin st.shared.f32 [%0], %3;\n"
              "@pin add.u32 %0, %0, %7;\n"
              "}\n"
              : "+r"(paddr), "+r"(above), "+f"(sab)
              : "f"(s[g + j]), "f"(lo), "f"(hi), "f"(ev[j]), "n"(NT * 4)
              : "memory");
        }
      }
      const int slot = (int)((paddr - park0) / (NT * 4));
      // the warp's parked elements -> one contiguous segment (exclusive scan of the per-thread counts)
      int incl = slot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      const int off = incl - slot;
      const int wtot = __shfl_sync(0xffffffffu, incl, 31);
      const int wmax = __reduce_max_sync(0xffffffffu, slot);
      const int wa = __reduce_add_sync(0xffffffffu, above);
      const bool clamped = CLAMP && __any_sync(0xffffffffu, paddr >= plimit);
      if (clamped || wmax > PS || wtot > kSegCap) {
        wcnt = kWarpOverflow;     // more parked elements than the hand-over area holds: the row goes to the redo path
      } else {
        float* dst = seg + (hb * NW + warp) * kSegCap + off;
        Classifier64 cls;     // same classifier as the select warp's first iteration
        cls.scale = __fdividef(61.0f, hi - lo);
        cls.bias23 = fmaf(-lo, cls.scale, 1.0f) + 8388608.0f;
#pragma unroll 4
        for (int i = 0; i < wmax; ++i) {
          if (i < slot) {
            const float v = park[i * NT + tid];
            dst[i] = v;
            atomicAdd(&fs.hist[hb][cls(v)], 1u);
          }
        }
        wcnt = (unsigned)wa | ((unsigned)wtot << 16);
      }
      sab = warp_reduce(sab, OpSum());
    }
    if (lane == 0) {
      fs.cnt_part[hb][warp] = wcnt;
      fs.sab_part[hb][warp] = sab;
    }
    if (tid == 0) {
      RowDesc d;
      d.lo = lo; d.hi = hi; d.mean = mean; d.inv_sd = inv_sd; d.mx = m;
      d.flags = good ? 0 : kRowRedo;
      fs.desc[hb] = d;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&fs.mbar_ready[hb]);
  }
  }   // roles
  __syncthreads();   // one barrier for both roles: publishes redo_n and the flagged records to the main threads
  if (tid >= NT || fs.redo_n == 0) return;
  uint32_t parity = (uint32_t)(rows_of_cta & 1);   // the TMA barrier has completed one phase per row of this CTA
  int item, trow;

  // ---- redo: rows the streaming select could not finish; exact tiers 2 / 3 on the re-staged row ----
  item = (int)blockIdx.x / cfg.n_rows; trow = (int)blockIdx.x % cfg.n_rows;
  for (int row = blockIdx.x; row < n_rows_total; row += gridDim.x) {
    const int64_t base = base_of(item, trow);
    advance(item, trow);
    if (P.stats[row].pad0 != 1) continue;      // block-uniform: written before the barrier above
    const int lead_c = (int)((reinterpret_cast<uintptr_t>(P.in.logits_cond) + (uintptr_t)base * EB) & 15);
    const int lead_u = MODE == 1 ? (int)((reinterpret_cast<uintptr_t>(P.in.logits_uncond) + (uintptr_t)base * EB) & 15) : 0;
    if (tid == 0) issue_row(base);
    mbar_wait(&fs.mbar_tma, parity);
    parity ^= 1;
    float s[NE];
    float fmn = INFINITY, fmx = -INFINITY;
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
        fmn = fminf(fmn, v);
        fmx = fmaxf(fmx, v);
      }
    }
    fmn = -group_reduce<MainBar>(-fmn, OpMaxF(), -INFINITY, slow_scr);
    fmx = group_reduce<MainBar>(fmx, OpMaxF(), -INFINITY, slow_scr);
    float thr = -INFINITY;
    if (P.do_topk) {
      float tmp[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) tmp[e] = s[e];
      thr = (fmn == fmx) ? fmx : select_slow<NE, MainBar>(tmp, cfg.top_k, fmn, fmx, sm);
      MainBar::sync();
    }
    const ExpShift ex(fmx);
    float part = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) part += (s[e] >= thr) ? ex(s[e]) : 0.f;
    const float tot = group_reduce<MainBar>(part, OpSum(), 0.f, slow_scr);
    if (tid == 0) {
      RowStats st;
      st.thr = thr; st.mx = fmx; st.sum = tot; st.vcut = -INFINITY; st.icut = -1;
      st.kind = LANTERN_ROW_IMAGE; st.pad0 = st.pad1 = 0;
      P.stats[row] = st;
    }
    MainBar::sync();   // the stage buffers are re-used by the next redo row
  }
}

}  // namespace lantern
