// Fused relaxed-acceptance verify step (sm_100a).
//
// Two phases behind one entry point (lantern_accept_fused):
//
//   row_stats_kernel  — HBM-bound.  One CTA per logits row (persistent grid over B*T rows): streams the
//                       cond/uncond rows once with 128-bit loads, applies the CFG mix and temperature,
//                       finds the top-k threshold (exact k-th largest, ties kept) and the softmax max / sum
//                       over the kept columns.  Only 32 bytes per row go back to HBM.
//   walk_kernel       — one CTA per prompt.  Literal restatement of evaluate_posterior[_v1]: the level
//                       loop, candidate dedup, latent-proximity relaxation (neighbour-table gather from the
//                       shared-memory probability vector, fp64 warp-shuffle scan, threshold search),
//                       accept test against the supplied / Philox uniform, residual bookkeeping, tail
//                       distribution and the bonus-token inverse-CDF draw.  The [L,D,V] gather of
//                       tree_decoding is never materialised: rows are addressed through retrieve_indices.
//
// Reference semantics: SURVEY.md Appendix A; file:line citations are in include/lantern_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "accept_types.cuh"
#include "select.cuh"
#include "stats_fast.cuh"
#include "stats_stream.cuh"
#include "stats_wide.cuh"
#include "topp.cuh"

namespace lantern {

// Phase trace of the walk (profiling builds only: -DLANTERN_WALK_TRACE, profiles/walk_trace.py).  Thread 0 of every CTA
// keeps (tag, clock) pairs in shared memory and copies them out once at the end, so a mark costs a few cycles.
#ifdef LANTERN_WALK_TRACE
constexpr int kTraceMax = 120, kTraceBlocks = 256;
__device__ long long g_trace[kTraceBlocks][2 * kTraceMax + 2];
__shared__ long long tr_buf[2 * kTraceMax];
__shared__ int tr_n;
#define TR(tag) do { if (threadIdx.x == 0 && tr_n < kTraceMax) { tr_buf[2 * tr_n] = (tag); tr_buf[2 * tr_n + 1] = clock64(); ++tr_n; } } while (0)
#else
#define TR(tag) do { } while (0)
#endif

constexpr int kWalkThreads = 1024;
// Lazy schedule: few, fat threads (32 row values each at 8192 columns).  The walk CTA is alone on its SM and most of its
// instructions are per-thread bookkeeping every warp repeats (reduction tails, scans, address arithmetic): with W warps
// on four schedulers each such instruction costs W/4 issue cycles, so 8 warps beat 16 (and 16 beat 32, round 1).
#ifndef LANTERN_LAZY_THREADS
#define LANTERN_LAZY_THREADS 256
#endif
constexpr int kLazyThreads = LANTERN_LAZY_THREADS;
constexpr int kMaxSib = 64;


// ----------------------------------------------------------------------------------------------
// Phase 1: per-row statistics
// ----------------------------------------------------------------------------------------------
template <int DT>
__host__ __device__ constexpr int cfg_ncols_bytes(int ncols) { return ncols * Elem<DT>::kBytes; }

// Generic row-statistics kernel: any window width up to NT * NQ * 4 columns, bounds-checked loads, runtime CFG
// switch, -inf aware statistics.  The TMA-staged kernel for the power-of-two windows lives in stats_fast.cuh.
template <int DT, int NT, int NQ, bool VEC>
__global__ void __launch_bounds__(NT, (NQ <= 8 ? 1024 / NT : 512 / NT)) row_stats_kernel(const AcceptParams P) {
  constexpr int NE = NQ * 4;
  constexpr int NW = NT / 32;
  __shared__ SelectSmem sm;
  __shared__ float fscratch[33];
  extern __shared__ __align__(128) unsigned char dyn_smem[];   // the [NE][NT] parking columns of the bracket select
  float* park = reinterpret_cast<float*>(dyn_smem);

  const lantern_accept_cfg& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n_rows_total = (long long)cfg.n_items * cfg.n_rows;
  const MixParams mix = P.mix;
  const bool has_uncond = mix.has_uncond;

  for (long long row = blockIdx.x; row < n_rows_total; row += gridDim.x) {
    const int b = (int)(row / cfg.n_rows), t = (int)(row % cfg.n_rows);
    RowStats st;
    st.thr = -INFINITY; st.mx = 0.f; st.sum = 1.f; st.vcut = -INFINITY; st.icut = -1;
    st.kind = P.in.row_kinds ? (int)P.in.row_kinds[row] : LANTERN_ROW_IMAGE;
    st.pad0 = st.pad1 = 0;
    if (st.kind != LANTERN_ROW_IMAGE) {   // one-hot rows: nothing to read
      if (tid == 0) P.stats[row] = st;
      continue;
    }
    const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)t * cfg.row_stride + cfg.col0;
    // pull the next row of this CTA towards L2 while this one is being processed
    if (tid == 0 && row + gridDim.x < n_rows_total) {
      const long long nr = row + gridDim.x;
      const int64_t nbase = (nr / cfg.n_rows) * cfg.item_stride + (nr % cfg.n_rows) * cfg.row_stride + cfg.col0;
      const size_t eb = Elem<DT>::kBytes;
      const uintptr_t a0 = (reinterpret_cast<uintptr_t>(P.in.logits_cond) + nbase * eb + 15) & ~uintptr_t(15);
      const uintptr_t a1 = (reinterpret_cast<uintptr_t>(P.in.logits_cond) + (nbase + cfg.ncols) * eb) & ~uintptr_t(15);
      if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)));
      if (has_uncond) {
        const uintptr_t b0 = (reinterpret_cast<uintptr_t>(P.in.logits_uncond) + nbase * eb + 15) & ~uintptr_t(15);
        const uintptr_t b1 = (reinterpret_cast<uintptr_t>(P.in.logits_uncond) + (nbase + cfg.ncols) * eb) & ~uintptr_t(15);
        if (b1 > b0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b0), "r"((unsigned)(b1 - b0)));
      }
    }
    // ---- lift the row into registers (fully unrolled) ----
    float s[NE];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e0 = (q * NT + tid) * 4;
      float c4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, u4[4] = {0.f, 0.f, 0.f, 0.f};
      if (VEC) {
        if (e0 < cfg.ncols) {
          Elem<DT>::load4(P.in.logits_cond, base + e0, c4);
          if (has_uncond) Elem<DT>::load4(P.in.logits_uncond, base + e0, u4);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (e0 + j < cfg.ncols) {
            c4[j] = Elem<DT>::load1(P.in.logits_cond, base + e0 + j);
            if (has_uncond) u4[j] = Elem<DT>::load1(P.in.logits_uncond, base + e0 + j);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) s[q * 4 + j] = mix_temper(c4[j], u4[j], mix);   // padding: -inf stays -inf
    }
    // ---- row statistics: sum, sum of squares, min, max, finite count; one fused two-level reduction ----
    float fsum = 0.f, fsq = 0.f, fmin_ = INFINITY, fmax_ = -INFINITY;
    int nfin = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const float v = s[e];
      fmax_ = fmaxf(fmax_, v);
      if (v > -INFINITY) {
        fsum += v;
        fsq = fmaf(v, v, fsq);
        fmin_ = fminf(fmin_, v);
        ++nfin;
      }
    }
    fsum = warp_reduce(fsum, OpSum());
    fsq = warp_reduce(fsq, OpSum());
    fmin_ = -warp_reduce(-fmin_, OpMaxF());
    fmax_ = warp_reduce(fmax_, OpMaxF());
    nfin = __reduce_add_sync(0xffffffffu, nfin);
    __syncthreads();
    if (lane == 0) {
      sm.f4[0][warp] = fsum; sm.f4[1][warp] = fsq; sm.f4[2][warp] = fmin_; sm.f4[3][warp] = fmax_;
      sm.warp_cnt[warp][0] = (unsigned)nfin;
    }
    __syncthreads();
    if (warp == 0) {
      float a = lane < NW ? sm.f4[0][lane] : 0.f, q2 = lane < NW ? sm.f4[1][lane] : 0.f;
      float mn = lane < NW ? sm.f4[2][lane] : INFINITY, mx = lane < NW ? sm.f4[3][lane] : -INFINITY;
      int n = lane < NW ? (int)sm.warp_cnt[lane][0] : 0;
      a = warp_reduce(a, OpSum()); q2 = warp_reduce(q2, OpSum());
      mn = -warp_reduce(-mn, OpMaxF()); mx = warp_reduce(mx, OpMaxF());
      n = __reduce_add_sync(0xffffffffu, n);
      if (lane == 0) { sm.f_scr[2] = a; sm.f_scr[3] = q2; sm.f_scr[4] = mn; sm.f_scr[5] = mx; sm.i_scr[4] = n; }
    }
    __syncthreads();
    fsum = sm.f_scr[2]; fsq = sm.f_scr[3]; fmin_ = sm.f_scr[4]; fmax_ = sm.f_scr[5];
    nfin = sm.i_scr[4];
    const float m = fmax_;
    // ---- top-k threshold (exact k-th largest; ties are kept by the >= test below) ----
    if (P.do_topk) {
      bool found = false;
      float thr = -INFINITY;
      const bool finite = isfinite(fmin_) && isfinite(fmax_) && isfinite(fsq);
      if (finite && fmin_ == fmax_) { thr = fmax_; found = true; }
      if (!found && finite) {                                   // tier 1: moment bracket
        const float inv_n = 1.0f / (float)nfin;
        const float mean = fsum * inv_n;
        const float sd = sqrtf(fmaxf(fsq * inv_n - mean * mean, 0.f));
        const float lo = mean + (P.z_guess - P.win_sd_first) * sd, hi = mean + (P.z_guess + P.win_sd_first) * sd;
        if (lo < hi && cfg.top_k <= nfin) found = bracket_select<NE, NT>(s, cfg.top_k, lo, hi, park, sm, &thr);
      }
      if (!found) {                                             // tiers 2 and 3 work on a copy: s[] stays in registers
        float tmp[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) tmp[e] = s[e];
        thr = select_slow<NE>(tmp, cfg.top_k, fmin_, fmax_, sm);
      }
      st.thr = thr;
    }
    // ---- softmax sum over the kept columns ----
    const ExpShift ex(m);
    float part = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const float ev = ex(s[e]);           // padded / masked slots: ex2(-inf) = 0
      part += (s[e] >= st.thr) ? ev : 0.f;
    }
    st.mx = m;
    st.sum = block_reduce(part, OpSum(), 0.f, fscratch);
    if (tid == 0) P.stats[row] = st;
  }
}

// ----------------------------------------------------------------------------------------------
// Phase 2: the acceptance walk, one CTA per prompt
// ----------------------------------------------------------------------------------------------
struct WalkSmem {
  float* p;          // [ncols] probability window (gtp)
  unsigned* nbmask;  // [ceil(ncols/32)] neighbour bitmap (static LlamaGen/Anole rejection)
  int* ri;           // [L*D]
  int* tok;          // [T]
  int* tried;        // [L] distinct child tokens of the current node, first-occurrence order (kid_x)
  int* kid_j;        // [L] candidates row that first reaches each child
  int* kid_node;     // [L] tree node of each child
  float* csv;        // [kWalkThreads] per-thread cumulative sums of the current neighbour chunk
  float* uni;        // [T + 1] this item's uniforms (supplied or Philox)
  int* sib;          // [kMaxSib] tokens of the rejected node's earlier siblings
  int* pid;          // [L] rows still matching the accepted prefix
  int* ctok;         // [L*D] token at every (row, level) of the path table, -1 = padding
  int* csyn;         // [L*D] Lumina: the token at (row, level) is a syntax token (accepted with p = 1)
  int* maxcp;        // [L] longest token prefix a row shares with an earlier row (static: dedup of the children)
  int* kid_syn;      // [L] Lumina: the child token is a syntax token (accepted with p = 1)
  double* dscr;      // [34]
  float* fscr;       // [34]
  int* iscr;         // [40]
};

template <int DT, bool VEC>
__device__ __forceinline__ void load_probs(const AcceptParams& P, int b, int node, const RowStats& st, float* p,
                                           bool raw) {
  const lantern_accept_cfg& cfg = P.cfg;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)node * cfg.row_stride + cfg.col0;
  MixParams mix = P.mix;
  if (raw) mix.do_temp = 0;
  const float inv = __fdiv_rn(1.0f, st.sum);
  const ExpShift ex(st.mx);
  const int nquads = (cfg.ncols + 3) >> 2;
  for (int g = threadIdx.x; g < nquads; g += blockDim.x) {
    const int e0 = g * 4;
    float c4[4], u4[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {
      Elem<DT>::load4(P.in.logits_cond, base + e0, c4);
      if (mix.has_uncond) Elem<DT>::load4(P.in.logits_uncond, base + e0, u4);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (e0 + j < cfg.ncols) {
          c4[j] = Elem<DT>::load1(P.in.logits_cond, base + e0 + j);
          if (mix.has_uncond) u4[j] = Elem<DT>::load1(P.in.logits_uncond, base + e0 + j);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (e0 + j < cfg.ncols) {
        const float s = mix_temper(c4[j], u4[j], mix);
        float v = 0.f;
        if (raw || kept_col(s, e0 + j, st)) v = __fmul_rn(ex(s), inv);
        p[e0 + j] = v;
      }
    }
  }
}

// raw softmax statistics of one row (vanilla tail row, drafters/utils.py:408-409)
template <int DT, bool VEC>
__device__ __forceinline__ RowStats raw_row_stats(const AcceptParams& P, int b, int node, float* fscr, double* dscr) {
  const lantern_accept_cfg& cfg = P.cfg;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)node * cfg.row_stride + cfg.col0;
  MixParams mix = P.mix;
  mix.do_temp = 0;
  float m = -INFINITY;
  for (int e = threadIdx.x; e < cfg.ncols; e += blockDim.x) {
    const float c = Elem<DT>::load1(P.in.logits_cond, base + e);
    const float u = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + e) : 0.f;
    m = fmaxf(m, mix_temper(c, u, mix));
  }
  m = block_reduce(m, OpMaxF(), -INFINITY, fscr);
  float part = 0.f;
  const ExpShift ex(m);
  for (int e = threadIdx.x; e < cfg.ncols; e += blockDim.x) {
    const float c = Elem<DT>::load1(P.in.logits_cond, base + e);
    const float u = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + e) : 0.f;
    part += ex(mix_temper(c, u, mix));
  }
  const double tot = block_reduce((double)part, OpSum(), 0.0, dscr);
  RowStats st;
  st.thr = -INFINITY; st.mx = m; st.sum = (float)tot; st.vcut = -INFINITY; st.icut = -1;
  st.kind = LANTERN_ROW_IMAGE; st.pad0 = st.pad1 = 0;
  return st;
}

// Lazy mode (phases bit 2): statistics of a visited row computed inside the walk CTA, straight from the logits,
// and the probability vector written in place (one global read of the row instead of two, and no statistics for the
// ~85 % of tree rows the walk never visits).  Same arithmetic as the streamed kernels.
//
// The walk waits for every row it visits, so this routine is built for latency, not throughput (round 2: the first
// version spent 13 block barriers per row, ~7800 cycles after the loads had landed):
//   * four barriers on the common path: moments; bracket counts; candidate list + kept mass of the fields above the
//     target field; threshold;
//   * no shared-memory atomics; the scan of the 16 field counts is done by every warp itself (no broadcast barrier),
//     the ranking of the candidates one candidate per thread;
//   * `hook` runs between issuing the row's loads and their first use: the walk lists the children of the current node
//     (one warp) inside the load shadow; `post` runs behind the first barrier, where the hook's results are visible;
//   * the row lives in shared memory and the per-element passes are rolled loops (see lazy_probs below).
// Rows the moment bracket misses (non-Gaussian rows, heavy ties, non-finite values) take the exact tier-2/3 selectors
// of select.cuh as before.  Returns true (and writes nothing) when every live column is -inf: a pre-masked one-hot
// row (see set_distribution).
constexpr int kLazyListMax = 128;
struct __align__(16) LazySmem {
  float4 mom[kLazyThreads / 32];        // per-warp (sum, sum of squares, min, max)
  int4 cnt[kLazyThreads / 32];          // per-warp (elements above the bracket, parked elements, bits of their exp-sum, -)
  unsigned wh[kLazyThreads / 32][8];    // per-warp counts of the 16 bracket fields, two 16-bit counts per word
  float list[kLazyListMax];             // the members of the field that holds the k-th largest value (16-byte aligned: the
                                        // compiler reads it with vector loads)
  float part[kLazyThreads / 32];        // per-warp kept mass
  float thr;
  float pad[3];
};

// one pass of the CFG mix + moments over the thread's raw row values (unrolled: the values sit in registers), the mixed
// logits stored to the thread's own slots of `p`.  HAS_UNCOND / DO_TEMP are compile-time here: with run-time flags
// the compiler emits both arms for every pair, which tripled the code of this pass.
template <int NQ, int NT, bool HAS_UNCOND, bool DO_TEMP>
__device__ __forceinline__ void lazy_lift(const float (&cc)[NQ][4], const float (&uu)[NQ][4], const MixParams& mix_in,
                                          float* p, int tid, float& fsum, float& fsq, float& fmx) {
  MixParams mix = mix_in;
  mix.has_uncond = HAS_UNCOND ? 1 : 0;
  mix.do_temp = DO_TEMP ? 1 : 0;
  uint64_t sum2 = pack2(0.f, 0.f), sq2 = pack2(0.f, 0.f);
  float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float4 o;
    const uint64_t a2 = mix_temper2(cc[q][0], cc[q][1], uu[q][0], uu[q][1], mix);
    const uint64_t b2 = mix_temper2(cc[q][2], cc[q][3], uu[q][2], uu[q][3], mix);
    unpack2(a2, o.x, o.y);
    unpack2(b2, o.z, o.w);
    sum2 = add2(sum2, a2); sq2 = fma2(a2, a2, sq2);
    sum2 = add2(sum2, b2); sq2 = fma2(b2, b2, sq2);
    mx = fmaxf(mx, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
    *reinterpret_cast<float4*>(p + (q * NT + tid) * 4) = o;
  }
  float a0, a1;
  unpack2(sum2, a0, a1); fsum = a0 + a1;
  unpack2(sq2, a0, a1); fsq = a0 + a1;
  fmx = mx;
}

// Round 2, second form.  ncu on the first latency-oriented version (registers-resident row, every pass fully unrolled)
// showed the walk stalled on instruction fetch above everything else (no_instruction 4.2 warps per issue, issue slots
// 18 % busy): the kernel executes ~8000 instructions per warp exactly once, so it runs at the rate the instruction
// cache can be refilled from L2.  The row therefore lives in shared memory - in `p`, every thread touching only its
// own slots, so no barrier is involved - and the per-element passes are rolled loops of a few dozen instructions.
template <int DT, int NE, class Hook, class Post>
__device__ __forceinline__ bool lazy_probs(const AcceptParams& P, int b, int node, bool raw, float* p, float* park,
                                           LazySmem& lz, SelectSmem& sm, float& z_run, float& win_run, Hook&& hook,
                                           Post&& post) {
  constexpr int NT = kLazyThreads, NW = NT / 32, NQ = NE / 4;
  constexpr unsigned kFull = 0xffffffffu;
  static_assert(NW % 4 == 0 && NW <= 32, "the field totals below are gathered by 4 lane groups x NW/4 warps");
  const lantern_accept_cfg& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)node * cfg.row_stride + cfg.col0;
  MixParams mix = P.mix;
  if (raw) mix.do_temp = 0;
  float fsum, fsq, fmx;
  {
    // ---- the row's loads go out first; nothing touches them until the hook has run ----
    float cc[NQ][4], uu[NQ][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e0 = (q * NT + tid) * 4;
      Elem<DT>::load4(P.in.logits_cond, base + e0, cc[q]);
      if (mix.has_uncond) Elem<DT>::load4(P.in.logits_uncond, base + e0, uu[q]);
      else { uu[q][0] = uu[q][1] = uu[q][2] = uu[q][3] = 0.f; }
    }
    hook();
    TR(39);
    if (mix.has_uncond) {
      if (mix.do_temp) lazy_lift<NQ, NT, true, true>(cc, uu, mix, p, tid, fsum, fsq, fmx);
      else lazy_lift<NQ, NT, true, false>(cc, uu, mix, p, tid, fsum, fsq, fmx);
    } else {
      if (mix.do_temp) lazy_lift<NQ, NT, false, true>(cc, uu, mix, p, tid, fsum, fsq, fmx);
      else lazy_lift<NQ, NT, false, false>(cc, uu, mix, p, tid, fsum, fsq, fmx);
    }
  }
  TR(40);
  {
    // the maximum through integer keys (one REDUX); the two sums share one butterfly: after the first exchange the lower
    // half-warp carries the sum, the upper one the sum of squares
    const unsigned kmx = __reduce_max_sync(kFull, float_key(fmx));
    const bool up = (lane & 16) != 0;
    float keep = up ? fsq : fsum;
    keep += __shfl_xor_sync(kFull, up ? fsum : fsq, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(kFull, keep, o);
    if (lane == 0) { lz.mom[warp].x = keep; lz.mom[warp].w = key_float(kmx); }
    if (lane == 16) lz.mom[warp].y = keep;
  }
  __syncthreads();   // #1 (also publishes what the hook wrote)
  post();
  fsum = 0.f; fsq = 0.f; fmx = -INFINITY;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const float4 m = lz.mom[w];
    fsum += m.x; fsq += m.y; fmx = fmaxf(fmx, m.w);
  }
  if (fmx == -INFINITY) return true;   // block-uniform
  TR(41);
  const ExpShift ex(fmx);
  float4* const mine = reinterpret_cast<float4*>(p) + tid;   // the thread's quads: mine[q * NT], q < NQ
  float thr = -INFINITY, tot = 0.f;
  bool have_tot = false;
  if (P.do_topk && !raw) {
    // a finite sum of squares means every element is finite (a constant row ends on the slow path: its bracket is
    // empty or holds the whole row)
    const bool finite = isfinite(fmx) && isfinite(fsq);
    // mean, sd, the bracket and the field scale only steer the search (any values keep it exact, and they are the same
    // in every thread): approximate reciprocal / square root instead of the IEEE sequences - on this one-shot
    // instruction stream ~80 instructions per row
    const float inv_n = P.inv_ncols;
    const float mean = fsum * inv_n;
    const float var = fmaxf(fsq * inv_n - mean * mean, 0.f);
    const float sd = var > 0.f ? var * rsqrtf(var) : 0.f;
    const int k = cfg.top_k;
    bool slow = false;
    {
      const float lo = mean + (z_run - win_run) * sd, hi = mean + (z_run + win_run) * sd;
      Classifier cls;   // 16 fields; the bracket itself is tiled by fields 1..14 (make_classifier with a fast division)
      cls.scale = __fdividef(13.0f, hi - lo);
      cls.bias23 = fmaf(-lo, cls.scale, 1.0f) + 8388608.0f;
      const bool fast = finite && lo < hi && isfinite(cls.scale) && isfinite(cls.bias23);
      slow = !fast;
      if (fast) {
        // One sweep: count and exp-sum of the elements above the bracket, the elements inside it parked in the thread's
        // private shared-memory column (branch-free: every element is stored at the running slot, only an element
        // inside the bracket keeps it).  No shared-memory atomics anywhere in this routine: they cost two cycles per
        // lane on this machine, which made an atomic histogram the most expensive step of the row.
        int above = 0;
        float sum_ab = 0.f;
        float* pp = park + tid;
#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
          const float4 v4 = mine[q * NT];
          const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v = vv[j];
            const bool ab = v > hi;
            if (ab) { sum_ab += ex(v); ++above; }
            *pp = v;
            if (!ab && v >= lo) pp += NT;
          }
        }
        const int slot = (int)(pp - (park + tid)) / NT;
        // the thread's parked elements by field: sixteen 8-bit counters in two registers (a thread parks <= NE <= 64)
        unsigned long long ca = 0ull, cb = 0ull;
        for (int i = 0; i < slot; ++i) {
          const unsigned f = cls(park[i * NT + tid]);
          const unsigned long long inc = 1ull << ((f & 7u) * 8u);
          if (f < 8u) ca += inc; else cb += inc;
        }
        // widen to 16-bit pairs and add across the warp: word 2q + r holds fields 4q + r and 4q + r + 2 (r = 0, 1)
        unsigned wsum[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned w32 = q < 2 ? (unsigned)(ca >> (32 * q)) : (unsigned)(cb >> (32 * (q - 2)));
          wsum[2 * q] = __reduce_add_sync(kFull, w32 & 0x00ff00ffu);
          wsum[2 * q + 1] = __reduce_add_sync(kFull, (w32 >> 8) & 0x00ff00ffu);
        }
        const int w_above = __reduce_add_sync(kFull, above), w_in = __reduce_add_sync(kFull, slot);
        const int max_slot = __reduce_max_sync(kFull, slot);
        sum_ab = warp_reduce(sum_ab, OpSum());
        if (lane == 0) {
          lz.cnt[warp] = make_int4(w_above, w_in, __float_as_int(sum_ab), 0);
          *reinterpret_cast<uint4*>(&lz.wh[warp][0]) = make_uint4(wsum[0], wsum[1], wsum[2], wsum[3]);
          *reinterpret_cast<uint4*>(&lz.wh[warp][4]) = make_uint4(wsum[4], wsum[5], wsum[6], wsum[7]);
        }
        TR(45);
        __syncthreads();   // #2
        int tot_above = 0, tot_in = 0;
        float sum_above = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const int4 c = lz.cnt[w];
          tot_above += c.x; tot_in += c.y; sum_above += __int_as_float(c.z);
        }
        TR(46);
        if (tot_above < k && k <= tot_above + tot_in) {
          const int krem = k - tot_above;   // rank among the parked elements
          // Every warp finds the field F that holds the target itself.  Block totals of the packed words (they still
          // fit 16 bits: at most 16384 elements): lane l adds word l & 7 of NW/4 warps, two exchanges finish the sum.
          unsigned wt = 0u;
          {
            const int t = lane & 7, g = lane >> 3;
#pragma unroll
            for (int i = 0; i < NW / 4; ++i) wt += lz.wh[g * (NW / 4) + i][t];
            wt += __shfl_xor_sync(kFull, wt, 8);
            wt += __shfl_xor_sync(kFull, wt, 16);
          }
          // lane f (and its mirror f + 16) takes field f: word 2 (f >> 2) + (f & 1), upper half when f & 2
          const int f = lane & 15;
          const unsigned wf = __shfl_sync(kFull, wt, 2 * (f >> 2) + (f & 1));
          const unsigned cf = (f & 2) ? (wf >> 16) : (wf & 0xffffu);
          unsigned incl = cf;   // suffix sums from the top field down
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) {
            const unsigned n = __shfl_down_sync(kFull, incl, o, 16);
            if (f + o < 16) incl += n;
          }
          const unsigned above_f = incl - cf;
          const bool own = above_f < (unsigned)krem && above_f + cf >= (unsigned)krem;
          const unsigned owner_mask = __ballot_sync(kFull, own) & 0xffffu;
          const int F = owner_mask ? __ffs(owner_mask) - 1 : -1;
          const int above2 = (int)__shfl_sync(kFull, above_f, F >= 0 ? F : 0);
          const int cntF = (int)__shfl_sync(kFull, cf, F >= 0 ? F : 0);
          TR(47);
          if (F >= 0 && cntF <= kLazyListMax) {
            // the members of F go into one list, in (warp, slot, lane) order - a fixed order, so every sum over the
            // list is reproducible; the warp's offset comes from the per-warp field counts, no atomics
            int pos;
            {
              const unsigned ww = lane < NW ? lz.wh[lane][2 * (F >> 2) + (F & 1)] : 0u;
              const unsigned cw = (F & 2) ? (ww >> 16) : (ww & 0xffffu);
              pos = (int)__reduce_add_sync(kFull, lane < warp ? cw : 0u);
            }
            float part = 0.f;   // kept mass of the thread's parked elements in fields above F (they are above thr)
            for (int i = 0; i < max_slot; ++i) {
              const bool have = i < slot;
              const float v = have ? park[i * NT + tid] : 0.f;
              const unsigned fv = have ? cls(v) : 0xffffu;
              const unsigned m = __ballot_sync(kFull, fv == (unsigned)F);
              if (fv == (unsigned)F) lz.list[pos + __popc(m & ((1u << lane) - 1u))] = v;
              pos += __popc(m);
              if (have && fv > (unsigned)F) part += ex(v);
            }
            part = warp_reduce(part, OpSum());
            if (lane == 0) lz.part[warp] = part;
            __syncthreads();   // #3
            TR(48);
            // exact rank inside field F: one candidate per thread against the whole list
            const int kr = krem - above2;
            for (int i = tid; i < cntF; i += NT) {   // one candidate per lane: only the first cntF / 32 warps work
              const float vi = lz.list[i];
              int gt = 0, ge = 0;
#pragma unroll 4
              for (int j = 0; j < cntF; ++j) {
                const float vj = lz.list[j];   // broadcast read
                gt += vj > vi;
                ge += vj >= vi;
              }
              if (gt < kr && kr <= ge) lz.thr = vi;   // every qualifying candidate has the same value
            }
            __syncthreads();   // #4
            thr = lz.thr;
            TR(49);
            float fm = 0.f;   // kept mass inside F, the same fixed order in every warp
            for (int i = lane; i < cntF; i += 32) {
              const float vi = lz.list[i];
              fm += vi >= thr ? ex(vi) : 0.f;
            }
            fm = warp_reduce(fm, OpSum());
            tot = sum_above;
#pragma unroll
            for (int w = 0; w < NW; ++w) tot += lz.part[w];
            tot += fm;
            have_tot = true;
          } else {
            slow = true;
          }
        } else {
          slow = true;
        }
      }
    }
    TR(42);
    if (slow) {   // block-uniform; rare: the exact selectors of select.cuh work on a register copy of the row
      __syncthreads();
      float tmp[NE];
      float fmn = INFINITY;   // the fast path has no use for the row minimum
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 v4 = mine[q * NT];
        tmp[q * 4 + 0] = v4.x; tmp[q * 4 + 1] = v4.y; tmp[q * 4 + 2] = v4.z; tmp[q * 4 + 3] = v4.w;
        fmn = fminf(fmn, fminf(fminf(v4.x, v4.y), fminf(v4.z, v4.w)));
      }
      fmn = -block_reduce(-fmn, OpMaxF(), -INFINITY, sm.f4[0]);
      thr = select_slow<NE>(tmp, k, fmn, fmx, sm);
      __syncthreads();
    }
    const float z_obs = __fdividef(thr - mean, sd);
    if (isfinite(z_obs)) { z_run = z_obs; win_run = P.win_sd; }   // a walk visits too few rows to adapt the width
  }
  TR(43);
  if (!have_tot) {
    float part = 0.f;
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
      const float4 v4 = mine[q * NT];
      part += v4.x >= thr ? ex(v4.x) : 0.f;
      part += v4.y >= thr ? ex(v4.y) : 0.f;
      part += v4.z >= thr ? ex(v4.z) : 0.f;
      part += v4.w >= thr ? ex(v4.w) : 0.f;
    }
    part = warp_reduce(part, OpSum());
    __syncthreads();
    if (lane == 0) lz.part[warp] = part;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) tot += lz.part[w];
  }
  const float inv = __fdiv_rn(1.0f, tot);
  TR(44);
  // logits -> probabilities in place (the exp is taken a second time here: cheaper than carrying it)
#pragma unroll 2
  for (int q = 0; q < NQ; ++q) {
    float4 v4 = mine[q * NT];
    v4.x = v4.x >= thr ? __fmul_rn(ex(v4.x), inv) : 0.f;
    v4.y = v4.y >= thr ? __fmul_rn(ex(v4.y), inv) : 0.f;
    v4.z = v4.z >= thr ? __fmul_rn(ex(v4.z), inv) : 0.f;
    v4.w = v4.w >= thr ? __fmul_rn(ex(v4.w), inv) : 0.f;
    mine[q * NT] = v4;
  }
  return false;
}

// LNE > 0: lazy form (kLazyThreads threads, the walk computes the statistics of the rows it visits); LNE == 0: the walk behind
// the streamed row-statistics kernel (1024 threads).
template <int DT, bool VEC, int LNE>
__global__ void __launch_bounds__(LNE > 0 ? kLazyThreads : kWalkThreads) walk_kernel(const AcceptParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using WB = BlockBar;   // the threads that walk: the whole CTA
  const lantern_accept_cfg& cfg = P.cfg;
  const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const int L = cfg.n_paths, D = cfg.depth, T = cfg.n_rows, V = cfg.vocab;
  const int ncols = cfg.ncols, col0 = cfg.col0, col1 = cfg.col0 + cfg.ncols;
  const int off = cfg.tok_offset;

  // ---- carve shared memory ----
  WalkSmem S;
  size_t o = 0;
  // (making S.p provably shared memory in the lazy form - LDS / STS instead of generic loads - was measured 2.7 us
  // slower per step, same box, three alternations: the generic form stays)
  if (P.p_spill) {   // very wide windows: the probability vector lives in global memory (L2), one slice per prompt
    S.p = P.p_spill + (size_t)b * (size_t)((ncols + 3) & ~3);
  } else {
    S.p = reinterpret_cast<float*>(smem_raw + o);          o += (size_t)((ncols + 3) & ~3) * 4;
  }
  S.nbmask = reinterpret_cast<unsigned*>(smem_raw + o);     o += (size_t)(((ncols + 31) >> 5) + 3 & ~3) * 4;
  S.dscr = reinterpret_cast<double*>(smem_raw + o);         o += 34 * 8;
  S.ri = reinterpret_cast<int*>(smem_raw + o);              o += (size_t)((L * D + 3) & ~3) * 4;
  S.tok = reinterpret_cast<int*>(smem_raw + o);             o += (size_t)((T + 3) & ~3) * 4;
  S.tried = reinterpret_cast<int*>(smem_raw + o);           o += (size_t)((L + 3) & ~3) * 4;
  S.kid_j = reinterpret_cast<int*>(smem_raw + o);           o += (size_t)((L + 3) & ~3) * 4;
  S.kid_node = reinterpret_cast<int*>(smem_raw + o);        o += (size_t)((L + 3) & ~3) * 4;
  S.csv = reinterpret_cast<float*>(smem_raw + o);           o += (size_t)kWalkThreads * 4;
  S.uni = reinterpret_cast<float*>(smem_raw + o);           o += (size_t)((T + 4) & ~3) * 4;
  S.sib = reinterpret_cast<int*>(smem_raw + o);             o += kMaxSib * 4;
  S.fscr = reinterpret_cast<float*>(smem_raw + o);          o += 36 * 4;
  S.iscr = reinterpret_cast<int*>(smem_raw + o);            o += 40 * 4;
  S.pid = reinterpret_cast<int*>(smem_raw + o);                 o += (size_t)((L + 3) & ~3) * 4;
  S.maxcp = reinterpret_cast<int*>(smem_raw + o);               o += (size_t)((L + 3) & ~3) * 4;
  S.ctok = reinterpret_cast<int*>(smem_raw + o);                o += (size_t)((L * D + 3) & ~3) * 4;
  S.csyn = reinterpret_cast<int*>(smem_raw + o);                o += (size_t)((L * D + 3) & ~3) * 4;
  S.kid_syn = reinterpret_cast<int*>(smem_raw + o);             o += (size_t)((L + 3) & ~3) * 4;
  o = (o + 7) & ~size_t(7);
  double* rsum = reinterpret_cast<double*>(smem_raw + o);       o += 2 * 32 * 8;   // removed-mass partials, two parities
  float* lazy_park = reinterpret_cast<float*>(smem_raw + o);   // [LNE][kLazyThreads] (lazy modes only)
  __shared__ SelectSmem lazy_sm;
  __shared__ LazySmem lazy_lz;
  float z_run = P.z_guess, win_run = P.win_sd_first;

#ifdef LANTERN_WALK_TRACE
  if (threadIdx.x == 0) tr_n = 0;
#endif
  TR(1);
  const int* ri_g = P.in.retrieve + (cfg.retrieve_shared ? 0 : (size_t)b * L * D);
  const int* tok_g = P.in.tree_tokens + (size_t)b * T;
  auto cand = [&](int j, int i) -> int { return S.ctok[j * D + i]; };   // token of row j at level i, -1 = padding
  // Rows still matching the accepted prefix (the reference's `is_eq`, ea_model_llamagen.py:721): every row that starts
  // with row 0's root token; an acceptance at level i keeps the rows whose token at level i is the accepted one.
  int* member = S.pid;   // [L]
  // Staging of the prompt's tree.  In the lazy form it runs inside the load shadow of the root's logits row (as part of
  // the first level's hook, see the level loop): the walk is a latency chain and this is ~1.5 us of it.
  auto prologue = [&]() {
    for (int i = tid; i < L * D; i += NT) S.ri[i] = ri_g[i];
    for (int i = tid; i < T; i += NT) S.tok[i] = tok_g[i];
    // at most one draw per tree node plus the bonus token (SURVEY.md Appendix A "uniform budget")
    for (int i = tid; i < T + 1; i += NT) {
      float uv;
      if (P.in.uniforms) uv = i < cfg.n_uniforms ? P.in.uniforms[(size_t)b * cfg.n_uniforms + i] : 0.f;
      else uv = philox_uniform(cfg.philox_seed, cfg.philox_step, (uint32_t)b, (uint32_t)i);
      S.uni[i] = uv;
    }
    TR(21);
    WB::sync();
    TR(22);
    for (int i = tid; i < L * D; i += NT) {
      const int n = S.ri[i];
      const int x = n >= 0 ? S.tok[n] : -1;
      S.ctok[i] = x;
      int syn = 0;
      if (P.lumina)
        for (int q = 0; q < cfg.n_syntax; ++q) syn |= (cfg.syntax_tokens[q] == x) ? 1 : 0;
      S.csyn[i] = syn;
    }
    WB::sync();
    TR(23);
    {
      const int root_tok = cand(0, 0);
      for (int j = tid; j < L; j += NT) member[j] = cand(j, 0) == root_tok ? 1 : 0;
    }
    // The reference lists the distinct child tokens of the current node in row order (`candidates_set`,
    // ea_model_llamagen.py:728-739).  The rows in play at level l all carry the accepted tokens at levels < l, so row
    // j repeats an earlier row's token at level l exactly when an earlier row shares j's token prefix through level l:
    // a property of the tree alone.  maxcp[j] = the longest token prefix row j shares with any earlier row.
    for (int j = tid >> 5; j < L; j += NT >> 5) {   // one warp per row, the lanes take the earlier rows
      int longest = 0;
      for (int jj = tid & 31; jj < j; jj += 32) {
        int cp = D;
        for (int l = D - 1; l >= 0; --l)
          if (S.ctok[jj * D + l] != S.ctok[j * D + l]) cp = l;
        longest = max(longest, cp);
      }
      longest = __reduce_max_sync(0xffffffffu, longest);
      if ((tid & 31) == 0) S.maxcp[j] = longest;
    }
    TR(24);
    WB::sync();
    TR(25);
  };
  const bool defer_prologue = LNE > 0 && D > 1;
  if (!defer_prologue) prologue();

  // ---- distribution state: window S.p + one explicit out-of-window token + uniform remainder ----
  // The residual is stored unnormalised: probability = stored value * scale.  A rejection then only zeroes
  // entries and re-derives scale = 1 / sum instead of dividing the whole vector (reference: gtp /= gtp.sum()).
  int extra_tok = -1;
  float p_extra = 0.f, p_out = 0.f, scale = 1.0f;
  auto prob_of = [&](int tkn) -> float {
    if (tkn >= col0 && tkn < col1) return S.p[tkn - col0];
    if (tkn == extra_tok) return p_extra;
    return p_out;
  };
  int rows_read = 0;   // logits rows materialised into S.p (reported in the flags word; one-hot rows read nothing)
  // CFG-mixed logit of one column of a tree row (any column of the vocabulary, not only the live window)
  auto logit_at = [&](int node, int col) -> float {
    const int64_t o = (int64_t)b * cfg.item_stride + (int64_t)node * cfg.row_stride + col;
    const float c = Elem<DT>::load1(P.in.logits_cond, o);
    const float u = P.mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, o) : 0.f;
    MixParams m = P.mix;
    m.do_temp = 0;
    return mix_temper(c, u, m);
  };
  // which node's distribution S.p holds, and whether a rejection has modified it since (the fresh tail can reuse it)
  int dist_node = -1;
  bool dist_dirty = false, dist_raw = false;
  // fp64 sum of the window S.p, kept up to date across zeroing rejections (a rejection then only reduces the mass it
  // removed instead of re-summing the whole vector); invalid after a fresh distribution or a static-tree subtraction
  double win_tot = 0.0;
  bool win_tot_valid = false;
  unsigned rej_parity = 0;
  // `hook` (lazy form only): work for one warp that does not depend on the row - it runs while the row's loads are in
  // flight and is published by the barriers below
  auto set_distribution = [&](int node, bool raw, auto&& hook, auto&& post) {
    const long long row = (long long)b * T + node;
    RowStats st;
    if (LNE > 0) {   // lazy mode: only the row class is needed up front
      st.kind = P.in.row_kinds ? (int)P.in.row_kinds[row] : LANTERN_ROW_IMAGE;
      // no barrier here: S.p is rewritten only behind the barriers of lazy_probs, which every earlier reader has passed
      if (st.kind != LANTERN_ROW_IMAGE) { hook(); WB::sync(); post(); }
    } else {
      st = P.stats[row];
      WB::sync();
    }
    extra_tok = -1; p_extra = 0.f; p_out = 0.f; scale = 1.0f;
    dist_node = node; dist_dirty = false; dist_raw = raw;
    win_tot_valid = false;
    int kind = st.kind;
    if (kind == LANTERN_ROW_IMAGE) {
      bool empty;
      if (LNE > 0) {
        ++rows_read;
        empty = lazy_probs<DT, (LNE > 0 ? LNE : 8)>(P, b, node, raw, S.p, lazy_park, lazy_lz, lazy_sm, z_run, win_run, hook, post);
      } else {
        ++rows_read;
        if (raw) st = raw_row_stats<DT, VEC>(P, b, node, S.fscr, S.dscr);
        empty = st.mx == -INFINITY;
        if (!empty) load_probs<DT, VEC>(P, b, node, st, S.p, raw);
      }
      if (empty) {
        // Every live column is -inf: the caller handed over rows that MultiModalLogitsProcessor already turned into
        // one-hot newline / end-of-image rows (the reference's gathered [L, D, V] form, ea_model_lumina_mgpt.py:71-84,
        // where no row_kinds travel).  The finite syntax column says which; an all -inf row keeps a zero window.
        kind = -1;
        if (cfg.eoi_token >= 0 && cfg.eoi_token < V && logit_at(node, cfg.eoi_token) > -INFINITY) kind = LANTERN_ROW_EOI;
        else if (cfg.newline_token >= 0 && cfg.newline_token < V && logit_at(node, cfg.newline_token) > -INFINITY)
          kind = LANTERN_ROW_NEWLINE;
      }
    }
    if (kind != LANTERN_ROW_IMAGE) {
      for (int e = tid; e < ncols; e += NT) S.p[e] = 0.f;
      if (kind >= 0) {
        extra_tok = kind == LANTERN_ROW_NEWLINE ? cfg.newline_token : cfg.eoi_token;
        p_extra = 1.0f;
        if (extra_tok >= col0 && extra_tok < col1) {   // degenerate configs: keep it inside the window
          WB::sync();
          if (tid == 0) S.p[extra_tok - col0] = 1.0f;
          extra_tok = -1; p_extra = 0.f;
        }
      }
    }
    WB::sync();
  };
  auto uniform = [&](int d) -> float { return S.uni[min(d, T)]; };
  // speculative L2 prefetch of a child's logits row (a wasted prefetch costs 64 KiB of otherwise idle HBM bandwidth)
  auto prefetch_row = [&](int cnode, int lvl) {
    if (!P.prefetch_rows || cnode < 0 || lvl + 1 >= D) return;
    constexpr size_t eb = Elem<DT>::kBytes;
    const int64_t nbase = (int64_t)b * cfg.item_stride + (int64_t)cnode * cfg.row_stride + col0;
    const uintptr_t a0 = (reinterpret_cast<uintptr_t>(P.in.logits_cond) + nbase * eb + 15) & ~uintptr_t(15);
    const uintptr_t a1 = (reinterpret_cast<uintptr_t>(P.in.logits_cond) + (nbase + ncols) * eb) & ~uintptr_t(15);
    if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)));
    if (P.mix.has_uncond) {
      const uintptr_t b0 = (reinterpret_cast<uintptr_t>(P.in.logits_uncond) + nbase * eb + 15) & ~uintptr_t(15);
      const uintptr_t b1 = (reinterpret_cast<uintptr_t>(P.in.logits_uncond) + (nbase + ncols) * eb) & ~uintptr_t(15);
      if (b1 > b0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b0), "r"((unsigned)(b1 - b0)));
    }
  };
  int accept_length = 1, best = 0, draws = 0, out_flags = 0;
  bool adjust = false;
  const int kk = cfg.lantern ? min(cfg.lantern_k, cfg.table_cols) : 0;
  const int kk1 = cfg.lantern ? min(cfg.lantern_k + 1, cfg.table_cols) : 0;

  TR(2);
  for (int lvl = 1; lvl < D; ++lvl) {
    if (lvl != accept_length) break;
    adjust = false;
    TR(10 + lvl);
    // fi, the first row still matching the accepted prefix (`member` was last written behind a barrier): every warp
    // finds it for itself
    const bool staging = defer_prologue && lvl == 1;   // the tree is not in shared memory yet: row 0 is always in play
    int fi = 0;
    for (int jb = 0; jb < L && !staging; jb += 32) {
      const int j = jb + (tid & 31);
      const unsigned mm = __ballot_sync(0xffffffffu, j < L && member[j] != 0);
      if (mm) { fi = jb + __ffs(mm) - 1; break; }
    }
    int node = staging ? __ldg(ri_g) : S.ri[fi * D + lvl - 1];
    if (node < 0) node += T;
    TR(35);
    // Warp 0 lists the distinct children of the current node in row order (the reference's `candidates_set`
    // dedup, ea_model_llamagen.py:728-739).  In the lazy form this runs while the node's logits row is in flight.
    auto list_children = [&]() {
      if (tid >= 32) return;
      int n = 0;
#pragma unroll 2
      for (int jb = 0; jb < L; jb += 32) {
        const int j = jb + tid;
        const bool mem = j < L && member[j] != 0;
        int x = -1, cn = -1;
        if (mem) {
          cn = S.ri[j * D + lvl];
          x = S.ctok[j * D + lvl];
        }
        const bool first = x != -1 && S.maxcp[j] <= lvl;   // no earlier row in play carries the same token here
        const unsigned mf = __ballot_sync(0xffffffffu, first);
        if (first) {
          const int pos = n + __popc(mf & ((1u << tid) - 1u));
          S.tried[pos] = x; S.kid_j[pos] = j; S.kid_node[pos] = cn;
          S.kid_syn[pos] = S.csyn[j * D + lvl];
        }
        n += __popc(mf);
        __syncwarp();
      }
      if (tid == 0) S.iscr[38] = n;
      TR(36);
    };
    // Once the list is published: pull the children's neighbour-table rows towards L2, and the logits rows of the first
    // two children (the next row the walk needs if one of them is accepted; later children are prefetched when they are
    // tried).  A bulk prefetch keeps its issuing lane busy for ~100 cycles, so the children are dealt out to the warps.
    auto prefetch_children = [&]() {
      if ((tid & 31) != 0) return;
      const int n = S.iscr[38], w = tid >> 5, nw = NT >> 5;
      if (w == nw - 1 && n > 0) prefetch_row(S.kid_node[0], lvl);
      if (w == nw - 2 && n > 1) prefetch_row(S.kid_node[1], lvl);
      if (!cfg.lantern) return;
      for (int c = w; c < n; c += nw) {
        const int xk = S.tried[c] - off;
        if (xk >= 0 && xk < ncols) {
          const uintptr_t a0 = (reinterpret_cast<uintptr_t>(P.in.nbr_table + (size_t)xk * cfg.table_cols) + 15) & ~uintptr_t(15);
          const uintptr_t a1 = reinterpret_cast<uintptr_t>(P.in.nbr_table + (size_t)xk * cfg.table_cols + kk1) & ~uintptr_t(15);
          if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)));
        }
      }
    };
    if (LNE > 0) {
      set_distribution(node, false, [&]() {
        if (staging) prologue();
        list_children();
      }, prefetch_children);
    } else {
      list_children();
      WB::sync();
      prefetch_children();
      set_distribution(node, false, [] {}, [] {});
    }
    const int n_kids = S.iscr[38];
    TR(30 + lvl);

    bool accepted = false;
    for (int c = 0; c < n_kids && !accepted; ++c) {
      const int j = S.kid_j[c], cnode = S.kid_node[c], x = S.tried[c];
      if (c >= 2 && tid == 0) prefetch_row(cnode, lvl);   // the first two went out when the children were listed
      TR(100 + c);
      const float r = uniform(draws++);
      float px = __fmul_rn(prob_of(x), scale);
      // only image tokens own a neighbour-table row; anything else (masked to probability 0 by every family's
      // tree_decoding) is tested on its own probability
      bool relaxable = x >= col0 && x < col1;
      if (P.lumina) {
        if (S.kid_syn[c]) { px = 1.0f; relaxable = false; }
        else if (!relaxable) px = 0.0f;
      }
      int idx = -1;
      const int* nb_row = nullptr;
      float qx = 1.0f;
      if (cfg.static_tree) {
        qx = P.in.node_q[(size_t)b * T + cnode];
        if (qx <= 0.f) continue;   // the draw above is consumed first, like the reference (:636-638)
      }
      // The relaxation only ever adds mass (px += cs[idx] >= 0): a candidate that passes on its own probability is
      // accepted whatever the neighbours hold, so their gather and scan are skipped for it.
      TR(700 + c);
      // (qx is exactly 1 on a dynamic tree and x / 1 = x: the IEEE division sequences are skipped there - on this
      // one-shot instruction stream every instruction that is not executed counts)
      auto over_q = [&](float v) -> float { return cfg.static_tree ? __fdiv_rn(v, qx) : v; };
      const bool sure = r <= over_q(px);
      bool scan = cfg.lantern && relaxable && !sure;
      float bound = 0.f;
      if (scan) {
        nb_row = P.in.nbr_table + (size_t)(x - off) * cfg.table_cols;
        bound = cfg.lantern_delta > 1.0f ? __fmul_rn(cfg.lantern_delta_m1, px) : cfg.lantern_delta;
        // The added mass never exceeds the bound (cs[idx] <= bound, rounding is monotone), so a draw above
        // (px + bound) / qx is a rejection whatever the neighbours hold.  The residual update then only needs to know
        // whether any neighbour was aggregated (idx != -1), i.e. whether the first prefix sum is within the bound.
        if (r > over_q(__fadd_rn(px, bound))) {
          scan = false;
          const float cs0 = (float)((double)prob_of(__ldg(nb_row) + off) * (double)scale);
          idx = (kk > 0 && cs0 <= bound) ? 0 : -1;
        }
      }
      if (scan) {
        // Prefix sums of the neighbour masses in fp64 (rounded to fp32 per prefix, like the CPU cumsum), in chunks of
        // EPT * NT neighbours: each thread owns EPT <= 4 consecutive neighbours, so the usual k = 1000 is one block scan
        // for both CTA sizes (1024 threads x 1, 256 threads x 4).
        const int ept = min(4, (kk + NT - 1) / NT);
        const int span = NT * ept;
        double carry = 0.0;
        int n_ok = 0;
        for (int base_t = 0; base_t < kk; base_t += span) {
          const int chunk = min(span, kk - base_t);
          if (tid == 0) S.iscr[39] = chunk;             // index (within the chunk) of the first sum above the bound
          const int t0 = base_t + tid * ept;
          double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < ept && t0 + q < kk) v[q] = (double)prob_of(__ldg(nb_row + t0 + q) + off);
          const double mine = (v[0] + v[1]) + (v[2] + v[3]);
          double total;
          double run = carry + block_scan_incl<WB>(mine, S.dscr, &total) - mine;   // prefix before the thread's first neighbour
          int fail = 0x7fffffff;                         // sums are non-decreasing: the failures form a suffix
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < ept) {
              run += v[q];
              const float cs = (float)(run * (double)scale);
              S.csv[tid * ept + q] = cs;
              if (fail == 0x7fffffff && t0 + q < kk && !(cs <= bound)) fail = tid * ept + q;
            }
          }
          const int wfail = __reduce_min_sync(0xffffffffu, fail);
          if ((tid & 31) == 0 && wfail != 0x7fffffff) atomicMin(&S.iscr[39], wfail);
          WB::sync();
          const int cnt = S.iscr[39];
          if (cnt > 0) S.fscr[35] = S.csv[cnt - 1];      // every thread stores the same value
          n_ok += cnt;
          carry += total;
          WB::sync();
          if (cnt < chunk) break;
        }
        if (n_ok > 0) {
          idx = n_ok - 1;
          px = __fadd_rn(px, S.fscr[35]);
        }
      }
      TR(800 + c);
      const float acp = over_q(px);
      if (r <= acp) {
        ++accept_length;
        best = j;
        accepted = true;
        for (int jj = tid; jj < L; jj += NT)     // rows that continue with token x
          if (member[jj] && cand(jj, lvl) != x) member[jj] = 0;
        WB::sync();
        break;
      }
      TR(200 + c);
      // ---------------- rejection: residual distribution ----------------
      dist_dirty = true;
      // every thread must have read this candidate's probabilities (px, first prefix sum) before any entry is zeroed:
      // the shortcuts above can reach this point without passing a barrier
      WB::sync();
      TR(300 + c);
      const bool zero_nb = cfg.lantern && relaxable && idx != -1;
      double removed = 0.0;
      if (cfg.static_tree) {
        const float* q = P.in.draft_op + ((size_t)b * cfg.n_q_rows + P.in.node_qrow[cnode]) * (size_t)V;
        const int s0 = P.in.sib_off[cnode], s1 = min(P.in.sib_off[cnode + 1], s0 + kMaxSib);
        const int* sib_tok = P.in.sib_tokens + (size_t)b * P.in.sib_tokens_stride;
        WB::sync();
        for (int s = s0 + tid; s < s1; s += NT) S.sib[s - s0] = sib_tok[P.in.sib_idx[s]];
        WB::sync();
        const int nsib = s1 - s0;
        auto is_sib = [&](int tkn) -> bool {
          bool hit = false;
          for (int s = 0; s < nsib; ++s) hit |= (S.sib[s] == tkn);
          return hit;
        };
        float qsum = 1.0f;
        if (s1 > s0) {   // q[earlier siblings] = 0; q /= q.sum()
          double part = 0.0;
          for (int v = tid; v < V; v += NT) part += is_sib(v) ? 0.0 : (double)q[v];
          qsum = (float)group_reduce<WB>(part, OpSum(), 0.0, S.dscr);
        }
        if (zero_nb) {
          if (P.static_zero_q) {
            for (int w = tid; w < ((ncols + 31) >> 5); w += NT) S.nbmask[w] = 0u;
            WB::sync();
            for (int tt = tid; tt < kk1; tt += NT) {
              const int nbt = __ldg(nb_row + tt) + off - col0;
              if (nbt >= 0 && nbt < ncols) atomicOr(&S.nbmask[nbt >> 5], 1u << (nbt & 31));
            }
          } else {
            for (int tt = tid; tt < kk1; tt += NT) {
              const int nbt = __ldg(nb_row + tt) + off - col0;
              if (nbt >= 0 && nbt < ncols) S.p[nbt] = 0.f;
            }
          }
        }
        WB::sync();
        const bool use_mask = zero_nb && P.static_zero_q;
        for (int e = tid; e < ncols; e += NT) {
          const int tkn = e + col0;
          float qv = is_sib(tkn) ? 0.f : q[tkn];
          if (s1 > s0) qv = __fdiv_rn(qv, qsum);
          if (use_mask && ((S.nbmask[e >> 5] >> (e & 31)) & 1u)) qv = 0.f;
          S.p[e] = fmaxf(__fsub_rn(__fmul_rn(S.p[e], scale), qv), 0.f);
        }
        if (extra_tok >= 0) {
          float qv = is_sib(extra_tok) ? 0.f : q[extra_tok];
          if (s1 > s0) qv = __fdiv_rn(qv, qsum);
          p_extra = fmaxf(__fsub_rn(__fmul_rn(p_extra, scale), qv), 0.f);
        }
        p_out = __fmul_rn(p_out, scale);
        scale = 1.0f;   // the subtraction pass stored normalised values
      } else {
        // gtp[x] = 0 and, if the candidate was relaxed, gtp[its k+1 nearest] = 0; atomicExch hands every entry's old
        // value to exactly one thread, whatever the table holds
        if (zero_nb) {
          for (int t0 = tid; t0 < kk1; t0 += 4 * NT) {   // four table entries in flight per thread
            int nbt[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) nbt[q] = t0 + q * NT < kk1 ? __ldg(nb_row + t0 + q * NT) + off - col0 : -1;
            float old[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) old[q] = (nbt[q] >= 0 && nbt[q] < ncols) ? atomicExch(&S.p[nbt[q]], 0.f) : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) removed += (double)old[q];
          }
        }
        if (x >= col0 && x < col1) { if (tid == 0) removed += (double)atomicExch(&S.p[x - col0], 0.f); }
        else if (x == extra_tok) p_extra = 0.f;
      }
      TR(400 + c);
      double tot;
      bool resum = cfg.static_tree || !win_tot_valid;
      if (!resum) {
        // one barrier: per-warp partials of the removed mass (double-buffered by rejection parity), summed by everyone
        removed = warp_reduce(removed, OpSum());
        double* rs = rsum + (rej_parity & 1) * 32;
        ++rej_parity;
        if ((tid & 31) == 0) rs[tid >> 5] = removed;
        WB::sync();
        double rem = 0.0;
        for (int w = 0; w < (NT >> 5); ++w) rem += rs[w];
        win_tot -= rem;
        if (win_tot < 1e-6) resum = true;     // nearly everything is gone: take the exact sum (the == 0 test below)
      }
      if (resum) {
        WB::sync();
        // fp64 from the first add: later rejections subtract their exact removed mass from this sum, so an fp32
        // partial (1e-7 of the whole window) would be amplified by whole / remaining mass in the residual's scale
        double part = 0.0;
        for (int e = tid; e < ncols; e += NT) part += (double)S.p[e];
        win_tot = group_reduce<WB>(part, OpSum(), 0.0, S.dscr);
        win_tot_valid = !cfg.static_tree;
      }
      TR(500 + c);
      tot = win_tot + (double)p_extra + (double)p_out * (double)(V - ncols - (extra_tok >= 0 ? 1 : 0));
      if ((float)tot == 0.f) {   // gtp = ones_like(gtp)
        for (int e = tid; e < ncols; e += NT) S.p[e] = 1.0f;
        if (extra_tok >= 0) p_extra = 1.0f;
        p_out = 1.0f;
        tot = (double)V;
        win_tot = (double)ncols;
        out_flags |= LANTERN_OUT_UNIFORM_FALLBACK;
        WB::sync();
      }
      scale = __fdiv_rn(1.0f, (float)tot);   // gtp /= gtp.sum(), applied lazily
      adjust = true;
      TR(600 + c);
    }
  }

  TR(3);
  // ---------------- tail distribution ----------------
  const bool residual_tail = adjust && (accept_length != D);
  if (residual_tail) {
    out_flags |= LANTERN_OUT_RESIDUAL_TAIL;
  } else {
    int node = S.ri[best * D + accept_length - 1];
    if (node < 0) node += T;
    // the distribution of the last accepted node is often still in place, untouched (it was made for a level that had
    // no children to try): the reference recomputes the same softmax (:784-786)
    if (!(dist_node == node && !dist_dirty && dist_raw == (P.tail_raw != 0))) set_distribution(node, P.tail_raw != 0, [] {}, [] {});
  }
  WB::sync();

  TR(4);
  // ---------------- bonus token: inverse CDF, fp64, index order ----------------
  const float u = (cfg.bonus_uniform_last && P.in.uniforms)
                      ? P.in.uniforms[(size_t)b * cfg.n_uniforms + cfg.n_uniforms - 1]
                      : uniform(draws);
  ++draws;
  const int per = (((ncols + NT - 1) / NT) + 3) & ~3;   // contiguous, float4-aligned chunk per thread
  const int i0 = min(tid * per, ncols), i1 = min(i0 + per, ncols);
  double loc = 0.0;
  int last_nz = -1;
  for (int i = i0; i + 4 <= i1; i += 4) {
    const float4 v = *reinterpret_cast<const float4*>(S.p + i);
    loc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    if (v.x > 0.f) last_nz = i;
    if (v.y > 0.f) last_nz = i + 1;
    if (v.z > 0.f) last_nz = i + 2;
    if (v.w > 0.f) last_nz = i + 3;
  }
  for (int i = i0 + ((i1 - i0) & ~3); i < i1; ++i) {
    loc += (double)S.p[i];
    if (S.p[i] > 0.f) last_nz = i;
  }
  double wtot;
  if (tid == 0) { S.iscr[36] = 0x7fffffff; S.iscr[37] = -1; }   // published by the barriers of the scan
  const double incl = block_scan_incl<WB>(loc, S.dscr, &wtot);
  const double massA = (double)p_out * (double)col0;
  const int n_suffix_uniform = V - col1 - ((extra_tok >= col1) ? 1 : 0);
  const double massB = (double)p_extra + (double)p_out * (double)n_suffix_uniform;
  const double total = massA + wtot + massB;
  const double target = (double)u * total;
  {
    // the thread whose chunk holds the target hands the chunk to its warp: 32 entries per step, one warp scan each
    const double run0 = massA + incl - loc;
    const unsigned om = __ballot_sync(0xffffffffu, target >= run0 && target < run0 + loc);
    if (om) {
      const int src = __ffs(om) - 1, ln = tid & 31;
      const int oi0 = __shfl_sync(0xffffffffu, i0, src), oi1 = __shfl_sync(0xffffffffu, i1, src);
      double carry = __shfl_sync(0xffffffffu, run0, src);
      int found = -1;
      for (int c0 = oi0; c0 < oi1 && found < 0; c0 += 32) {
        const int i = c0 + ln;
        double v = i < oi1 ? (double)S.p[i] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double n = __shfl_up_sync(0xffffffffu, v, o);
          if (ln >= o) v += n;
        }
        const unsigned hm = __ballot_sync(0xffffffffu, i < oi1 && carry + v > target);
        if (hm) found = c0 + __ffs(hm) - 1;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      if (found >= 0 && ln == 0) atomicMin(&S.iscr[36], found + col0);
    }
    const int wl = __reduce_max_sync(0xffffffffu, last_nz);   // one shared-memory atomic per warp, not per thread
    if ((tid & 31) == 0 && wl >= 0) atomicMax(&S.iscr[37], wl + col0);
  }
  WB::sync();
  int token = S.iscr[36];
  if (token == 0x7fffffff) {
    if (target < massA && p_out > 0.f) {
      token = min(col0 - 1, (int)(target / (double)p_out));
    } else {
      const double rem = target - massA - wtot;
      if (p_out > 0.f && rem >= 0.0) {
        token = min(V - 1, col1 + (int)(rem / (double)p_out));
      } else if (p_extra > 0.f) {
        token = extra_tok;
      } else {
        token = S.iscr[37] >= 0 ? S.iscr[37] : col0;   // numerical fall-through: last non-zero entry
      }
    }
  }

  TR(5);
  // ---------------- outputs ----------------
  const int a = accept_length - 1;
  if (tid == 0) {
    P.out.accept_length[b] = a;
    P.out.best_candidate[b] = best;
    P.out.token[b] = token;
    if (P.out.n_draws) P.out.n_draws[b] = draws;
    if (P.out.flags) P.out.flags[b] = out_flags | (min(rows_read, 255) << LANTERN_OUT_ROWS_READ_SHIFT);
  }
  for (int i = tid; i < D; i += NT) {
    if (P.out.path_tokens) P.out.path_tokens[(size_t)b * D + i] = i <= a ? cand(best, i) : -1;
    if (P.out.select_indices) P.out.select_indices[(size_t)b * D + i] = i <= a ? S.ri[best * D + i] : -1;
  }
  if (P.out.sample_p) {
    float* sp = P.out.sample_p + (size_t)b * V;
    for (int v = tid; v < V; v += NT) sp[v] = __fmul_rn(prob_of(v), scale);
  }
  TR(6);
#ifdef LANTERN_WALK_TRACE
  if (tid == 0 && b < kTraceBlocks) {
    g_trace[b][0] = tr_n;
    g_trace[b][1] = draws;
    for (int i = 0; i < 2 * tr_n; ++i) g_trace[b][2 + i] = tr_buf[i];
  }
#endif
}

// ----------------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------------
static size_t walk_smem_bytes(const lantern_accept_cfg& c, int lazy_ne, bool spill = false) {
  size_t o = (size_t)lazy_ne * kLazyThreads * 4;
  if (!spill) o += (size_t)((c.ncols + 3) & ~3) * 4;
  o += (size_t)((((c.ncols + 31) >> 5) + 3) & ~3) * 4;
  o += 34 * 8;
  o += (size_t)((c.n_paths * c.depth + 3) & ~3) * 4;
  o += (size_t)((c.n_rows + 3) & ~3) * 4;
  o += 3 * (size_t)((c.n_paths + 3) & ~3) * 4;
  o += (size_t)kWalkThreads * 4;
  o += (size_t)((c.n_rows + 4) & ~3) * 4;
  o += kMaxSib * 4 + 36 * 4 + 40 * 4;
  o += 3 * (size_t)((c.n_paths + 3) & ~3) * 4 + 16;
  o += 2 * (size_t)((c.n_paths * c.depth + 3) & ~3) * 4;
  o += 2 * 32 * 8 + 8;
  return o;
}

template <int DT, bool VEC>
static int launch_all(const AcceptParams& P_in, cudaStream_t stream, int phases) {
  AcceptParams P = P_in;
  P.prefetch_rows = (phases & 16) ? 0 : 1;   // bit 4: logits live behind PCIe (in-place host rows): no speculative reads
  const lantern_accept_cfg& c = P.cfg;
  const long long rows = (long long)c.n_items * c.n_rows;
  if (phases & 8) {
    // automatic policy (measured, profiles/sweep_r1.md): streaming every tree row pays off while the step is
    // latency-bound; from ~2K rows on, computing the statistics of the visited rows inside the walk is faster
    const int ne = c.ncols / kLazyThreads;
    const bool lazy_ok = VEC && !P.do_topp && c.ncols % (4 * kLazyThreads) == 0 && (c.ncols == 2048 || c.ncols == 4096 || c.ncols == 8192 || c.ncols == 16384);
    // (round 2, rewritten lazy walk: from ~48 tree rows per prompt on it is no slower at any batch size, cold or warm -
    // profiles/r2/sweep_r2.md, profiles/r2/auto_policy_ab.txt; tiny trees visit most of their rows and keep the
    // streamed form, which spreads them over the SMs)
    phases = (lazy_ok && (rows >= 2048 || c.n_rows >= 48)) ? 6 : 3;
  }
  const int nquads = (c.ncols + 3) / 4;
  // Thread/element split.  Generic mode: 256 threads up to 8192 columns, 512 beyond.  Fast (TMA-staged) modes:
  // 32 elements per thread (256 threads for 8192 columns, 512 for 16384): measured faster than 16 per thread
  // because the per-thread fixed cost of the reductions is amortised over twice the elements.
  constexpr int EB = Elem<DT>::kBytes;
  const bool starts_aligned = ((size_t)c.col0 * EB) % 16 == 0 && ((size_t)c.row_stride * EB) % 16 == 0 &&
                              ((size_t)c.item_stride * EB) % 16 == 0 &&
                              reinterpret_cast<uintptr_t>(P.in.logits_cond) % 16 == 0 &&
                              reinterpret_cast<uintptr_t>(P.in.logits_uncond) % 16 == 0;
  const bool stageable = VEC && (starts_aligned || c.col0 + c.ncols + 8 <= c.row_stride);
  int nt = c.ncols <= 8192 ? 256 : 512, nq_inst = 1;
  bool full = false;
  static const int fast_shapes[][3] = {{2048, 256, 2}, {4096, 256, 4}, {8192, 256, 8}, {16384, 512, 8}, {32768, 1024, 8}};
  const char* env_nt = getenv("LANTERN_STATS_NT");
  for (auto& fsz : fast_shapes) {
    if (stageable && c.ncols == fsz[0]) { nt = fsz[1]; nq_inst = fsz[2]; full = true; }
  }
  if (full && env_nt && c.ncols % (4 * atoi(env_nt)) == 0) {   // tuning knob for experiments
    nt = atoi(env_nt);
    nq_inst = c.ncols / (4 * nt);
  }
  if (!full) {
    const int nq = (nquads + nt - 1) / nt;
    while (nq_inst < nq) nq_inst <<= 1;
    if (nt == 512 && nq_inst < 8) nq_inst = 8;
  }
  const int mode = full ? (P.mix.has_uncond ? 1 : 2) : 0;
  const size_t stage_bytes = mode ? (((size_t)c.ncols * EB + 32 + 127) & ~size_t(127)) : 0;
  // streaming form (stats_stream.cuh): NT main threads + one select warp; LANTERN_STATS_OLD=1 keeps the round-1 kernel
  const bool use_stream = mode && nt <= 512 && c.ncols <= 16384 && !getenv("LANTERN_STATS_OLD");
  const int ps_slots = std::min(nq_inst * 4, 20);
  const size_t park_bytes = (use_stream ? (size_t)ps_slots * nt * sizeof(float) + 2 * (size_t)(nt / 32) * kSegCap * sizeof(float)
                                    : (size_t)nq_inst * 4 * nt * sizeof(float)) +
                            (mode == 1 ? 2 : (mode == 2 ? 1 : 0)) * stage_bytes;
  const bool wide = !full && c.ncols > 16 * 4 * 512;   // multi-pass kernel, no register-resident row
  if (!wide && park_bytes + 6 * 1024 > 227 * 1024) {
    set_error("row statistics kernel needs %zu bytes of shared memory", park_bytes);
    return LANTERN_E_UNSUPPORTED;
  }
  int per_sm = mode ? (nt <= 512 ? 2 : 1) : (nq_inst <= 8 ? 1024 / nt : 512 / nt);
  per_sm = std::max(1, std::min<int>(per_sm, (int)((227 * 1024) / (park_bytes + 5 * 1024))));
  const int grid = (int)std::min<long long>(rows, (long long)kNumSMs * per_sm);
#define LAUNCH_STATS(NT, NQ)                                                                              \
  do {                                                                                                    \
    auto kk = row_stats_kernel<DT, NT, NQ, VEC>;                                                          \
    if (park_bytes > 48 * 1024)                                                                           \
      LANTERN_CUDA(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)park_bytes)); \
    kk<<<grid, NT, park_bytes, stream>>>(P);                                                              \
  } while (0)
#define LAUNCH_FAST_MODE(NT, NQ, M)                                                                        \
  do {                                                                                                    \
    auto kk = row_stats_fast_kernel<DT, NT, NQ, M>;                                                       \
    if (park_bytes > 48 * 1024)                                                                           \
      LANTERN_CUDA(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)park_bytes)); \
    kk<<<grid, NT, park_bytes, stream>>>(P);                                                              \
  } while (0)
#define LAUNCH_STREAM_MODE(NT, NQ, M, TT)                                                                  \
  do {                                                                                                    \
    auto kk = row_stats_stream_kernel<DT, NT, NQ, M, TT>;                                                 \
    if (park_bytes > 48 * 1024)                                                                           \
      LANTERN_CUDA(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)park_bytes)); \
    kk<<<grid, NT + 64, park_bytes, stream>>>(P);                                                        \
  } while (0)
#define LAUNCH_STREAM(NT, NQ)                                   \
  do {                                                          \
    if (mode == 1) {                                            \
      if (P.mix.do_temp) LAUNCH_STREAM_MODE(NT, NQ, 1, true);   \
      else LAUNCH_STREAM_MODE(NT, NQ, 1, false);                \
    } else {                                                    \
      if (P.mix.do_temp) LAUNCH_STREAM_MODE(NT, NQ, 2, true);   \
      else LAUNCH_STREAM_MODE(NT, NQ, 2, false);                \
    }                                                           \
  } while (0)
#define LAUNCH_FAST(NT, NQ)                     \
  do {                                          \
    if (use_stream) LAUNCH_STREAM(NT, NQ);      \
    else if (mode == 1) LAUNCH_FAST_MODE(NT, NQ, 1); \
    else LAUNCH_FAST_MODE(NT, NQ, 2);           \
  } while (0)
  if (!(phases & 1) || (phases & 4)) {
  } else if (wide) {   // wider than the register-resident limit: multi-pass kernel
    row_stats_wide_kernel<DT><<<(unsigned)std::min<long long>(rows, 4LL * kNumSMs), kWideThreads, 0, stream>>>(P);
  } else if (mode) {
    if (nt == 256 && nq_inst == 2) LAUNCH_FAST(256, 2);
    else if (nt == 256 && nq_inst == 4) LAUNCH_FAST(256, 4);
    else if (nt == 256 && nq_inst == 8) LAUNCH_FAST(256, 8);
    else if (nt == 512 && nq_inst == 4) LAUNCH_FAST(512, 4);
    else if (nt == 512 && nq_inst == 8) LAUNCH_FAST(512, 8);
    else if (nt == 1024 && nq_inst == 2) LAUNCH_FAST(1024, 2);
    else if (nt == 1024 && nq_inst == 4) LAUNCH_FAST(1024, 4);
    else if (nt == 1024 && nq_inst == 8) LAUNCH_FAST(1024, 8);
    else {
      set_error("no fast row-statistics instantiation for %d threads x %d quads", nt, nq_inst);
      return LANTERN_E_UNSUPPORTED;
    }
  } else if (nt == 256) {
    if (nq_inst == 1) LAUNCH_STATS(256, 1);
    else if (nq_inst == 2) LAUNCH_STATS(256, 2);
    else if (nq_inst == 4) LAUNCH_STATS(256, 4);
    else LAUNCH_STATS(256, 8);
  } else if (nq_inst == 8) LAUNCH_STATS(512, 8);
  else if (nq_inst == 16) LAUNCH_STATS(512, 16);
  else {
    set_error("ncols=%d exceeds the register-resident row limit (%d)", c.ncols, 16 * 4 * 512);
    return LANTERN_E_UNSUPPORTED;
  }
#undef LAUNCH_FAST
#undef LAUNCH_STREAM
#undef LAUNCH_STREAM_MODE
#undef LAUNCH_FAST_MODE
#undef LAUNCH_STATS
  LANTERN_CUDA(cudaGetLastError());
  if ((phases & 1) && !(phases & 4) && P.do_topp) {   // nucleus cut on top of the row statistics (slow path, one CTA per row)
    row_topp_kernel<DT, VEC><<<(unsigned)rows, kToppThreads, 0, stream>>>(P);
    LANTERN_CUDA(cudaGetLastError());
  }
  if (!(phases & 2)) return LANTERN_OK;
  // lazy mode (phases bit 2): the walk computes the statistics of the rows it visits itself
  int lazy_ne = 0;
  if ((phases & 4) && VEC && !P.do_topp && c.ncols % (4 * kLazyThreads) == 0) {
    const int ne = c.ncols / kLazyThreads;
    if (c.ncols == 2048 || c.ncols == 4096 || c.ncols == 8192 || c.ncols == 16384) lazy_ne = ne;
  }
  if ((phases & 4) && !lazy_ne) {
    set_error("lazy statistics need a vector-aligned window of 2048/4096/8192/16384 columns and no top-p");
    return LANTERN_E_UNSUPPORTED;
  }
  const size_t smem = walk_smem_bytes(c, lazy_ne, P.p_spill != nullptr);
  if (smem + 8 * 1024 > 227 * 1024) {   // + the kernel's static arrays
    set_error("walk kernel needs %zu bytes of shared memory (> 227 KB): ncols too large", smem);
    return LANTERN_E_UNSUPPORTED;
  }
#define LAUNCH_WALK(V, NE)                                                                          \
  do {                                                                                              \
    auto kern = walk_kernel<DT, V, NE>;                                                             \
    if (smem > 36 * 1024)   /* the 48 KB default covers static + dynamic: the kernel has up to ~8 KB of static arrays */ \
      LANTERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<c.n_items, (NE) > 0 ? kLazyThreads : kWalkThreads, smem, stream>>>(P);                   \
  } while (0)
  switch (lazy_ne) {
    case 0: LAUNCH_WALK(VEC, 0); break;
    case 2048 / kLazyThreads: LAUNCH_WALK(true, 2048 / kLazyThreads); break;
    case 4096 / kLazyThreads: LAUNCH_WALK(true, 4096 / kLazyThreads); break;
    case 8192 / kLazyThreads: LAUNCH_WALK(true, 8192 / kLazyThreads); break;
    default: LAUNCH_WALK(true, 16384 / kLazyThreads); break;
  }
#undef LAUNCH_WALK
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}

}  // namespace lantern

using namespace lantern;

// the walk's probability vector spills to global memory when the shared-memory carve-up would not fit
static bool walk_spills(const lantern_accept_cfg& c) { return walk_smem_bytes(c, 0, false) + 8 * 1024 > 227 * 1024; }
static size_t stats_bytes(const lantern_accept_cfg& c) {
  return ((size_t)c.n_items * (size_t)c.n_rows * sizeof(RowStats) + 255) & ~size_t(255);
}

extern "C" size_t lantern_accept_workspace_bytes(const lantern_accept_cfg* cfg) {
  if (!cfg) return 0;
  size_t n = stats_bytes(*cfg);
  if (walk_spills(*cfg)) n += (size_t)cfg->n_items * (size_t)((cfg->ncols + 3) & ~3) * sizeof(float);
  return n;
}

// Acklam's rational approximation of the inverse normal CDF (|error| < 1.2e-9); only seeds a search bracket.
static double norm_ppf(double p) {
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  if (p <= 0.0) return -8.0;
  if (p >= 1.0) return 8.0;
  if (p < 0.02425) {
    const double q = sqrt(-2 * log(p));
    return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  if (p > 1 - 0.02425) return -norm_ppf(1 - p);
  const double q = p - 0.5, r = q * q;
  return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
         (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

static int validate(const lantern_accept_cfg& c, const lantern_accept_in& in, const lantern_accept_out& out) {
#define REQUIRE(cond, msg)                 \
  if (!(cond)) {                           \
    set_error("lantern_accept_fused: " msg); \
    return LANTERN_E_INVALID;              \
  }
  REQUIRE(c.n_items > 0 && c.n_rows > 0 && c.n_paths > 0 && c.depth > 0, "n_items/n_rows/n_paths/depth must be > 0");
  REQUIRE(c.vocab > 0 && c.ncols > 0 && c.col0 >= 0 && c.col0 + c.ncols <= c.vocab, "bad column window");
  REQUIRE(c.row_stride >= c.ncols && c.item_stride >= 0, "bad strides");
  REQUIRE(c.logits_dtype >= LANTERN_F32 && c.logits_dtype <= LANTERN_F16, "bad logits_dtype");
  REQUIRE(c.family >= LANTERN_FAMILY_VANILLA && c.family <= LANTERN_FAMILY_LUMINA, "bad family");
  REQUIRE(in.logits_cond && in.tree_tokens && in.retrieve, "logits_cond/tree_tokens/retrieve are required");
  REQUIRE(out.accept_length && out.best_candidate && out.token, "accept_length/best_candidate/token outputs are required");
  REQUIRE(c.temperature > 1e-5f, "temperature must be > 1e-5 (greedy decoding has no sampling walk)");
  REQUIRE(c.n_syntax >= 0 && c.n_syntax <= LANTERN_MAX_SYNTAX_TOKENS, "too many syntax tokens");
  if (c.lantern) {
    REQUIRE(in.nbr_table != nullptr, "lantern=1 needs nbr_table");
    REQUIRE(c.lantern_k >= 1 && c.lantern_k <= c.table_cols, "lantern_k must be in [1, table_cols]");
  }
  if (c.static_tree) {
    REQUIRE(in.node_q && in.draft_op && in.node_qrow && in.sib_off && in.sib_idx && in.sib_tokens,
            "static_tree=1 needs node_q/draft_op/node_qrow/sib_off/sib_idx/sib_tokens");
    REQUIRE(c.n_q_rows > 0, "static_tree=1 needs n_q_rows");
  }
  if (in.uniforms) REQUIRE(c.n_uniforms >= 1, "n_uniforms must be >= 1");
#undef REQUIRE
  return LANTERN_OK;
}

static void fill_params(AcceptParams& P, const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                        const lantern_accept_out* out, void* workspace_dev) {
  P.cfg = *cfg;
  P.in = *in;
  if (out) P.out = *out; else memset(&P.out, 0, sizeof(P.out));
  P.stats = static_cast<RowStats*>(workspace_dev);
  P.p_spill = (workspace_dev && walk_spills(*cfg))
                  ? reinterpret_cast<float*>(static_cast<unsigned char*>(workspace_dev) + stats_bytes(*cfg)) : nullptr;
  P.mix.cfg_scale = cfg->cfg_scale;
  P.mix.temperature = cfg->temperature;
  P.mix.has_uncond = in->logits_uncond != nullptr;
  P.mix.do_temp = cfg->temperature != 1.0f;
  P.do_topk = cfg->top_k > 0 && cfg->top_k < cfg->ncols;
  P.do_topp = (1e-8f <= cfg->top_p && cfg->top_p < 1.0f) ? 1 : 0;
  P.z_guess = P.do_topk ? (float)norm_ppf(1.0 - (double)cfg->top_k / (double)cfg->ncols) : 0.f;
  // Half-width of the tracked bracket: 2.9 standard deviations of the row-to-row difference of the sample quantile,
  // sqrt(2) * sqrt(p (1 - p) / n) / pdf(z) under the Gaussian prior (0.062 sd for top-k 2000 of 8192, 0.052 for 2000 of
  // 16384; measured optimum on B200 0.05-0.06).  Misses stay exact (tiers 2/3) and widen the bracket adaptively.
  P.win_sd = 0.08f;
  if (P.do_topk) {
    const double pk = (double)cfg->top_k / (double)cfg->ncols, z = P.z_guess;
    const double pdf = exp(-0.5 * z * z) / 2.5066282746310002;
    const double w = 2.9 * 1.4142135623730951 * sqrt(pk * (1.0 - pk) / (double)cfg->ncols) / (pdf > 1e-6 ? pdf : 1e-6);
    P.win_sd = (float)(w < 0.02 ? 0.02 : (w > 0.2 ? 0.2 : w));
  }
  if (const char* w = getenv("LANTERN_WIN_SD")) P.win_sd = (float)atof(w);   // tuning knob (any value keeps the select exact)
  // A CTA's first rows only have the Gaussian prior: twice the tracked width (the prior is exact for Gaussian rows, so
  // the noise of one sample quantile is all that has to fit; a miss costs one exact redo and doubles the width, up to
  // 0.25 sd).  Wider first brackets were measured slower: every parked element is work for one select warp.
  P.win_sd_first = fminf(0.25f, 2.0f * P.win_sd);
  if (const char* w = getenv("LANTERN_WIN_FIRST")) P.win_sd_first = (float)atof(w);
  P.inv_ncols = 1.0f / (float)cfg->ncols;
  P.tail_raw = cfg->family == LANTERN_FAMILY_VANILLA;
  P.lumina = cfg->family == LANTERN_FAMILY_LUMINA;
  P.static_zero_q = cfg->static_tree && cfg->family != LANTERN_FAMILY_LUMINA;
  const int eb = cfg->logits_dtype == LANTERN_F32 ? 4 : 2;
  const uintptr_t align = eb == 4 ? 16 : 8;
  auto aligned = [&](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % align) == 0; };
  P.vec_ok = (cfg->col0 % 4 == 0) && (cfg->ncols % 4 == 0) && (cfg->row_stride % 4 == 0) &&
             (cfg->item_stride % 4 == 0) && aligned(in->logits_cond) && aligned(in->logits_uncond);
}

static int dispatch_launch(const AcceptParams& P, cudaStream_t s, int phases) {
#define DISPATCH(DT) return P.vec_ok ? launch_all<DT, true>(P, s, phases) : launch_all<DT, false>(P, s, phases)
  switch (P.cfg.logits_dtype) {
    case LANTERN_F32: DISPATCH(LANTERN_F32);
    case LANTERN_BF16: DISPATCH(LANTERN_BF16);
    default: DISPATCH(LANTERN_F16);
  }
#undef DISPATCH
}

// Row statistics alone (CFG mix, warpers) for callers that consume RowStats themselves (draft_sample.cu).
int accept_row_stats_only(const lantern_accept_cfg* cfg, const lantern_accept_in* in, void* workspace_dev,
                          size_t workspace_bytes, void* stream, AcceptParams* params_out) {
  if (cfg->n_items <= 0 || cfg->n_rows <= 0 || cfg->vocab <= 0 || cfg->ncols <= 0 || cfg->col0 < 0 ||
      cfg->col0 + cfg->ncols > cfg->vocab || cfg->row_stride < cfg->ncols || !in->logits_cond ||
      cfg->logits_dtype < LANTERN_F32 || cfg->logits_dtype > LANTERN_F16 || !(cfg->temperature > 1e-5f)) {
    set_error("row statistics: bad shape / window / dtype / temperature");
    return LANTERN_E_INVALID;
  }
  if (!workspace_dev || workspace_bytes < lantern_accept_workspace_bytes(cfg)) {
    set_error("row statistics: workspace too small");
    return LANTERN_E_WORKSPACE;
  }
  fill_params(*params_out, cfg, in, nullptr, workspace_dev);
  return dispatch_launch(*params_out, static_cast<cudaStream_t>(stream), 1);
}

extern "C" int lantern_accept_fused(const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                    const lantern_accept_out* out, void* workspace_dev, size_t workspace_bytes,
                                    void* stream) {
  return lantern_accept_phases(cfg, in, out, workspace_dev, workspace_bytes, stream, 3);
}

extern "C" int lantern_accept_phases(const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                     const lantern_accept_out* out, void* workspace_dev, size_t workspace_bytes,
                                     void* stream, int phases) {
  if (!cfg || !in || !out) {
    set_error("lantern_accept_fused: null argument");
    return LANTERN_E_INVALID;
  }
  int rc = validate(*cfg, *in, *out);
  if (rc) return rc;
  if (!workspace_dev || workspace_bytes < lantern_accept_workspace_bytes(cfg)) {
    set_error("lantern_accept_fused: workspace too small (%zu < %zu)", workspace_bytes,
              lantern_accept_workspace_bytes(cfg));
    return LANTERN_E_WORKSPACE;
  }
  AcceptParams P;
  fill_params(P, cfg, in, out, workspace_dev);
  return dispatch_launch(P, static_cast<cudaStream_t>(stream), phases);
}

#ifdef LANTERN_WALK_TRACE
// profiling builds only: per-CTA phase marks of the last walk launch ([blocks][2 + 2 * kTraceMax] int64: count, draws, pairs)
extern "C" LANTERN_API int lantern_debug_walk_trace(long long* host, int max_blocks) {
  cudaDeviceSynchronize();
  const int n = max_blocks < kTraceBlocks ? max_blocks : kTraceBlocks;
  return (int)cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * (size_t)n * (2 * kTraceMax + 2));
}
#endif
