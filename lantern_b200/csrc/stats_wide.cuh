// Row statistics for windows wider than the register-resident limit (more than 32768 live columns: plain EAGLE
// verification on a 65536-entry vocabulary, drafters/utils.py:333-410).  One CTA per row, the row is re-read from
// global memory (L2) for every pass: maximum, MSB-first radix select of the exact k-th largest value over ordered
// integer keys (four passes of 8 bits, 256-field shared-memory histogram), softmax sum over the kept columns.
// Parity path, not tuned: no BASELINE configuration has such rows (image-token windows are 8192 / 16384 wide).
#pragma once

#include "accept_types.cuh"

namespace lantern {

constexpr int kWideThreads = 1024;

template <int DT>
__global__ void __launch_bounds__(kWideThreads) row_stats_wide_kernel(const AcceptParams P) {
  __shared__ unsigned hist[258];
  __shared__ float fscr[33];
  __shared__ double dscr[33];
  const lantern_accept_cfg& cfg = P.cfg;
  const int tid = threadIdx.x;
  const long long n_rows_total = (long long)cfg.n_items * cfg.n_rows;
  const MixParams mix = P.mix;
  for (long long row = blockIdx.x; row < n_rows_total; row += gridDim.x) {
    const int b = (int)(row / cfg.n_rows), t = (int)(row % cfg.n_rows);
    RowStats st;
    st.thr = -INFINITY; st.mx = 0.f; st.sum = 1.f; st.vcut = -INFINITY; st.icut = -1;
    st.kind = P.in.row_kinds ? (int)P.in.row_kinds[row] : LANTERN_ROW_IMAGE;
    st.pad0 = st.pad1 = 0;
    if (st.kind != LANTERN_ROW_IMAGE) {
      if (tid == 0) P.stats[row] = st;
      continue;
    }
    const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)t * cfg.row_stride + cfg.col0;
    auto value = [&](int e) -> float {
      const float c = Elem<DT>::load1(P.in.logits_cond, base + e);
      const float u = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + e) : 0.f;
      return mix_temper(c, u, mix);
    };
    float m = -INFINITY;
    for (int e = tid; e < cfg.ncols; e += kWideThreads) m = fmaxf(m, value(e));
    m = block_reduce(m, OpMaxF(), -INFINITY, fscr);
    float thr = -INFINITY;
    if (P.do_topk) {
      uint32_t prefix = 0, mask = 0;
      int krem = cfg.top_k;
#pragma unroll 1
      for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kWideThreads) hist[i] = 0;
        __syncthreads();
        for (int e = tid; e < cfg.ncols; e += kWideThreads) {
          const uint32_t key = float_key(value(e));
          if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {   // descend from the top digit: the digit that holds the krem-th largest key
          unsigned run = 0;
          for (int dgt = 255; dgt >= 0; --dgt) {
            if (run < (unsigned)krem && run + hist[dgt] >= (unsigned)krem) { hist[256] = dgt; hist[257] = run; break; }
            run += hist[dgt];
          }
        }
        __syncthreads();
        prefix |= hist[256] << shift;
        mask |= 0xffu << shift;
        krem -= (int)hist[257];
        __syncthreads();
      }
      thr = key_float(prefix);
    }
    const ExpShift ex(m);
    float part = 0.f;
    for (int e = tid; e < cfg.ncols; e += kWideThreads) {
      const float s = value(e);
      part += (s >= thr) ? ex(s) : 0.f;
    }
    const double tot = block_reduce((double)part, OpSum(), 0.0, dscr);
    st.thr = thr; st.mx = m; st.sum = (float)tot;
    if (tid == 0) P.stats[row] = st;
    __syncthreads();
  }
}

}  // namespace lantern
