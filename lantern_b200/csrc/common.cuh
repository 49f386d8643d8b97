// Shared device/host helpers for the lantern_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lantern_b200.h"

namespace lantern {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define LANTERN_CUDA(expr)                                        \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::lantern::cuda_fail(_e, #expr); \
  } while (0)

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200

// ----------------------------------------------------------------------------------------------
// Streaming 4-element loads of a logits row (128-bit for fp32, 64-bit for bf16/fp16), L1 bypassed:
// each row is consumed once per kernel.
// ----------------------------------------------------------------------------------------------
template <int DT>
struct Elem;
template <>
struct Elem<LANTERN_F32> {
  static constexpr int kBytes = 4;
  static __device__ __forceinline__ void load4(const void* base, int64_t off, float (&o)[4]) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t off) {
    return __ldg(static_cast<const float*>(base) + off);
  }
};
template <>
struct Elem<LANTERN_BF16> {
  static constexpr int kBytes = 2;
  static __device__ __forceinline__ void load4(const void* base, int64_t off, float (&o)[4]) {
    const uint2* p = reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + off);
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    o[0] = __uint_as_float(v.x << 16);
    o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16);
    o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t off) {
    return __uint_as_float(static_cast<uint32_t>(__ldg(static_cast<const uint16_t*>(base) + off)) << 16);
  }
};
template <>
struct Elem<LANTERN_F16> {
  static constexpr int kBytes = 2;
  static __device__ __forceinline__ void load4(const void* base, int64_t off, float (&o)[4]) {
    const uint2* p = reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + off);
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t off) {
    __half h = __ushort_as_half(__ldg(static_cast<const uint16_t*>(base) + off));
    return __half2float(h);
  }
};

// 4-element reads of a row staged in shared memory (same element order as Elem<DT>::load4).
template <int DT>
__device__ __forceinline__ void lds4(const unsigned char* buf, int elem_off, float (&o)[4]) {
  if (DT == LANTERN_F32) {
    const float4 v = *reinterpret_cast<const float4*>(buf + (size_t)elem_off * 4);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(buf + (size_t)elem_off * 2);
    if (DT == LANTERN_BF16) {
      o[0] = __uint_as_float(v.x << 16);
      o[1] = __uint_as_float(v.x & 0xffff0000u);
      o[2] = __uint_as_float(v.y << 16);
      o[3] = __uint_as_float(v.y & 0xffff0000u);
    } else {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA 1-D bulk copy (cp.async.bulk) + mbarrier: the DMA engine streams the next logits row into shared
// memory while the SM works on the current one.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// CFG mix + temperature, one separately rounded fp32 op per reference ATen kernel
// (ea_model_llamagen.py:28: uncond + (cond - uncond) * scale; HF TemperatureLogitsWarper: scores / T).
// The intrinsics forbid FMA contraction so the value is bit-identical to the reference arithmetic.
struct MixParams {
  float cfg_scale;
  float temperature;
  int has_uncond;
  int do_temp;
};
__device__ __forceinline__ float mix_temper(float c, float u, const MixParams& m) {
  float v = c;
  if (m.has_uncond) v = __fadd_rn(u, __fmul_rn(__fsub_rn(c, u), m.cfg_scale));
  if (m.do_temp) v = __fdiv_rn(v, m.temperature);
  return v;
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2: one instruction, two lanes).  ptxas contracts a packed multiply feeding
// a packed add into one FFMA2 whatever the rounding modifiers say, which would change the reference's separately
// rounded CFG mix: mix_temper2 therefore keeps the multiply scalar.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t p, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two elements of the CFG mix + temperature, bit-identical to mix_temper on each: c - u as fma(u, -1, c) (one rounding
// of the exact difference), the scale as a scalar multiply, u + t as a packed add, the temperature as a scalar division
__device__ __forceinline__ uint64_t mix_temper2(float c0, float c1, float u0, float u1, const MixParams& m) {
  float v0 = c0, v1 = c1;
  if (m.has_uncond) {
    const uint64_t u = pack2(u0, u1);
    float t0, t1;
    unpack2(fma2(u, pack2(-1.0f, -1.0f), pack2(c0, c1)), t0, t1);
    unpack2(add2(u, pack2(__fmul_rn(t0, m.cfg_scale), __fmul_rn(t1, m.cfg_scale))), v0, v1);
  }
  if (m.do_temp) {
    v0 = __fdiv_rn(v0, m.temperature);
    v1 = __fdiv_rn(v1, m.temperature);
  }
  return pack2(v0, v1);
}

// exp(s - m) for the softmax: one FFMA + MUFU.EX2.  The same function is used wherever a probability is formed
// (row statistics, walk, tail), so numerator and denominator always agree; relative error <= ~3e-6 for
// s - m >= -60 (2 ulp of ex2.approx plus the rounding of the scaled argument), inside the 1e-5 budget.
struct ExpShift {
  float neg_m_log2e;
  __device__ __forceinline__ explicit ExpShift(float m) : neg_m_log2e(-m * 1.4426950408889634f) {}
  __device__ __forceinline__ float operator()(float s) const {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(s, 1.4426950408889634f, neg_m_log2e)));
    return r;
  }
};

// Monotone float -> uint32 key (larger float => larger key; -0 < +0 only as keys).
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

// ----------------------------------------------------------------------------------------------
// Warp / block reductions.  `scratch` must hold >= 33 elements of T and is reusable after return.
// ----------------------------------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T identity, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_reduce(v, op);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarp ? scratch[lane] : identity;
    w = warp_reduce(w, op);
    if (lane == 0) scratch[32] = w;
  }
  __syncthreads();
  return scratch[32];
}

// Thread groups the block-wide helpers below (and select.cuh) can be run by: the whole CTA, or the first N threads
// of it behind a named barrier (the streaming row-statistics kernel keeps one warp out of its main group).
struct BlockBar {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ int size() { return (int)blockDim.x; }
  static __device__ __forceinline__ int tid() { return (int)threadIdx.x; }
};
// threads OFF .. OFF + N - 1 of the CTA (a multiple of 32 each) behind named barrier ID
template <int ID, int N, int OFF = 0>
struct NamedBar {
  static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
  static __device__ __forceinline__ int size() { return N; }
  static __device__ __forceinline__ int tid() { return (int)threadIdx.x - OFF; }
};
template <class Bar, typename T, typename Op>
__device__ __forceinline__ T group_reduce(T v, Op op, T identity, T* scratch) {
  const int lane = Bar::tid() & 31, warp = Bar::tid() >> 5, nwarp = (Bar::size() + 31) >> 5;
  v = warp_reduce(v, op);
  Bar::sync();  // protect scratch from a previous use
  if (lane == 0) scratch[warp] = v;
  Bar::sync();
  if (warp == 0) {
    T w = lane < nwarp ? scratch[lane] : identity;
    w = warp_reduce(w, op);
    if (lane == 0) scratch[32] = w;
  }
  Bar::sync();
  return scratch[32];
}

struct OpSum {
  template <typename T>
  __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMaxF {
  __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
struct OpMinI {
  __device__ __forceinline__ int operator()(int a, int b) const { return a < b ? a : b; }
};
struct OpMaxI {
  __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// Block-wide scan of doubles: returns this thread's inclusive prefix, *total gets the block sum.
// `scratch` holds one double per warp (<= 32 warps).
template <class Bar = BlockBar>
__device__ __forceinline__ double block_scan_incl(double v, double* scratch, double* total) {
  const int lane = Bar::tid() & 31, warp = Bar::tid() >> 5, nwarp = (Bar::size() + 31) >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  Bar::sync();
  if (lane == 31) scratch[warp] = v;
  Bar::sync();
  // every warp scans the warp totals itself (no second barrier)
  double t = lane < nwarp ? scratch[lane] : 0.0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, t, o);
    if (lane >= o) t += n;
  }
  *total = __shfl_sync(0xffffffffu, t, nwarp - 1);
  const double pre = __shfl_sync(0xffffffffu, t, warp > 0 ? warp - 1 : 0);
  return (warp > 0 ? pre : 0.0) + v;
}

// ----------------------------------------------------------------------------------------------
// Philox-4x32-10 (Salmon et al. 2011).  Stream contract (shared with oracle/lantern_oracle.py
// philox_uniforms and lantern_philox_uniforms): draw d of item b at verify step s is word d%4 of
// philox(counter = (d/4, s_lo, b, s_hi), key = seed), mapped to [0,1) as (w >> 8) * 2^-24.
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = 0xD2511F53ull * c[0];
    uint64_t p1 = 0xCD9E8D57ull * c[2];
    uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = static_cast<uint32_t>(p1);
    uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = static_cast<uint32_t>(p0);
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__host__ __device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t step, uint32_t item,
                                                         uint32_t draw) {
  uint32_t c[4] = {draw >> 2, static_cast<uint32_t>(step), item, static_cast<uint32_t>(step >> 32)};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return static_cast<float>(c[draw & 3] >> 8) * (1.0f / 16777216.0f);
}

}  // namespace lantern
