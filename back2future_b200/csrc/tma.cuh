// TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a, and the host-side tensor-map
// encoder.  libcuda is NOT linked: cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so that the library still loads on a box without a driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"

namespace b2f {

// ---- host: tensor map for a 4-D fp32 tensor (dims fastest first) -----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  // function-local static with an initialiser: C++11 guarantees one thread runs it and the others wait, so no
  // caller can observe "tried" without the pointer (the library is re-entrant, include/b2f.h)
  static const EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return nullptr;
  }();
  return fn;
}

// dims[0] is the contiguous dimension.  strides_elems[i] is the element stride of dims[i+1].
// Out-of-bounds elements of a box read as zero -- exactly the reference's "terms with an
// out-of-range source are dropped" (models/CostVolMulti.lua:77-88).
inline int make_tmap4(CUtensorMap* tm, const float* base, const uint64_t dims[4],
                      const uint64_t strides_elems[3], const uint32_t box[4], bool swizzle128 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(B2F_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_elems[0] * 4, strides_elems[1] * 4, strides_elems[2] * 4};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2F_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return B2F_OK;
}

// 2-D variant (packed convolution weights: rows of Cout floats, one row per (input channel, tap))
inline int make_tmap2(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t stride1_elems,
                      uint32_t b0, uint32_t b1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(B2F_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstr[1] = {stride1_elems * 4};
  cuuint32_t bx[2] = {b0, b1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2F_EINVAL, "cuTensorMapEncodeTiled (2-D) failed with CUresult %d", (int)r);
  return B2F_OK;
}

// 5-D variant (the gradOut slab: x, y, window row, window column, batch)
inline int make_tmap5(CUtensorMap* tm, const float* base, const uint64_t dims[5],
                      const uint64_t strides_elems[4], const uint32_t box[5], bool swizzle128 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(B2F_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5] = {dims[0], dims[1], dims[2], dims[3], dims[4]};
  cuuint64_t gstr[4] = {strides_elems[0] * 4, strides_elems[1] * 4, strides_elems[2] * 4, strides_elems[3] * 4};
  cuuint32_t bx[5] = {box[0], box[1], box[2], box[3], box[4]};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2F_EINVAL, "cuTensorMapEncodeTiled (5-D) failed with CUresult %d", (int)r);
  return B2F_OK;
}

// ---- device: mbarrier ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// Non-blocking probe (mbarrier.test_wait): lets a consumer ask for the NEXT stage while it still has
// arithmetic to issue, so the ~100-cycle barrier round trip overlaps the FFMAs.
__device__ __forceinline__ bool mbar_test(uint32_t bar_addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_addr(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                                 int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// ---- device: TMA tile store / reduce-add, shared -> global (bulk async group) ------------------
__device__ __forceinline__ void tma_store_4d_addr(uint32_t smem_src, const CUtensorMap* tm, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d_addr(uint32_t smem_src, const CUtensorMap* tm, int c0, int c1, int c2,
                                                  int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_5d_addr(uint32_t smem_src, const CUtensorMap* tm, int c0, int c1,
                                                       int c2, int c3, int c4) {
  asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d_addr(uint32_t smem_src, const CUtensorMap* tm, int c0, int c1,
                                                       int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk stores issued so far have finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// ---- packed fp32 pairs (Blackwell FFMA2: two independent fp32 FMAs per instruction, each rounded as fmaf) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// accumulating form: the accumulator is ONE read-write operand, so the PTX has a single virtual register per accumulator
// across a loop back-edge (with the value-returning form ptxas left 30 MOVs per two channels at the end of the
// cost-volume forward's channel loop: the phi copies of accumulators it had renamed, 19 % of the loop's issue slots)
__device__ __forceinline__ void fma2_acc(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// (hi half of a, lo half of b): the pair that starts one float later
__device__ __forceinline__ f32x2 straddle2(f32x2 a, f32x2 b) {
  f32x2 d;
  // volatile: assembled ONCE per use site (a plain asm is rematerialised by the compiler at every FFMA2)
  asm volatile("{\n.reg .b32 a0, a1, b0, b1;\nmov.b64 {a0, a1}, %1;\nmov.b64 {b0, b1}, %2;\nmov.b64 %0, {a1, b0};\n}"
      : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 16-byte shared-memory load as two packed pairs
__device__ __forceinline__ void lds128_pairs(uint32_t addr, f32x2& p0, f32x2& p1) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(p0), "=l"(p1) : "r"(addr));
}
__device__ __forceinline__ void tma_load_5d_addr(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                                 int c3, int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}
// 16-byte shared-memory load from a 32-bit shared address (no generic-pointer conversion in the loop)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ---- device: TMA tile load, 4-D, global -> shared, completes on an mbarrier ----------
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// ---- device: 4-byte cp.async with zero-fill, completion signalled on an mbarrier --------
// Used where a TMA box cannot be: TMA needs the innermost start coordinate 16-byte aligned
// (an unaligned one faults with "illegal instruction" on sm_100a).
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const float* gsrc, bool valid) {
  const int src_bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
// The executing thread's prior cp.async operations arrive on `bar` when they complete; the
// arrival was pre-counted in mbar_init (.noinc).
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

}  // namespace b2f
