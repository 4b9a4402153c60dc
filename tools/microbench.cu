// Micro-benchmarks that fix the numbers the kernel design and bench.py lean on (sm_100a):
//   fma     : fp32 FMA peak of the chip with FFMA (3-register form) and FFMA2 (packed pairs)     -> TFLOP/s
//   tmem    : lane/column mapping of tcgen05.ld.16x256b (checked against a pattern written with 32x32b stores)
//             and the TMEM read rate of 32x32b.x4 / 16x256b.x2 loads                               -> B/clk/SM
//   gather  : rows of `row_bytes` gathered from a 32 K-row table into shared memory, per SM, with
//             (a) one cp.async.bulk per row, (b) 16-byte cp.async per lane                         -> GB/s
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/microbench tools/microbench.cu
// Run  :  tools/bin/microbench            (prints one JSON object per test)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

static float time_ms(void (*launch)(void*), void* arg, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch(arg);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch(arg);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

// ------------------------------------------------------------------------------------------------ fma
constexpr int kFmaIters = 4096;
template <int MODE>  // 0: FFMA, 1: FFMA2, 2: FFMA2 with one shared (broadcast) multiplicand
__global__ void __launch_bounds__(512) fma_kernel(float* out, float a, float b) {
  if (MODE == 0) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    float m = a + threadIdx.x * 1e-9f, c = b;
#pragma unroll 1
    for (int it = 0; it < kFmaIters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], m, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 m = make_float2(a + threadIdx.x * 1e-9f, MODE == 2 ? a + threadIdx.x * 1e-9f : a * 1.0001f);
    float2 c = make_float2(b, b * 0.5f);
#pragma unroll 1
    for (int it = 0; it < kFmaIters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], m, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}
struct FmaArg { float* out; int mode; int grid; };
static void fma_launch(void* p) {
  FmaArg* a = (FmaArg*)p;
  if (a->mode == 0) fma_kernel<0><<<a->grid, 512>>>(a->out, 1.0000001f, 1e-7f);
  else if (a->mode == 1) fma_kernel<1><<<a->grid, 512>>>(a->out, 1.0000001f, 1e-7f);
  else fma_kernel<2><<<a->grid, 512>>>(a->out, 1.0000001f, 1e-7f);
}

// ------------------------------------------------------------------------------------------------ tmem
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) tmem_map_kernel(uint32_t* out) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base;
  // pattern: TMEM[lane L][column c] = L * 1000 + c   (32x32b: thread t of warp w owns lane 32 w + t)
  const uint32_t my = base + ((uint32_t)(32 * warp) << 16);
  for (int c = 0; c < 64; ++c) {
    const uint32_t v = (uint32_t)((32 * warp + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(my + (uint32_t)c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read back with 16x256b.x1 at lane bases 32 w and 32 w + 16, columns 8 .. 15
  for (int half = 0; half < 2; ++half) {
    uint32_t r0, r1, r2, r3;
    const uint32_t a = base + ((uint32_t)(32 * warp + 16 * half) << 16) + 8u;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(a)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)::"memory");
    uint32_t* o = out + ((warp * 2 + half) * 32 + lane) * 4;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
  }
  // and 16x256b.x2 (16 columns): 8 registers
  {
    uint32_t r[8];
    const uint32_t a = base + ((uint32_t)(32 * warp) << 16) + 16u;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t* o = out + 1024 + (warp * 32 + lane) * 8;
    for (int i = 0; i < 8; ++i) o[i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
}

constexpr int kTmemIters = 2048;
template <int MODE>  // 0: 32x32b.x4 (512 B / instr), 1: 16x256b.x2 (512 B / instr), 2: 32x32b.x16 (2 KB / instr)
__global__ void __launch_bounds__(512, 1) tmem_rate_kernel(uint32_t* out) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
#pragma unroll 1
  for (int it = 0; it < kTmemIters; ++it) {
    const uint32_t col = (uint32_t)((it * 16 + warp * 4) & 255);
    if (MODE == 0) {
      uint32_t r0, r1, r2, r3, q0, q1, q2, q3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + col) : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(base + col + 128u) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(q0), "+r"(q1), "+r"(q2), "+r"(q3)::"memory");
      acc += r0 ^ r1 ^ r2 ^ r3 ^ q0 ^ q1 ^ q2 ^ q3;
    } else if (MODE == 1) {
      uint32_t r[8], q[8];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(base + col) : "memory");
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                   : "r"(base + col + 128u) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7])::"memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += r[i] ^ q[i];
    } else {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(base + col) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += r[i];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512u) : "memory");
}
struct TmemArg { uint32_t* out; int mode; };
static void tmem_launch(void* p) {
  TmemArg* a = (TmemArg*)p;
  if (a->mode == 0) tmem_rate_kernel<0><<<148, 512>>>(a->out);
  else if (a->mode == 1) tmem_rate_kernel<1><<<148, 512>>>(a->out);
  else tmem_rate_kernel<2><<<148, 512>>>(a->out);
}

// ------------------------------------------------------------------------------------------------ gather
constexpr int kGatherRows = 64;       // rows per stage
constexpr int kGatherStages = 2;
constexpr int kGatherChunks = 256;    // chunks per CTA
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
// MODE 0: cp.async.bulk per row (lane = row), MODE 1: cp.async 16 B per lane (warp = row)
template <int MODE>
__global__ void __launch_bounds__(128, 1) gather_kernel(const float* __restrict__ table, const int* __restrict__ idx,
                                                       int row_floats, int nwarps, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kGatherStages];
  float* buf = reinterpret_cast<float*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)row_floats * 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGatherStages; ++s) mbar_init(&full[s], MODE == 0 ? 1 : nwarps * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  const int* my_idx = idx + (size_t)blockIdx.x * kGatherChunks * kGatherRows;
  for (int kk = -1; kk < kGatherChunks; ++kk) {
    // issue chunk kk + 1, then wait for chunk kk: the copy of the next chunk overlaps the wait
    const int k = kk + 1;
    const int s = k % kGatherStages;
    float* dst = buf + (size_t)s * kGatherRows * row_floats;
    if (warp < nwarps && k < kGatherChunks) {
      if (MODE == 0) {
        if (warp == 0 && lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])),
                       "r"(row_bytes * kGatherRows) : "memory");
        __syncwarp();
        for (int r = warp * 32 + lane; r < kGatherRows; r += nwarps * 32) {
          const int src = my_idx[k * kGatherRows + r];
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(dst + (size_t)r * row_floats)),
                       "l"(table + (size_t)src * row_floats), "r"(row_bytes), "r"(smem_u32(&full[s]))
                       : "memory");
        }
      } else {
        const int pieces = row_floats / 4;
        const int my_src0 = my_idx[k * kGatherRows + lane], my_src1 = my_idx[k * kGatherRows + 32 + lane];
        for (int r = warp; r < kGatherRows; r += nwarps) {
          const int src = __shfl_sync(0xffffffffu, r < 32 ? my_src0 : my_src1, r & 31);
          const float* g = table + (size_t)src * row_floats;
          for (int p = lane; p < pieces; p += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + (size_t)r * row_floats + p * 4)),
                         "l"(g + p * 4) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      }
    }
    if (kk >= 0) {
      mbar_wait(&full[kk % kGatherStages], (kk / kGatherStages) & 1);
      acc += buf[(size_t)(kk % kGatherStages) * kGatherRows * row_floats + (threadIdx.x * 13) % (kGatherRows * row_floats)];
    }
    __syncthreads();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// MODE 2: TMA tile::gather4 (4 rows per instruction, lane = group of 4 rows); box = {row_floats, 1}
__global__ void __launch_bounds__(128, 1) gather4_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx,
                                                        int row_floats, int nwarps, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kGatherStages];
  float* buf = reinterpret_cast<float*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)row_floats * 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGatherStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  const int* my_idx = idx + (size_t)blockIdx.x * kGatherChunks * kGatherRows;
  for (int kk = -1; kk < kGatherChunks; ++kk) {
    // issue chunk kk + 1, then wait for chunk kk: the copy of the next chunk overlaps the wait
    const int k = kk + 1;
    const int s = k % kGatherStages;
    float* dst = buf + (size_t)s * kGatherRows * row_floats;
    if (warp < nwarps && k < kGatherChunks) {
      if (warp == 0 && lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])),
                     "r"(row_bytes * kGatherRows) : "memory");
      __syncwarp();
      for (int g = warp * 32 + lane; g < kGatherRows / 4; g += nwarps * 32) {
        const int4 r4 = *reinterpret_cast<const int4*>(my_idx + k * kGatherRows + 4 * g);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 "
            "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst + (size_t)4 * g * row_floats)),
            "l"(&tmap), "r"(0), "r"(r4.x), "r"(r4.y), "r"(r4.z), "r"(r4.w), "r"(smem_u32(&full[s]))
            : "memory");
      }
    }
    if (kk >= 0) {
      mbar_wait(&full[kk % kGatherStages], (kk / kGatherStages) & 1);
      acc += buf[(size_t)(kk % kGatherStages) * kGatherRows * row_floats + (threadIdx.x * 13) % (kGatherRows * row_floats)];
    }
    __syncthreads();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// correctness of gather4: table[r][c] = r * 1000 + c, gather rows {5, 3, 100, 7} -> smem -> out
__global__ void gather4_check_kernel(const __grid_constant__ CUtensorMap tmap, int row_floats, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  float* buf = reinterpret_cast<float*>(smem);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)row_floats * 16u) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(buf)), "l"(&tmap), "r"(0), "r"(5), "r"(3), "r"(100), "r"(7),
        "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 4 * row_floats; i += blockDim.x) out[i] = buf[i];
}
struct Gather4Arg { CUtensorMap tmap; const int* idx; int row_floats; int nwarps; float* out; };
static void gather4_launch(void* p) {
  Gather4Arg* a = (Gather4Arg*)p;
  const size_t smem = (size_t)kGatherStages * kGatherRows * a->row_floats * 4;
  gather4_kernel<<<148, 128, smem>>>(a->tmap, a->idx, a->row_floats, a->nwarps, a->out);
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_row_tmap(CUtensorMap* m, const float* table, int rows, int row_floats, int box_floats) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)row_floats, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_floats * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_floats, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)table, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("{\"error\": \"cuTensorMapEncodeTiled -> %d\"}\n", (int)r);
  return r == CUDA_SUCCESS;
}
struct GatherArg { const float* table; const int* idx; int row_floats; int nwarps; float* out; int mode; };
static void gather_launch(void* p) {
  GatherArg* a = (GatherArg*)p;
  const size_t smem = (size_t)kGatherStages * kGatherRows * a->row_floats * 4;
  if (a->mode == 0) gather_kernel<0><<<148, 128, smem>>>(a->table, a->idx, a->row_floats, a->nwarps, a->out);
  else gather_kernel<1><<<148, 128, smem>>>(a->table, a->idx, a->row_floats, a->nwarps, a->out);
}

int main(int argc, char** argv) {
  const char* only = argc > 1 ? argv[1] : "";
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", prop.name, prop.multiProcessorCount, clock_khz / 1000);
  float* out;
  CK(cudaMalloc(&out, 64 << 20));
  if (!*only || !strcmp(only, "fma")) {
    const char* names[3] = {"ffma", "ffma2", "ffma2_shared_operand"};
    for (int mode = 0; mode < 3; ++mode) {
      FmaArg a{out, mode, prop.multiProcessorCount * 4};
      const float ms = time_ms(fma_launch, &a, 20);
      const double flops = (double)a.grid * 512 * kFmaIters * 16 * (mode == 0 ? 2.0 : 4.0);
      printf("{\"test\": \"fma_peak\", \"instr\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f}\n", names[mode], ms,
             flops / ms * 1e-9);
    }
  }
  if (!*only || !strcmp(only, "tmem")) {
    uint32_t* d = reinterpret_cast<uint32_t*>(out);
    CK(cudaMemset(d, 0xff, 4096 * 4));
    tmem_map_kernel<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> h(2048);
    CK(cudaMemcpy(h.data(), d, 2048 * 4, cudaMemcpyDeviceToHost));
    // expectation: thread t, lane base B, column base C: r0,r1 = (B + t/4, C + 2 (t%4) + {0,1}), r2,r3 = (B + t/4 + 8, same)
    int bad = 0;
    for (int w = 0; w < 4; ++w)
      for (int half = 0; half < 2; ++half)
        for (int t = 0; t < 32; ++t) {
          const uint32_t* r = &h[((w * 2 + half) * 32 + t) * 4];
          const int B = 32 * w + 16 * half, C = 8;
          const uint32_t e0 = (B + t / 4) * 1000 + C + 2 * (t % 4), e2 = (B + t / 4 + 8) * 1000 + C + 2 * (t % 4);
          if (r[0] != e0 || r[1] != e0 + 1 || r[2] != e2 || r[3] != e2 + 1) ++bad;
        }
    int bad2 = 0;
    for (int w = 0; w < 4; ++w)
      for (int t = 0; t < 32; ++t) {
        const uint32_t* r = &h[1024 + (w * 32 + t) * 8];
        const int B = 32 * w, C = 16;
        for (int rep = 0; rep < 2; ++rep) {
          const uint32_t e0 = (B + t / 4) * 1000 + C + 8 * rep + 2 * (t % 4), e2 = (B + t / 4 + 8) * 1000 + C + 8 * rep + 2 * (t % 4);
          if (r[4 * rep] != e0 || r[4 * rep + 1] != e0 + 1 || r[4 * rep + 2] != e2 || r[4 * rep + 3] != e2 + 1) ++bad2;
        }
      }
    printf("{\"test\": \"tmem_16x256b_map\", \"x1_mismatches\": %d, \"x2_mismatches\": %d, \"t5_x1\": [%u, %u, %u, %u], "
           "\"t5_x2\": [%u, %u, %u, %u, %u, %u, %u, %u]}\n",
           bad, bad2, h[5 * 4], h[5 * 4 + 1], h[5 * 4 + 2], h[5 * 4 + 3], h[1024 + 40], h[1024 + 41], h[1024 + 42],
           h[1024 + 43], h[1024 + 44], h[1024 + 45], h[1024 + 46], h[1024 + 47]);
    const char* names[3] = {"32x32b.x4", "16x256b.x2", "32x32b.x16"};
    for (int mode = 0; mode < 3; ++mode) {
      TmemArg a{d, mode};
      const float ms = time_ms(tmem_launch, &a, 10);
      const double bytes_per_sm = 16.0 * kTmemIters * (mode == 2 ? 2048.0 : 1024.0);
      const double clk = ms * 1e-3 * clock_khz * 1e3;
      printf("{\"test\": \"tmem_read_rate\", \"shape\": \"%s\", \"ms\": %.4f, \"bytes_per_clk_per_sm\": %.1f}\n", names[mode],
             ms, bytes_per_sm / clk);
    }
  }
  if (!*only || !strcmp(only, "gather")) {
    const int rows = 32768;
    for (int row_floats : {16, 100, 200, 248}) {
      float* table;
      int* idx;
      CK(cudaMalloc(&table, (size_t)rows * row_floats * 4));
      CK(cudaMemset(table, 0, (size_t)rows * row_floats * 4));
      const size_t nidx = (size_t)148 * kGatherChunks * kGatherRows;
      std::vector<int> hidx(nidx);
      uint64_t s = 12345;
      for (size_t i = 0; i < nidx; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        // senders are near the receiver in the bench graph: a window of 64 rows around a moving centre
        hidx[i] = (int)(((i / 28) + (s >> 33) % 64) % rows);
      }
      CK(cudaMalloc(&idx, nidx * 4));
      CK(cudaMemcpy(idx, hidx.data(), nidx * 4, cudaMemcpyHostToDevice));
      const size_t smem = (size_t)kGatherStages * kGatherRows * row_floats * 4;
      CK(cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int mode = 0; mode < 2; ++mode)
        for (int nw : {1, 2, 4}) {
          GatherArg a{table, idx, row_floats, nw, out, mode};
          const float ms = time_ms(gather_launch, &a, 5);
          const double bytes = (double)nidx * row_floats * 4;
          const double clk_per_row = ms * 1e-3 * clock_khz * 1e3 / (kGatherChunks * kGatherRows);
          printf("{\"test\": \"gather\", \"how\": \"%s\", \"row_bytes\": %d, \"warps\": %d, \"ms\": %.4f, \"gbps\": %.0f, "
                 "\"clk_per_row_per_sm\": %.1f}\n",
                 mode == 0 ? "cp.async.bulk/row" : "cp.async 16B/lane", row_floats * 4, nw, ms, bytes / ms * 1e-6, clk_per_row);
        }
      {
        // gather4: box padded to a multiple of 8 floats so that 4 rows are a multiple of 128 bytes
        const int box = (row_floats + 7) / 8 * 8;
        Gather4Arg a;
        if (make_row_tmap(&a.tmap, table, rows, row_floats, box)) {
          std::vector<float> ht((size_t)rows * row_floats);
          for (int r = 0; r < rows; ++r) for (int c = 0; c < row_floats; ++c) ht[(size_t)r * row_floats + c] = (float)(r * 1000 + c);
          CK(cudaMemcpy(table, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice));
          CK(cudaFuncSetAttribute(gather4_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * box * 4));
          gather4_check_kernel<<<1, 128, 4 * box * 4>>>(a.tmap, box, out);
          CK(cudaDeviceSynchronize());
          std::vector<float> ho(4 * box);
          CK(cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost));
          const int want_rows[4] = {5, 3, 100, 7};
          int bad = 0;
          for (int j = 0; j < 4; ++j) for (int c = 0; c < row_floats; ++c) if (ho[j * box + c] != (float)(want_rows[j] * 1000 + c)) ++bad;
          printf("{\"test\": \"gather4_check\", \"row_floats\": %d, \"box\": %d, \"mismatches\": %d, \"pad0\": %.0f}\n", row_floats, box, bad,
                 box > row_floats ? ho[row_floats] : -1.f);
          a.idx = idx; a.row_floats = box; a.out = out;
          const size_t smem4 = (size_t)kGatherStages * kGatherRows * box * 4;
          CK(cudaFuncSetAttribute(gather4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
          for (int nw : {1, 2}) {
            a.nwarps = nw;
            const float ms = time_ms(gather4_launch, &a, 5);
            const double bytes = (double)nidx * row_floats * 4;
            const double clk_per_row = ms * 1e-3 * clock_khz * 1e3 / (kGatherChunks * kGatherRows);
            printf("{\"test\": \"gather\", \"how\": \"tma gather4\", \"row_bytes\": %d, \"warps\": %d, \"ms\": %.4f, \"gbps\": %.0f, "
                   "\"clk_per_row_per_sm\": %.1f}\n", row_floats * 4, nw, ms, bytes / ms * 1e-6, clk_per_row);
          }
        }
      }
      CK(cudaFree(table));
      CK(cudaFree(idx));
    }
  }
  return 0;
}
