// Fused convolution forward, Blackwell-native version (fp32 storage, l <= 2 tensor-product types):
//
//   last radial-MLP layer  w[e, c] = sum_k h[e,k] W[k,c]   ->  tcgen05.mma (bf16 x3 split, fp32 accumulate)
//   accumulators                                           ->  TMEM  (row = weight column c, column = edge e)
//   uvu Clebsch-Gordan contraction + per-receiver sum      ->  FP32 FMA pipes, operands in shared memory,
//                                                              per-node sums in registers (CSR order, no atomics)
//
// Same contract as conv_fwd.cuh (reference src/matten/nn/utils.py:260-263 + src/matten/nn/conv.py:113-120):
// the per-edge tensor-product weights [E, weight_numel] and the messages [E, D_mid] never leave the SM.
//
// Two kernels:
//  (1) edge_prepare_kernel: one thread per receiver-sorted edge evaluates the small hidden layers of the radial MLP
//      (e.g. 8 -> 32 -> 32) and stores h as three bf16 planes (hi/mid/lo) in the K-major core-matrix layout the
//      MMA wants, plus the edge's spherical-harmonic row padded to 16 bytes.  192 + 48 bytes per edge.
//  (2) conv_fwd_tc_kernel: 1 CTA per SM, persistent over a contiguous node range (balanced by edge count),
//      warp specialised:
//        warp 0      builds node-aligned chunks (every node's edges padded to a multiple of 4 columns, <= 64
//                    columns) and issues cp.async.bulk copies: gathered sender rows x[src], sh rows, h planes.
//        warp 1      one thread issues the tcgen05.mma's of the chunk: D[t] (128 x 64, TMEM) = A[t] (128 rows of
//                    W^T, K = 32) x B^T (64 edges), 6 significant products of the 3 x 3 bf16 split (~2^-24).
//        warps 4..31 consumers.  Warp w reads TMEM lanes 32*(w%4).. .  A group of 32 TMEM lanes holds 32 weight
//                    columns of ONE (l1,l2,l3) type (or several small types packed; those run lane-phased).
//                    Work units (sub-item, node) are handed out dynamically.  Per edge a lane gets w from TMEM
//                    (tcgen05.ld 32x32b.x4, software pipelined), x / sh from shared memory and runs the generated
//                    CG contraction; pad columns carry zero weights, so the edge loop has no range checks.
//      Double buffered TMEM accumulators and x / sh staging; mbarriers: full[b] (bulk-copy bytes + tcgen05.commit),
//      empty[b] (28 consumer warps), bready / bfree (h planes landed / consumed by the MMAs).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "generated/cg_gen.cuh"

namespace mt {

// columns (padded edges) per chunk == MMA N: 2 stages x MT tiles x NE columns must fit the 512 TMEM columns, so
// small plans take wider chunks (more nodes per chunk: more units to balance, fewer producer round trips)
__host__ __device__ constexpr int tc_chunk_cols(int MT) { return MT <= 1 ? 256 : (MT == 2 ? 128 : 64); }
constexpr int kTcK = 32;              // padded size of the last hidden layer == MMA K total
constexpr int kTcProducerWarps = 4;   // warp 0: chunks + plane / sh copies, warp 1: MMA issue, warps 2-3: x-row copies
// consumer warps per chunk width (a multiple of 4: warp w reads TMEM quarter w % 4).  Fewer than the 28 that fill a
// 1024-thread CTA: with 16 (4 per quarter / scheduler) the kernel gets 92 registers per thread and each scheduler's
// instruction stream stays resident -- measured on the bench layers (ms per call, layers 1..3), 28 warps: 1.19 / 1.40 /
// 1.47, 24: 1.03 / 1.22 / 1.24, 20: 1.16 / 1.33 / 1.31, 16: 1.09 / 1.11 / 1.18, 12: 1.25 / 1.30 / 1.36.
__host__ __device__ constexpr int tc_consumer_warps(int NE) { return NE == 128 ? 24 : 16; }
__host__ __device__ constexpr int tc_threads(int NE) { return 32 * (kTcProducerWarps + tc_consumer_warps(NE)); }
constexpr int kTcMaxTiles = 4;        // M tiles of 128 rows -> <= 512 weight-column rows
constexpr int kTcMaxSub = 64;         // sub-items per plan
constexpr int kTcMaxNodes = 16;       // receiver nodes per chunk
constexpr int kTcD = 5;               // 2*lmax+1 of the types this kernel handles (l <= 2)

// debugging aid: per-warp progress codes of CTA `dbg_block` into a host-visible buffer (MT_CONV_TC_DEBUG)
// phase timing of CTA 0 (MT_CONV_TC_DEBUG = device pointer to >= 16384 int64): slot layout in tools/tc_debug.py
// (compiled in only with -DMT_TC_TIMING: the extra code in the consumer loop costs instruction-cache hits)
#ifdef MT_TC_TIMING
#define MT_TACC(var) do { if (p.dbg && blockIdx.x == 0) { const long long _n = clock64(); (var) += _n - _tl; _tl = _n; } } while (0)
#define MT_TIMING_ONLY(...) __VA_ARGS__
#define MT_DBG(code) do { if (p.dbg && (threadIdx.x & 31) == 0) ((volatile long long*)p.dbg)[256 + blockIdx.x * 32 + (threadIdx.x >> 5)] = (long long)(code); } while (0)
#define MT_DBG_BLOCK(code) do { if (p.dbg && threadIdx.x == 0) ((volatile long long*)p.dbg)[blockIdx.x] = (long long)(code); } while (0)
#else
#define MT_TACC(var) do { } while (0)
#define MT_TIMING_ONLY(...)
#define MT_DBG(code) do { } while (0)
#define MT_DBG_BLOCK(code) do { } while (0)
#endif

struct ConvTcParams {
  int x_dim, y_dim, out_dim;
  int num_tiles;             // MT
  int num_sub;               // sub-items
  const int32_t* row_wcol;   // [MT*128] weight column of every A row (-1: zero row)
  const int32_t* sub_hdr;    // [num_sub][8] {type, cpw, lane0 (first TMEM lane within the quarter), tile, quarter, D3,0,0}
  const int32_t* sub_slot;   // [num_sub][32][4] per lane {xoff, yoff, ooff, valid}
  const int32_t* q_list;     // [4][kTcMaxSub] sub-item ids per quarter (heavy first)
  int q_count[4];
  int nl;
  int sizes[MT_MAX_MLP_LAYERS + 1];
  int act;
  float act_cst;
  const float* w[MT_MAX_MLP_LAYERS];
  const float* x;
  const float* sh;
  const float* emb;
  const int32_t* rowptr;
  const int32_t* perm;
  const int32_t* src;
  float avg;
  const float* num_neigh;
  float* out;
  int64_t N, E;
  int y_pad;                   // y_dim rounded up to a multiple of 4 floats
  __nv_bfloat16* hplanes;      // workspace [3][4][E][8] bf16: plane, k-group, edge, k%8
  float* ysorted;              // workspace [E][y_pad]
  long long* dbg;              // optional clock64 stamps of CTA 0 (MT_CONV_TC_DEBUG), else nullptr
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug (or a faulted bulk copy) must end in a trap with a message, never in a hang.
__device__ __noinline__ void mbar_timeout(int id, uint32_t parity) {
  printf("[matten_b200] mbarrier wait timed out: barrier %d parity %u block %d warp %d\n", id, parity, (int)blockIdx.x,
         (int)(threadIdx.x >> 5));
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int id = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  long long t0 = 0;
  uint32_t spins = 0;
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
    if (!done && (++spins & 63u) == 0) {  // ~2 s at 2 GHz: far beyond any legitimate wait in this kernel
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) mbar_timeout(id, parity);
    }
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.ld is asynchronous: the destination registers are valid only after tcgen05.wait::ld.  The wait takes
// the registers as read-write operands so the compiler cannot move a use above it.  (Scalars, not arrays:
// arrays passed through asm operands end up in local memory.)
#define MT_TMEM_LD_X4(taddr, r0, r1, r2, r3)                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"     \
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)                            \
               : "r"(taddr)                                                        \
               : "memory")
#define MT_TMEM_LD_WAIT(r0, r1, r2, r3) \
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)::"memory")

// bulk asynchronous copy global -> shared (the non-tensor TMA path); completes bytes on an mbarrier.
// size and both addresses are multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// K-major, no-swizzle ("interleave") shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
//   element (row r, k) of a [rows x 16] bf16 K-step lives at
//   start + (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2   bytes
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t start_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((start_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// cute::UMMA::InstrDescriptor for kind::f16: BF16 x BF16 -> F32, both K-major
__device__ __forceinline__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_bf16x3(float v, __nv_bfloat16& hi, __nv_bfloat16& mi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(hi);
  mi = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mi);
  lo = __float2bfloat16_rn(r2);
}

struct __align__(16) TcMeta {
  int nnodes;  // -1: terminate
  int ncols;   // used (padded) columns; 0: only nodes without edges
  int pad0, pad1;
  int node_id[kTcMaxNodes];
  float den[kTcMaxNodes];            // sqrt(avg_num_neighbors) or sqrt(num_neigh[node])
  short cb[kTcMaxNodes];             // first column of the node (multiple of 4)
  short ngrp[kTcMaxNodes];           // padded edge count / 4
  unsigned char first[kTcMaxNodes];  // 1: this chunk holds the node's first edges (plain store), 0: accumulate
  int rlo[kTcMaxNodes];              // first receiver-sorted edge of the node's piece in this chunk
  int deg[kTcMaxNodes];              // its (unpadded) edge count
};

// packed per-lane slot of a sub-item (8 bytes)
struct __align__(8) TcSlot {
  unsigned short xoff;
  unsigned char yoff;
  unsigned char valid;
  int ooff;
};

// shared-memory carve-up (bytes)
struct TcSmemLayout {
  size_t a_off, a_plane;  // 3 planes of [MT*128 x 32] bf16
  size_t b_off, b_plane;  // 3 planes of [64 x 32] bf16
  size_t x_off, x_buf;    // 2 buffers [64][x_dim] fp32
  size_t y_off, y_buf;    // 2 buffers [64][y_pad] fp32
  size_t meta_off;        // 2 TcMeta
  size_t slot_off;        // [num_sub][32] TcSlot
  size_t hdr_off;         // [num_sub] packed {type, cpw, lane0, tile, D3}
  size_t qlist_off;       // [4][kTcMaxSub] uint8
  size_t total;
};
__host__ __device__ inline size_t tc_align16(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline TcSmemLayout tc_smem_layout(int MT, int x_dim, int y_pad, int num_sub, int NE) {
  TcSmemLayout L;
  size_t o = 0;
  L.a_off = o;
  L.a_plane = (size_t)MT * 128 * kTcK * 2;
  o += 3 * L.a_plane;
  L.b_off = o;
  L.b_plane = (size_t)NE * kTcK * 2;
  o += 3 * L.b_plane;
  L.x_off = o;
  L.x_buf = (size_t)NE * x_dim * 4;
  o += 2 * L.x_buf;
  L.y_off = o;
  L.y_buf = (size_t)NE * y_pad * 4;
  o += 2 * L.y_buf;
  L.meta_off = o;
  o += 2 * tc_align16(sizeof(TcMeta));
  L.slot_off = o;
  o += (size_t)num_sub * 32 * sizeof(TcSlot);
  L.hdr_off = o;
  o += tc_align16((size_t)num_sub * 4);
  L.qlist_off = o;
  o += 4 * kTcMaxSub;
  L.total = o;
  return L;
}

// =====================================================================================================
// (1) per-edge preparation: hidden layers of the radial MLP -> bf16 planes, sh row -> 16-byte padded row
// =====================================================================================================
constexpr int kPrepThreads = 128;

__global__ void __launch_bounds__(kPrepThreads) edge_prepare_kernel(const ConvTcParams p) {
  // weights of the hidden layers (pre-scaled by 1/sqrt(fan_in)) and one private activation row per thread
  __shared__ __align__(16) float sW[2][kTcK * kTcK];
  __shared__ float sH[kPrepThreads][kTcK + 1];
  const int nh = p.nl - 1;
  for (int li = 0; li < nh; ++li) {
    const int fi = p.sizes[li], fo = p.sizes[li + 1];
    const float s = rsqrtf((float)fi);
    for (int t = threadIdx.x; t < kTcK * kTcK; t += kPrepThreads) {
      const int k = t >> 5, j = t & 31;
      sW[li][t] = (k < fi && j < fo) ? p.w[li][(size_t)k * fo + j] * s : 0.f;
    }
  }
  __syncthreads();
  const int in0 = p.sizes[0];
  float* hrow = sH[threadIdx.x];
  for (int64_t e = blockIdx.x * (int64_t)kPrepThreads + threadIdx.x; e < p.E; e += (int64_t)gridDim.x * kPrepThreads) {
    const int64_t orig = p.perm[e];
    // sh row, padded
    {
      // loads first, stores after: a store waiting for its load would block the next load behind it (in-order issue)
      const float* __restrict__ yr = p.sh + orig * p.y_dim;
      float* __restrict__ yo = p.ysorted + e * p.y_pad;
      for (int j0 = 0; j0 < p.y_pad; j0 += 12) {
        float yv[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) yv[j] = (j0 + j < p.y_dim) ? yr[j0 + j] : 0.f;
#pragma unroll
        for (int j = 0; j < 12; j += 4)
          if (j0 + j < p.y_pad) *reinterpret_cast<float4*>(yo + j0 + j) = make_float4(yv[j], yv[j + 1], yv[j + 2], yv[j + 3]);
      }
    }
    float h[kTcK];
#pragma unroll
    for (int j = 0; j < kTcK; ++j) h[j] = 0.f;
    {
      const float* er = p.emb + orig * in0;
#pragma unroll
      for (int j = 0; j < kTcK; ++j)
        if (j < in0) h[j] = er[j];
    }
    for (int li = 0; li < nh; ++li) {
      const int fi = p.sizes[li], fo = p.sizes[li + 1];
#pragma unroll
      for (int j = 0; j < kTcK; ++j) hrow[j] = h[j];
      float2 a[kTcK / 2];  // output pairs: one FFMA2 (broadcast h[k]) per pair
#pragma unroll
      for (int j = 0; j < kTcK / 2; ++j) a[j] = make_float2(0.f, 0.f);
      const float* W = sW[li];
#pragma unroll 4
      for (int k = 0; k < fi; ++k) {
        const float hk = hrow[k];
        const float4* wr = reinterpret_cast<const float4*>(W + k * kTcK);
#pragma unroll
        for (int j4 = 0; j4 < kTcK / 4; ++j4) {
          const float4 w4 = wr[j4];
          fma_pair(hk, w4.x, w4.y, a[2 * j4]);
          fma_pair(hk, w4.z, w4.w, a[2 * j4 + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < kTcK; ++j) {
        const float aj = (j & 1) ? a[j >> 1].y : a[j >> 1].x;
        h[j] = (j < fo) ? apply_act<float>(p.act, aj) * p.act_cst : 0.f;
      }
    }
    // planes [3][4][E][8]
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      __align__(16) __nv_bfloat16 hi[8], mi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) split_bf16x3(h[g * 8 + i], hi[i], mi[i], lo[i]);
      uint4* base = reinterpret_cast<uint4*>(p.hplanes);
      base[(size_t)(0 * 4 + g) * p.E + e] = *reinterpret_cast<const uint4*>(hi);
      base[(size_t)(1 * 4 + g) * p.E + e] = *reinterpret_cast<const uint4*>(mi);
      base[(size_t)(2 * 4 + g) * p.E + e] = *reinterpret_cast<const uint4*>(lo);
    }
  }
}

// =====================================================================================================
// (2) fused kernel
// =====================================================================================================

// One (sub-item, node) unit for one (l1,l2,l3) type: ngrp groups of 4 columns starting at column cb.
// Compact on purpose (2 inlined contractions for lane==row items, 1 for packed items): the hot code of 28 warps
// on different types has to stay in the instruction cache (fully unrolled per-type loops measured a 74 % I-cache
// hit rate with the GPC instruction cache at 86 % of its request peak).
template <int L1, int L2, int L3>
__device__ __forceinline__ void tc_unit(uint32_t taddr, const float* __restrict__ xp, int xstride,
                                        const float* __restrict__ yp, int ystride, int ngrp, int cpw, int lane,
                                        int src_lane, float* __restrict__ acc) {
  constexpr int D1 = 2 * L1 + 1, D2 = 2 * L2 + 1;
  uint32_t c0, c1, c2, c3, n0 = 0, n1 = 0, n2 = 0, n3 = 0;
  MT_TMEM_LD_X4(taddr, c0, c1, c2, c3);
  MT_TMEM_LD_WAIT(c0, c1, c2, c3);
  if (cpw == 32) {
    // lane == TMEM row: every lane walks all columns, two edges per iteration packed into FFMA2 / FMUL2 (edge a in
    // the low half, edge b in the high half of every register pair); the two partial sums are added at the end
    constexpr int D3 = 2 * L3 + 1;
    f2 acc2[D3];
#pragma unroll
    for (int m = 0; m < D3; ++m) acc2[m] = f2(0.f, 0.f);
    auto two_edges = [&](const f2 w2) {
      f2 x2[D1], y2[D2];
#pragma unroll
      for (int m = 0; m < D1; ++m) x2[m] = f2(xp[m], xp[xstride + m]);
#pragma unroll
      for (int m = 0; m < D2; ++m) y2[m] = f2(yp[m], yp[ystride + m]);
      CG<L1, L2, L3>::template fwd<f2>(x2, y2, w2, acc2);
      xp += 2 * xstride;
      yp += 2 * ystride;
    };
#pragma unroll 1
    for (int g = 0; g < ngrp; ++g) {
      const bool more = g + 1 < ngrp;
      if (more) MT_TMEM_LD_X4(taddr + (uint32_t)(4 * g + 4), n0, n1, n2, n3);  // prefetch the next 4 columns
      // (rolled on purpose: straight-line code for the 4 columns of a group measured 4 % slower -- instruction fetch)
#pragma unroll 1
      for (int h = 0; h < 2; ++h) two_edges(f2(__uint_as_float(h ? c2 : c0), __uint_as_float(h ? c3 : c1)));
      if (more) {
        MT_TMEM_LD_WAIT(n0, n1, n2, n3);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      }
    }
#pragma unroll
    for (int m = 0; m < D3; ++m) acc[m] += acc2[m].v.x + acc2[m].v.y;
  } else {
    // packed small types: lane = (column j, phase ph); phase ph takes the columns == ph (mod nphase) of every
    // group and fetches its weight from the lane that owns the column's TMEM row
    const int nphase = 32 / cpw, phase = lane / cpw;
    xp += phase * xstride;
    yp += phase * ystride;
#pragma unroll 1
    for (int g = 0; g < ngrp; ++g) {
      const bool more = g + 1 < ngrp;
      if (more) MT_TMEM_LD_X4(taddr + (uint32_t)(4 * g + 4), n0, n1, n2, n3);
      const float t0 = __shfl_sync(0xffffffffu, __uint_as_float(c0), src_lane);
      const float t1 = __shfl_sync(0xffffffffu, __uint_as_float(c1), src_lane);
      const float t2 = __shfl_sync(0xffffffffu, __uint_as_float(c2), src_lane);
      const float t3 = __shfl_sync(0xffffffffu, __uint_as_float(c3), src_lane);
      // warp-uniform trip count (r), lane-dependent column j = r + phase selected without branches: the warp
      // must be converged when it reaches the next .sync.aligned TMEM instruction
#pragma unroll 1
      for (int r = 0; r < 4; r += nphase) {
        const int j = r + phase;
        const float w = (j == 0) ? t0 : (j == 1) ? t1 : (j == 2) ? t2 : t3;
        const float* xr = xp + r * xstride;
        const float* yr = yp + r * ystride;
        float xa[D1], ya[D2];
#pragma unroll
        for (int m = 0; m < D1; ++m) xa[m] = xr[m];
#pragma unroll
        for (int m = 0; m < D2; ++m) ya[m] = yr[m];
        CG<L1, L2, L3>::template fwd<float>(xa, ya, w, acc);
      }
      xp += 4 * xstride;
      yp += 4 * ystride;
      __syncwarp();
      if (more) {
        MT_TMEM_LD_WAIT(n0, n1, n2, n3);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      }
    }
  }
}

template <int NE>
__global__ void __launch_bounds__(tc_threads(NE), 1) conv_fwd_tc_kernel(const ConvTcParams p) {
  constexpr int kTcConsumerWarps = tc_consumer_warps(NE), kTcThreads = tc_threads(NE);
  extern __shared__ __align__(128) unsigned char smem[];  // no swizzle: descriptors need 16 B alignment
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_go[2], bar_bready, bar_bfree;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_cnt[2][4];
  __shared__ int s_mma_b;  // buffer of the chunk handed to the MMA warp (-1: terminate)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int MT = p.num_tiles;
  const TcSmemLayout L = tc_smem_layout(MT, p.x_dim, p.y_pad, p.num_sub, NE);
  unsigned char* sA = smem + L.a_off;
  unsigned char* sB = smem + L.b_off;
  TcMeta* meta0 = reinterpret_cast<TcMeta*>(smem + L.meta_off);
  TcMeta* meta1 = reinterpret_cast<TcMeta*>(smem + L.meta_off + tc_align16(sizeof(TcMeta)));
  TcSlot* sSlot = reinterpret_cast<TcSlot*>(smem + L.slot_off);
  uint32_t* sHdr = reinterpret_cast<uint32_t*>(smem + L.hdr_off);
  unsigned char* sQ = smem + L.qlist_off;
  const uint32_t tmem_cols = (2 * MT * NE <= 32) ? 32 : (2 * MT * NE <= 64) ? 64 : (2 * MT * NE <= 128) ? 128
                             : (2 * MT * NE <= 256) ? 256 : 512;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    mbar_init(&bar_full[0], 2);
    mbar_init(&bar_full[1], 2);
    mbar_init(&bar_empty[0], kTcConsumerWarps);
    mbar_init(&bar_empty[1], kTcConsumerWarps);
    mbar_init(&bar_go[0], 1);
    mbar_init(&bar_go[1], 1);
    mbar_init(&bar_bready, 1);
    mbar_init(&bar_bfree, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem_base, tmem_cols);
  // plan tables -> shared memory (with a ~225 KB carve-out the L1 is a few KB: every global load is an L2 trip)
  for (int t = tid; t < p.num_sub * 32; t += kTcThreads) {
    const int4 v = reinterpret_cast<const int4*>(p.sub_slot)[t];
    TcSlot sl;
    sl.xoff = (unsigned short)v.x;
    sl.yoff = (unsigned char)v.y;
    sl.valid = (unsigned char)v.w;
    sl.ooff = v.z;
    sSlot[t] = sl;
  }
  for (int t = tid; t < p.num_sub; t += kTcThreads) {
    const int* h = p.sub_hdr + t * 8;
    sHdr[t] = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)(h[3] & 15) << 24) |
              ((uint32_t)(h[5] & 15) << 28);  // type | cpw | lane0 | tile | D3
  }
  for (int t = tid; t < 4 * kTcMaxSub; t += kTcThreads) sQ[t] = (unsigned char)p.q_list[t];
  // staging buffers start zeroed: pad columns are read (with zero weights) and must hold finite values
  {
    uint4* z = reinterpret_cast<uint4*>(smem + L.b_off);
    const int n16 = (int)((L.meta_off - L.b_off) >> 4);
    for (int t = tid; t < n16; t += kTcThreads) z[t] = make_uint4(0u, 0u, 0u, 0u);
  }
  // A planes: rows of W_last^T (pre-scaled by 1/sqrt(H)) split into bf16 hi/mid/lo, canonical K-major layout
  {
    const int H = p.sizes[p.nl - 1], Wn = p.sizes[p.nl];
    const float* __restrict__ Wl = p.w[p.nl - 1];
    const float s = rsqrtf((float)H);
    const int rows = MT * 128;
    for (int t = tid; t < rows * 4; t += kTcThreads) {
      const int R = t >> 2, g = t & 3;
      const int wc = p.row_wcol[R];
      __align__(16) __nv_bfloat16 hi[8], mi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = g * 8 + i;
        const float v = (wc >= 0 && k < H) ? Wl[(size_t)k * Wn + wc] * s : 0.f;
        split_bf16x3(v, hi[i], mi[i], lo[i]);
      }
      const size_t off = (size_t)g * rows * 16 + (size_t)R * 16;
      *reinterpret_cast<uint4*>(sA + off) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(sA + L.a_plane + off) = *reinterpret_cast<const uint4*>(mi);
      *reinterpret_cast<uint4*>(sA + 2 * L.a_plane + off) = *reinterpret_cast<const uint4*>(lo);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ================================================================ chunk builder + bulk copies
    // node range of this CTA: boundaries at equal shares of the edge list
    int64_t n_cur, n_end;
    {
      auto bound = [&](int64_t i) -> int64_t {  // first node whose rowptr >= i*E/grid
        if (i <= 0) return 0;
        if (i >= (int64_t)gridDim.x) return p.N;
        const int64_t target = (p.E * i) / gridDim.x;
        int64_t lo = 0, hi = p.N;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (p.rowptr[mid] < target) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      if (p.E == 0) {
        n_cur = (p.N * blockIdx.x) / gridDim.x;
        n_end = (p.N * (blockIdx.x + 1)) / gridDim.x;
      } else {
        n_cur = bound(blockIdx.x);
        n_end = bound(blockIdx.x + 1);
      }
    }
    int e_carry = -1;  // >= 0: the next chunk continues node n_cur at this global edge
    int b_uses = 0;    // chunks that loaded the h planes so far
    MT_TIMING_ONLY(long long _tl = clock64(), t_we = 0, t_meta = 0, t_wb = 0, t_pad = 0, t_issue = 0;)
    long long n_chunks = 0;
    const uint32_t xrow_bytes = (uint32_t)p.x_dim * 4, yrow_bytes = (uint32_t)p.y_pad * 4;
    for (int k = 0;; ++k) {
      const int b = k & 1;
      TcMeta& M = b ? *meta1 : *meta0;
      MT_DBG(1000 + k * 10 + 0);
      if (n_cur >= n_end) {
        if (lane == 0) mbar_wait(&bar_empty[b], ((k >> 1) & 1) ^ 1, 10 + b);  // consumers released buffer b
        __syncwarp();
        if (lane == 0) {
          M.nnodes = -1;
          M.ncols = 0;
          mbar_arrive(&bar_go[b]);  // the x-row warps see the terminator
          if (b_uses > 0) mbar_wait(&bar_bfree, (b_uses - 1) & 1, 20);  // the MMA warp is done with its last chunk
          s_mma_b = -1;
          mbar_arrive(&bar_bready);  // terminator for the MMA warp
          mbar_arrive(&bar_full[b]);
          mbar_arrive(&bar_full[b]);
        }
        break;
      }
      // ---- candidates: lane l looks at node n_cur + l (one coalesced rowptr read)
      const int64_t cand = n_cur + lane;
      const bool in_range = lane < kTcMaxNodes && cand < n_end;
      int r_lo = in_range ? p.rowptr[cand] : 0;
      const int r_hi = in_range ? p.rowptr[cand + 1] : 0;
      if (lane == 0 && e_carry >= 0) r_lo = e_carry;
      int deg = r_hi - r_lo;
      bool split = false;
      if (lane == 0 && deg > NE) {  // node larger than a chunk: take 64 edges now, the rest next time
        deg = NE;
        split = true;
      }
      const int degp = (deg + 3) & ~3;
      int cum = in_range ? degp : 0x10000;  // inclusive prefix of padded degrees
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) cum += t;
      }
      const unsigned fm = __ballot_sync(0xffffffffu, in_range && cum <= NE);
      int m = (fm == 0xffffffffu) ? 32 : (__ffs(~fm) - 1);  // leading run of nodes that fit
      const bool split0 = __shfl_sync(0xffffffffu, (int)split, 0) != 0;
      if (split0) m = 1;
      const bool mine = lane < m;
      const int cb = cum - degp;
      const float den = mine ? (p.num_neigh ? sqrtf(p.num_neigh[cand]) : sqrtf(p.avg)) : 0.f;
      const int ncols = __shfl_sync(0xffffffffu, cum, m - 1);
      const int my_edges = mine ? deg : 0;
      int tot_edges = my_edges;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot_edges += __shfl_xor_sync(0xffffffffu, tot_edges, o);
      // unpadded and consecutive in the sorted edge list (always, for whole nodes): the chunk's columns are ONE
      // contiguous edge range -> one copy per plane segment instead of one per node
      const bool contiguous = (ncols == tot_edges);
      const int r_first = __shfl_sync(0xffffffffu, r_lo, 0);
      MT_TACC(t_meta);
      // ---- everything above only read global memory; now wait until the consumers released buffer b
      if (lane == 0) mbar_wait(&bar_empty[b], ((k >> 1) & 1) ^ 1, 10 + b);
      MT_DBG(1000 + k * 10 + 1);
      __syncwarp();
      MT_TACC(t_we);
      if (mine) {
        M.node_id[lane] = (int)cand;
        M.cb[lane] = (short)cb;
        M.ngrp[lane] = (short)(degp >> 2);
        M.first[lane] = (lane == 0 && e_carry >= 0) ? 0 : 1;
        M.den[lane] = den;
        M.rlo[lane] = r_lo;
        M.deg[lane] = deg;
      }
      const bool continues = (e_carry >= 0);
      if (lane == 0) {
        M.nnodes = m;
        M.ncols = ncols;
        // a chunk that continues a node split over chunks accumulates into its output row: keep the pieces
        // ordered (deterministic sum) by letting the previous chunk drain first
        if (continues && k > 0) mbar_wait(&bar_empty[b ^ 1], ((k - 1) >> 1) & 1, 12);
      }
      __syncwarp();
      float* xs = reinterpret_cast<float*>(smem + L.x_off + (size_t)b * L.x_buf);
      float* ys = reinterpret_cast<float*>(smem + L.y_off + (size_t)b * L.y_buf);
      if (ncols == 0) {
        if (lane == 0) {
          s_cnt[b][0] = 0; s_cnt[b][1] = 0; s_cnt[b][2] = 0; s_cnt[b][3] = 0;
          mbar_arrive(&bar_go[b]);
          mbar_arrive(&bar_full[b]);  // nothing to copy, no MMA: both arrivals from here
          mbar_arrive(&bar_full[b]);
        }
      } else {
        if (lane == 0) {
          s_cnt[b][0] = 0; s_cnt[b][1] = 0; s_cnt[b][2] = 0; s_cnt[b][3] = 0;
          if (b_uses > 0) mbar_wait(&bar_bfree, (b_uses - 1) & 1, 21);  // previous MMAs have consumed the h planes
        }
        __syncwarp();
        MT_TACC(t_wb);
        MT_DBG(1000 + k * 10 + 2);
        // pad rows of the h planes must be zero (zero weights for pad columns)
        for (int j = 0; j < (contiguous ? 0 : m); ++j) {
          const int dj = __shfl_sync(0xffffffffu, deg, j), cj = __shfl_sync(0xffffffffu, cb, j);
          const int npad = ((dj + 3) & ~3) - dj;
          for (int t = lane; t < npad * 12; t += 32) {
            const int r = t / 12, pg = t - r * 12;
            *reinterpret_cast<uint4*>(sB + (size_t)(pg >> 2) * L.b_plane + (size_t)(pg & 3) * NE * 16 +
                                      (size_t)(cj + dj + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          s_mma_b = b;
          mbar_arrive_expect_tx(&bar_full[b], (uint32_t)tot_edges * (xrow_bytes + yrow_bytes));
          mbar_arrive_expect_tx(&bar_bready, (uint32_t)tot_edges * 16u * 12u);
          mbar_arrive(&bar_go[b]);  // warps 2-3 issue the gathered x rows (one bulk copy per edge) from here
        }
        __syncwarp();
        MT_TACC(t_pad);
        MT_DBG(1000 + k * 10 + 3);
        // bulk copies: 12 plane segments + the sh rows per node (contiguous: sorted order) -- or per chunk when the
        // chunk's columns are one contiguous edge range; the x rows (one per edge) are issued by warps 2-3
        if (contiguous) {
          if (lane < 12) {
            const int pl = lane >> 2, g = lane & 3;
            bulk_g2s(sB + (size_t)pl * L.b_plane + (size_t)g * NE * 16,
                     reinterpret_cast<const unsigned char*>(p.hplanes) + ((size_t)(pl * 4 + g) * p.E + r_first) * 16,
                     (uint32_t)tot_edges * 16u, &bar_bready);
          } else if (lane == 12) {
            bulk_g2s(ys, p.ysorted + (size_t)r_first * p.y_pad, (uint32_t)tot_edges * yrow_bytes, &bar_full[b]);
          }
        }
        for (int j = 0; j < (contiguous ? 0 : m); ++j) {
          const int dj = __shfl_sync(0xffffffffu, deg, j), cj = __shfl_sync(0xffffffffu, cb, j);
          const int rj = __shfl_sync(0xffffffffu, r_lo, j);
          if (dj == 0) continue;
          if (lane < 12) {
            const int pl = lane >> 2, g = lane & 3;
            bulk_g2s(sB + (size_t)pl * L.b_plane + (size_t)g * NE * 16 + (size_t)cj * 16,
                     reinterpret_cast<const unsigned char*>(p.hplanes) + ((size_t)(pl * 4 + g) * p.E + rj) * 16,
                     (uint32_t)dj * 16u, &bar_bready);
          } else if (lane == 12) {
            bulk_g2s(ys + (size_t)cj * p.y_pad, p.ysorted + (size_t)rj * p.y_pad, (uint32_t)dj * yrow_bytes,
                     &bar_full[b]);
          }
        }
        ++b_uses;
        ++n_chunks;
        MT_TACC(t_issue);
        MT_DBG(1000 + k * 10 + 4);
      }
      // advance
      if (split0) {
        e_carry = __shfl_sync(0xffffffffu, r_lo, 0) + NE;
        if (e_carry >= __shfl_sync(0xffffffffu, r_hi, 0)) {  // exactly consumed
          e_carry = -1;
          n_cur += 1;
        }
      } else {
        e_carry = -1;
        n_cur += m;
      }
    }
    MT_TIMING_ONLY(if (p.dbg && blockIdx.x == 0 && lane == 0) {
      p.dbg[8192 + 0] = t_we; p.dbg[8192 + 1] = t_meta; p.dbg[8192 + 2] = t_wb; p.dbg[8192 + 3] = t_pad;
      p.dbg[8192 + 4] = t_issue; p.dbg[8192 + 6] = n_chunks;
    })
    (void)n_chunks;
  } else if (warp == 2 || warp == 3) {
    // ================================================================ gathered x rows: one bulk copy per edge
    // (a UBLKCP is issued lane by lane, ~100 cycles each: two warps on two schedulers halve the issue time)
    const int part = warp - 2;
    const uint32_t xrow_bytes = (uint32_t)p.x_dim * 4;
    MT_TIMING_ONLY(long long _tl = clock64(), t_wg = 0, t_cp = 0;)
    for (int k = 0;; ++k) {
      const int b = k & 1;
      const TcMeta& M = b ? *meta1 : *meta0;
      if (lane == 0) mbar_wait(&bar_go[b], (k >> 1) & 1, 50 + b);
      __syncwarp();
      MT_TACC(t_wg);
      const int nn = M.nnodes;
      if (nn < 0) break;
      if (M.ncols > 0) {
        float* xs = reinterpret_cast<float*>(smem + L.x_off + (size_t)b * L.x_buf);
        for (int j = 0; j < nn; ++j) {
          const int dj = M.deg[j], cj = M.cb[j], rj = M.rlo[j];
          for (int e = 2 * lane + part; e < dj; e += 64)
            bulk_g2s(xs + (size_t)(cj + e) * p.x_dim, p.x + (size_t)p.src[rj + e] * p.x_dim, xrow_bytes, &bar_full[b]);
        }
      }
      MT_TACC(t_cp);
    }
    MT_TIMING_ONLY(if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[8192 + 10 + 2 * part] = t_wg; p.dbg[8192 + 11 + 2 * part] = t_cp; })
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, NE);
    const uint32_t a_lbo = (uint32_t)MT * 128 * 16, b_lbo = (uint32_t)NE * 16;
    const uint64_t a_desc0 = make_kmajor_desc(smem_u32(sA), a_lbo, 128);
    const uint64_t b_desc0 = make_kmajor_desc(smem_u32(sB), b_lbo, 128);
    MT_TIMING_ONLY(long long _tl = clock64(), t_wr = 0, t_mma = 0;)
    for (int k = 0;; ++k) {  // k counts the chunks that carry edges (and the terminator)
      MT_DBG(2000 + k * 10 + 0);
      // lane 0 alone reads the hand-over word: by the time the other lanes get here the producer may already
      // have posted the next chunk (or the terminator) -- a per-lane read would split the warp
      int b = 0;
      if (lane == 0) {
        mbar_wait(&bar_bready, k & 1, 30);
        b = s_mma_b;
      }
      b = __shfl_sync(0xffffffffu, b, 0);
      MT_DBG(2000 + k * 10 + 1);
      MT_TACC(t_wr);
      if (b < 0) break;
      if (lane == 0) {
        tc_fence_after();
        // significant products of (hi+mid+lo) x (hi+mid+lo), small ones first
        const int pa[6] = {0, 2, 1, 0, 1, 0};
        const int pb[6] = {2, 0, 1, 1, 0, 0};
        for (int t = 0; t < MT; ++t) {
          const uint32_t d = tmem_base + (uint32_t)((b * MT + t) * NE);
          uint32_t accum = 0;
#pragma unroll
          for (int q = 0; q < 6; ++q) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              // descriptors differ only in the start-address field (units of 16 bytes, no carry: < 2^14)
              const uint32_t a_off = (uint32_t)(pa[q] * L.a_plane) + (uint32_t)(2 * ks) * a_lbo + (uint32_t)t * 128 * 16;
              const uint32_t b_off = (uint32_t)(pb[q] * L.b_plane) + (uint32_t)(2 * ks) * b_lbo;
              umma_bf16(d, a_desc0 + (uint64_t)(a_off >> 4), b_desc0 + (uint64_t)(b_off >> 4), idesc, accum);
              accum = 1;
            }
          }
        }
        umma_commit(&bar_full[b]);
        umma_commit(&bar_bfree);
      }
      __syncwarp();
      MT_TACC(t_mma);
      MT_DBG(2000 + k * 10 + 2);
    }
    MT_TIMING_ONLY(if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[8192 + 8] = t_wr; p.dbg[8192 + 9] = t_mma; })
  } else if (warp >= kTcProducerWarps) {
    // ================================================================ consumers
    const int q = warp & 3;
    const int nsubq = p.q_count[q];
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    MT_TIMING_ONLY(long long _tl = clock64(), t_wf = 0, t_work = 0;)
    for (int k = 0;; ++k) {
      const int b = k & 1;
      MT_DBG(3000 + k * 100 + 0);
      mbar_wait(&bar_full[b], (k >> 1) & 1, 40 + b);
      MT_TACC(t_wf);
      MT_DBG(3000 + k * 100 + 1);
      tc_fence_after();
      const TcMeta& M = b ? *meta1 : *meta0;
      const int nn = M.nnodes;
      if (nn < 0) break;
      const float* xs = reinterpret_cast<const float*>(smem + L.x_off + (size_t)b * L.x_buf);
      const float* ys = reinterpret_cast<const float*>(smem + L.y_off + (size_t)b * L.y_buf);
      const int num_units = nsubq * nn;
      while (true) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(&s_cnt[b][q], 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= num_units) break;
        const int si = unit / nn, nj = unit - si * nn;
        const int sub = sQ[q * kTcMaxSub + si];
        MT_DBG(3000 + k * 100 + 10 + unit);
        const uint32_t hd = sHdr[sub];
        const int type = hd & 0xff, cpw = (hd >> 8) & 0xff, lane0 = (hd >> 16) & 0xff, tile = (hd >> 24) & 15;
        const int d3 = hd >> 28;
        const TcSlot slot = sSlot[sub * 32 + lane];
        const int cb = M.cb[nj], ngrp = M.ngrp[nj];
        MT_TIMING_ONLY(const long long u0 = (p.dbg && blockIdx.x == 0) ? clock64() : 0;)
        float acc[kTcD];
#pragma unroll
        for (int m = 0; m < kTcD; ++m) acc[m] = 0.f;
        __syncwarp();
        if (ngrp > 0) {
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)((b * MT + tile) * NE + cb);
          const float* xp = xs + (size_t)cb * p.x_dim + slot.xoff;
          const float* yp = ys + (size_t)cb * p.y_pad + slot.yoff;
          const int src_lane = (lane0 + (lane & (cpw - 1))) & 31;
          switch (type) {
#define MT_TC_CASE(ID, A, B, C)                                                           \
  case ID:                                                                                \
    tc_unit<A, B, C>(taddr, xp, p.x_dim, yp, p.y_pad, ngrp, cpw, lane, src_lane, acc);    \
    break;
            MT_FOR_EACH_CG_TYPE_L2(MT_TC_CASE)
#undef MT_TC_CASE
            default: break;
          }
        }
        for (int off = cpw; off < 32; off <<= 1) {
#pragma unroll
          for (int m = 0; m < kTcD; ++m) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], off);
        }
        MT_TIMING_ONLY(if (p.dbg && blockIdx.x == 0 && lane == 0) {
          atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg) + 8192 + 128 + sub, (unsigned long long)(clock64() - u0));
          atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg) + 8192 + 256 + sub, 1ull);
        })
        if (slot.valid && lane < cpw) {
          float* o = p.out + (size_t)M.node_id[nj] * p.out_dim + slot.ooff;
          const float den = M.den[nj];
          // `first` is warp-uniform; keep the accumulate case apart so the common case issues no loads
          if (M.first[nj]) {
#pragma unroll
            for (int m = 0; m < kTcD; ++m)
              if (m < d3) o[m] = acc[m] / den;
          } else {
#pragma unroll
            for (int m = 0; m < kTcD; ++m)
              if (m < d3) o[m] += acc[m] / den;
          }
        }
      }
      MT_DBG(3000 + k * 100 + 90);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[b]);
      MT_TACC(t_work);
    }
    MT_TIMING_ONLY(if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[8192 + 16 + 2 * warp] = t_wf; p.dbg[8192 + 17 + 2 * warp] = t_work; })
  }
  // ---------------------------------------------------------------- teardown
  MT_DBG(9000);
  tc_fence_before();
  __syncthreads();
  MT_DBG_BLOCK(1);
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  MT_DBG_BLOCK(2);
}

}  // namespace mt
