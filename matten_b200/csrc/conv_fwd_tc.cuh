// Fused convolution forward, Blackwell-native version (fp32 storage):
//
//   radial-MLP last layer  w[e, c] = sum_k h[e,k] W[k,c]   ->  tcgen05.mma (bf16 x3 split, fp32 accum)
//   accumulators                                           ->  TMEM  (row = weight column c, column = edge e)
//   uvu Clebsch-Gordan contraction + per-receiver sum      ->  FP32 FMA pipes, operands in shared memory,
//                                                              per-node sums in registers (CSR order, no atomics)
//
// Same contract as conv_fwd.cuh (reference src/matten/nn/utils.py:260-263 + src/matten/nn/conv.py:113-120);
// per-edge weights and messages never leave the SM.
//
// CTA = 1 per SM, persistent over a contiguous range of receiver nodes (balanced by edge count).
//   warps 0..3   producers: build node-aligned chunks of <= 64 edges, cp.async-gather x[src] / sh / emb rows,
//                evaluate the hidden MLP layers, split h into three bf16 planes (hi/mid/lo) in the canonical
//                K-major core-matrix layout, and (one thread) issue the tcgen05.mma's:
//                  D[t] (128 x 64, TMEM) = A[t] (128 rows of W^T, K=32) * B^T (64 edges, K=32)
//                with the 6 significant products of the 3x3 bf16 split (error ~2^-24, i.e. fp32 grade).
//   warps 4..19  consumers: warp w reads TMEM lanes 32*(w%4).. (its "quarter").  A row group of 32 TMEM lanes
//                holds 32 weight columns of ONE (l1,l2,l3) type (or several small types packed, processed
//                as lane-phased sub-items).  Work units (sub-item, node) are handed out dynamically per
//                quarter.  Per edge a lane gets w from TMEM (tcgen05.ld 32x32b.x4), x / sh from shared memory
//                and runs the generated CG contraction.
//   Double buffering: TMEM accumulators and the x/sh staging buffers, so gather + MLP + MMA of chunk k+1
//   overlap the CG work of chunk k.  mbarriers: full[b] (producer arrive + tcgen05.commit), empty[b]
//   (16 consumer warps), bfree (tcgen05.commit: B operand / MLP scratch reusable).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "generated/cg_gen.cuh"

namespace mt {

constexpr int kTcNE = 64;             // edges per chunk == MMA N
constexpr int kTcK = 32;              // padded size of the last hidden layer == MMA K total
constexpr int kTcProducerWarps = 4;
constexpr int kTcConsumerWarps = 16;
constexpr int kTcThreads = 32 * (kTcProducerWarps + kTcConsumerWarps);
constexpr int kTcMaxTiles = 4;        // M tiles of 128 rows -> <= 512 weight-column rows
constexpr int kTcMaxSub = 64;         // sub-items per plan

struct ConvTcParams {
  int x_dim, y_dim, out_dim;
  int num_tiles;             // MT
  int num_sub;               // sub-items
  const int32_t* row_wcol;   // [MT*128] weight column of every A row (-1: zero row)
  const int32_t* sub_hdr;    // [num_sub][8] {type, cpw, lane0 (first TMEM lane within the quarter), tile, quarter, 0,0,0}
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
  int xs_stride;  // floats, multiple of 4 when x_dim % 4 == 0
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
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
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
  v[2] = __uint_as_float(r2);
  v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

struct __align__(16) TcMeta {
  int nnodes;  // -1: terminate
  int ne;      // staged edges
  int c0;      // first edge (global, receiver-sorted order)
  int pad;
  int node_id[kTcNE];
  short e0[kTcNE];  // chunk-relative edge range of each node
  short e1[kTcNE];
  unsigned char first[kTcNE];  // 1: this chunk holds the node's first edges (plain store), 0: accumulate
};

// shared-memory carve-up (bytes), all regions 1024-aligned where the tensor core reads them
struct TcSmemLayout {
  size_t a_off, a_plane;   // 3 planes of [MT*128 x 32] bf16
  size_t b_off, b_plane;   // 3 planes of [64 x 32] bf16 (also MLP scratch P0)
  size_t p1_off;           // MLP scratch P1 [64][32] fp32
  size_t x_off, x_buf;     // 2 buffers [64][xs_stride] fp32
  size_t y_off, y_buf;     // 2 buffers [64][y_dim] fp32
  size_t meta_off;         // 2 TcMeta
  size_t total;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int MT, int xs_stride, int y_dim) {
  TcSmemLayout L;
  size_t o = 0;
  L.a_off = o;
  L.a_plane = (size_t)MT * 128 * kTcK * 2;
  o += 3 * L.a_plane;
  L.b_off = o;
  L.b_plane = (size_t)kTcNE * kTcK * 2;
  o += 3 * L.b_plane;
  L.p1_off = o;
  o += (size_t)kTcNE * kTcK * 4;
  L.x_off = o;
  L.x_buf = (size_t)kTcNE * xs_stride * 4;
  o += 2 * L.x_buf;
  L.y_off = o;
  L.y_buf = ((size_t)kTcNE * y_dim * 4 + 15) & ~(size_t)15;
  o += 2 * L.y_buf;
  L.meta_off = o;
  o += 2 * sizeof(TcMeta);
  L.total = o;
  return L;
}

// one (sub-item, node) unit: edges [e0, e1) of the chunk (columns of the TMEM tile)
template <int L1, int L2, int L3>
__device__ __forceinline__ void tc_unit(uint32_t taddr_row, int lane, int cpw, int src_lane_base,
                                        const float* __restrict__ xs, int xstride, const float* __restrict__ ys,
                                        int ystride, int xoff, int yoff, int e0, int e1, float* __restrict__ o,
                                        float den, bool first, bool valid) {
  constexpr int D1 = 2 * L1 + 1, D2 = 2 * L2 + 1, D3 = 2 * L3 + 1;
  float acc[D3];
#pragma unroll
  for (int m = 0; m < D3; ++m) acc[m] = 0.f;
  const int nphase = 32 / cpw;
  const int phase = lane / cpw;
  const int src_lane = (src_lane_base + (lane & (cpw - 1))) & 31;
  for (int grp = e0 & ~3; grp < e1; grp += 4) {
    float v[4];
    tmem_ld_x4(taddr_row + (uint32_t)grp, v);
    if (nphase == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int el = grp + i;
        if (el >= e0 && el < e1) {  // warp-uniform
          float xv[D1], yv[D2];
          const float* xr = xs + (size_t)el * xstride + xoff;
#pragma unroll
          for (int m = 0; m < D1; ++m) xv[m] = xr[m];
          const float* yr = ys + (size_t)el * ystride + yoff;
#pragma unroll
          for (int m = 0; m < D2; ++m) yv[m] = yr[m];
          CG<L1, L2, L3>::template fwd<float>(xv, yv, v[i], acc);
        }
      }
    } else {
      // packed small types: lane = (column j, phase p); phase p takes edge grp + r*nphase + p.
      // its weight sits in TMEM lane src_lane (another lane of this warp) -> shuffle.
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = __shfl_sync(0xffffffffu, v[i], src_lane);
      for (int r = 0; r < 4; r += nphase) {
        const int k = r + phase;  // nphase in {2,4}: k in 0..3
        const int el = grp + k;
        const float w = (k == 0) ? t[0] : (k == 1) ? t[1] : (k == 2) ? t[2] : t[3];
        if (el >= e0 && el < e1) {
          float xv[D1], yv[D2];
          const float* xr = xs + (size_t)el * xstride + xoff;
#pragma unroll
          for (int m = 0; m < D1; ++m) xv[m] = xr[m];
          const float* yr = ys + (size_t)el * ystride + yoff;
#pragma unroll
          for (int m = 0; m < D2; ++m) yv[m] = yr[m];
          CG<L1, L2, L3>::template fwd<float>(xv, yv, w, acc);
        }
      }
    }
  }
  for (int off = cpw; off < 32; off <<= 1) {
#pragma unroll
    for (int m = 0; m < D3; ++m) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], off);
  }
  if (valid && phase == 0) {
#pragma unroll
    for (int m = 0; m < D3; ++m) {
      const float r = acc[m] / den;
      o[m] = first ? r : (o[m] + r);
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) conv_fwd_tc_kernel(const ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_bfree;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_cnt[2][4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int MT = p.num_tiles;
  const TcSmemLayout L = tc_smem_layout(MT, p.xs_stride, p.y_dim);
  unsigned char* sA = smem + L.a_off;
  unsigned char* sB = smem + L.b_off;
  float* sP0 = reinterpret_cast<float*>(sB);  // aliases the B planes (12 KB >= 8 KB)
  float* sP1 = reinterpret_cast<float*>(smem + L.p1_off);
  TcMeta* meta = reinterpret_cast<TcMeta*>(smem + L.meta_off);
  const uint32_t tmem_cols = (2 * MT * kTcNE <= 32) ? 32 : (2 * MT * kTcNE <= 64) ? 64 : (2 * MT * kTcNE <= 128) ? 128
                             : (2 * MT * kTcNE <= 256) ? 256 : 512;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    mbar_init(&bar_full[0], 2);
    mbar_init(&bar_full[1], 2);
    mbar_init(&bar_empty[0], kTcConsumerWarps);
    mbar_init(&bar_empty[1], kTcConsumerWarps);
    mbar_init(&bar_bfree, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem_base, tmem_cols);
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
        float v = (wc >= 0 && k < H) ? Wl[(size_t)k * Wn + wc] * s : 0.f;
        hi[i] = __float2bfloat16_rn(v);
        float r1 = v - __bfloat162float(hi[i]);
        mi[i] = __float2bfloat16_rn(r1);
        float r2 = r1 - __bfloat162float(mi[i]);
        lo[i] = __float2bfloat16_rn(r2);
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

  // node range of this CTA: boundaries at equal shares of the edge list
  int64_t n_begin, n_end;
  {
    auto bound = [&](int64_t i) -> int64_t {  // first node whose rowptr >= i*E/grid
      if (i <= 0) return 0;
      if (i >= (int64_t)gridDim.x) return p.N;
      const int64_t target = (p.E * i) / gridDim.x;
      int64_t lo = 0, hi = p.N;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (p.rowptr[mid] < target) lo = mid + 1; else hi = mid;
      }
      return lo;
    };
    if (p.E == 0) {
      n_begin = (p.N * blockIdx.x) / gridDim.x;
      n_end = (p.N * (blockIdx.x + 1)) / gridDim.x;
    } else {
      n_begin = bound(blockIdx.x);
      n_end = bound(blockIdx.x + 1);
    }
  }

  if (warp < kTcProducerWarps) {
    // ================================================================ producers
    const int ptid = tid;  // 0..127
    constexpr int NP = kTcProducerWarps * 32;
    int64_t n_cur = n_begin;
    int e_carry = -1;  // >= 0: next chunk continues node n_cur at this global edge
    int mma_issued = 0;
    const uint32_t idesc = make_idesc_bf16(128, kTcNE);
    for (int k = 0;; ++k) {
      const int b = k & 1;
      if (ptid == 0) {
        mbar_wait(&bar_empty[b], ((k >> 1) & 1) ^ 1);             // consumers released buffer b
        if (mma_issued > 0) mbar_wait(&bar_bfree, (mma_issued - 1) & 1);  // B / scratch reusable
        // ---- build the chunk: whole nodes while they fit; a node with > 64 edges is split
        TcMeta& M = meta[b];
        int nn = 0, ne = 0, c0 = 0;
        if (n_cur >= n_end) {
          nn = -1;
        } else {
          c0 = (e_carry >= 0) ? e_carry : p.rowptr[n_cur];
          while (n_cur < n_end && nn < kTcNE) {
            const int r0 = (e_carry >= 0) ? e_carry : p.rowptr[n_cur];
            const int r1 = p.rowptr[n_cur + 1];
            const int deg = r1 - r0;
            if (deg <= kTcNE - ne) {
              M.node_id[nn] = (int)n_cur;
              M.e0[nn] = (short)ne;
              M.e1[nn] = (short)(ne + deg);
              M.first[nn] = (e_carry < 0) ? 1 : 0;
              ne += deg;
              ++nn;
              ++n_cur;
              e_carry = -1;
            } else if (ne == 0) {  // node larger than a chunk: take 64 edges, continue next time
              M.node_id[nn] = (int)n_cur;
              M.e0[nn] = 0;
              M.e1[nn] = (short)kTcNE;
              M.first[nn] = (e_carry < 0) ? 1 : 0;
              ne = kTcNE;
              ++nn;
              e_carry = r0 + kTcNE;
              break;
            } else {
              break;
            }
          }
        }
        M.nnodes = nn;
        M.ne = ne;
        M.c0 = c0;
        // a chunk that continues a node split over chunks accumulates into its output row: keep the
        // pieces ordered (deterministic sum) by letting the previous chunk drain first
        if (nn > 0 && M.first[0] == 0 && k > 0) mbar_wait(&bar_empty[b ^ 1], ((k - 1) >> 1) & 1);
      }
      named_bar_sync(1, NP);
      const int nn = meta[b].nnodes, ne = meta[b].ne, c0 = meta[b].c0;
      if (nn < 0) {
        if (ptid == 0) {
          mbar_arrive(&bar_full[b]);
          mbar_arrive(&bar_full[b]);
        }
        break;
      }
      float* xs = reinterpret_cast<float*>(smem + L.x_off + (size_t)b * L.x_buf);
      float* ys = reinterpret_cast<float*>(smem + L.y_off + (size_t)b * L.y_buf);
      if (ne > 0) {
        // ---- gather (cp.async): sender rows, sh rows, radial embedding rows
        if ((p.x_dim & 3) == 0) {
          const int vec = p.x_dim >> 2;
          for (int t = ptid; t < ne * vec; t += NP) {
            const int el = t / vec, j = t - el * vec;
            cp_async_16(xs + (size_t)el * p.xs_stride + 4 * j, p.x + (size_t)p.src[c0 + el] * p.x_dim + 4 * j);
          }
        } else {
          for (int t = ptid; t < ne * p.x_dim; t += NP) {
            const int el = t / p.x_dim, j = t - el * p.x_dim;
            cp_async_4(xs + (size_t)el * p.xs_stride + j, p.x + (size_t)p.src[c0 + el] * p.x_dim + j);
          }
        }
        for (int t = ptid; t < ne * p.y_dim; t += NP) {
          const int el = t / p.y_dim, j = t - el * p.y_dim;
          cp_async_4(ys + t, p.sh + (size_t)p.perm[c0 + el] * p.y_dim + j);
        }
        const int in0 = p.sizes[0];
        for (int t = ptid; t < kTcNE * kTcK; t += NP) {
          const int el = t >> 5, j = t & 31;
          if (el < ne && j < in0) cp_async_4(sP0 + t, p.emb + (size_t)p.perm[c0 + el] * in0 + j);
          else sP0[t] = 0.f;
        }
        cp_async_wait_all();
        named_bar_sync(1, NP);
        // ---- hidden layers: P[li&1] -> P[(li+1)&1], task = (edge, 8 outputs)
        for (int li = 0; li + 1 < p.nl; ++li) {
          const float* hin = (li & 1) ? sP1 : sP0;
          float* hout = (li & 1) ? sP0 : sP1;
          const int fi = p.sizes[li], fo = p.sizes[li + 1];
          const float* __restrict__ Wl = p.w[li];
          const float s = rsqrtf((float)fi);
          for (int t = ptid; t < kTcNE * 4; t += NP) {
            const int el = t >> 2, j0 = (t & 3) * 8;
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = 0.f;
            const float* hr = hin + el * kTcK;
            if ((fo & 3) == 0 && j0 + 8 <= fo) {
              for (int kk = 0; kk < fi; ++kk) {
                const float hv = hr[kk];
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wl + (size_t)kk * fo + j0));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wl + (size_t)kk * fo + j0 + 4));
                a[0] = fmaf(hv, w0.x, a[0]); a[1] = fmaf(hv, w0.y, a[1]);
                a[2] = fmaf(hv, w0.z, a[2]); a[3] = fmaf(hv, w0.w, a[3]);
                a[4] = fmaf(hv, w1.x, a[4]); a[5] = fmaf(hv, w1.y, a[5]);
                a[6] = fmaf(hv, w1.z, a[6]); a[7] = fmaf(hv, w1.w, a[7]);
              }
            } else {
              for (int kk = 0; kk < fi; ++kk) {
                const float hv = hr[kk];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float wv = (j0 + i < fo) ? __ldg(Wl + (size_t)kk * fo + j0 + i) : 0.f;
                  a[i] = fmaf(hv, wv, a[i]);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              hout[el * kTcK + j0 + i] = (j0 + i < fo) ? apply_act<float>(p.act, a[i] * s) * p.act_cst : 0.f;
          }
          named_bar_sync(1, NP);
        }
        // ---- B planes: h (fp32, in P[(nl-1)&1]) -> bf16 hi/mid/lo; via registers because P0 aliases B
        {
          const float* hfin = ((p.nl - 1) & 1) ? sP1 : sP0;
          float hv[2][8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int t = ptid + q * NP;
            const int el = t >> 2, g = t & 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[q][i] = hfin[el * kTcK + g * 8 + i];
          }
          named_bar_sync(1, NP);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int t = ptid + q * NP;
            const int el = t >> 2, g = t & 3;
            __align__(16) __nv_bfloat16 hi[8], mi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float v = hv[q][i];
              hi[i] = __float2bfloat16_rn(v);
              const float r1 = v - __bfloat162float(hi[i]);
              mi[i] = __float2bfloat16_rn(r1);
              const float r2 = r1 - __bfloat162float(mi[i]);
              lo[i] = __float2bfloat16_rn(r2);
            }
            const size_t off = (size_t)g * kTcNE * 16 + (size_t)el * 16;
            *reinterpret_cast<uint4*>(sB + off) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(sB + L.b_plane + off) = *reinterpret_cast<const uint4*>(mi);
            *reinterpret_cast<uint4*>(sB + 2 * L.b_plane + off) = *reinterpret_cast<const uint4*>(lo);
          }
        }
        fence_proxy_async();
      }
      named_bar_sync(1, NP);
      if (ptid == 0) {
        s_cnt[b][0] = 0; s_cnt[b][1] = 0; s_cnt[b][2] = 0; s_cnt[b][3] = 0;
        if (ne > 0) {
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
          const uint32_t a_lbo = (uint32_t)MT * 128 * 16, b_lbo = (uint32_t)kTcNE * 16;
          // significant products of (hi+mid+lo) x (hi+mid+lo), small ones first
          const int pa[6] = {0, 2, 1, 0, 1, 0};
          const int pb[6] = {2, 0, 1, 1, 0, 0};
          for (int t = 0; t < MT; ++t) {
            const uint32_t d = tmem_base + (uint32_t)((b * MT + t) * kTcNE);
            uint32_t accum = 0;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t a_addr = a_base + (uint32_t)(pa[q] * L.a_plane) + (uint32_t)(2 * ks) * a_lbo + (uint32_t)t * 128 * 16;
                const uint32_t b_addr = b_base + (uint32_t)(pb[q] * L.b_plane) + (uint32_t)(2 * ks) * b_lbo;
                umma_bf16(d, make_kmajor_desc(a_addr, a_lbo, 128), make_kmajor_desc(b_addr, b_lbo, 128), idesc, accum);
                accum = 1;
              }
            }
          }
          umma_commit(&bar_full[b]);
          umma_commit(&bar_bfree);
          ++mma_issued;
        } else {
          mbar_arrive(&bar_full[b]);
        }
        mbar_arrive(&bar_full[b]);
      }
    }
  } else {
    // ================================================================ consumers
    const int q = warp & 3;
    const int nsubq = p.q_count[q];
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    for (int k = 0;; ++k) {
      const int b = k & 1;
      mbar_wait(&bar_full[b], (k >> 1) & 1);
      tc_fence_after();
      const TcMeta& M = meta[b];
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
        const int sub = p.q_list[q * kTcMaxSub + si];
        const int4 h0 = reinterpret_cast<const int4*>(p.sub_hdr)[sub * 2 + 0];
        const int4 h1 = reinterpret_cast<const int4*>(p.sub_hdr)[sub * 2 + 1];
        const int type = h0.x, cpw = h0.y, lane0 = h0.z, tile = h0.w;
        (void)h1;
        const int4 slot = reinterpret_cast<const int4*>(p.sub_slot)[sub * 32 + lane];
        const int node = M.node_id[nj];
        const int e0 = M.e0[nj], e1 = M.e1[nj];
        const bool first = M.first[nj] != 0;
        const float den = p.num_neigh ? sqrtf(p.num_neigh[node]) : sqrtf(p.avg);
        float* o = p.out + (size_t)node * p.out_dim + slot.z;
        const uint32_t taddr = tmem_base + lane_base + (uint32_t)((b * MT + tile) * kTcNE);
        switch (type) {
#define MT_TC_CASE(ID, A, B, C)                                                                                 \
  case ID:                                                                                                      \
    tc_unit<A, B, C>(taddr, lane, cpw, lane0, xs, p.xs_stride, ys, p.y_dim, slot.x, slot.y, e0, e1, o, den, first, \
                     slot.w != 0);                                                                              \
    break;
          MT_FOR_EACH_CG_TYPE(MT_TC_CASE)
#undef MT_TC_CASE
          default: break;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[b]);
    }
  }
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mt
