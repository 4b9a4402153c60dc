// Node-side kernels: species-indexed irreps linear (a8/a11/a13/a14), Gate + BatchNorm
// affine (a9/a10), segmented pooling (a12).  See include/matten_b200.h for the contract.
#include "common.cuh"

namespace mt {

// =========================================================================
// irreps-wise (species-indexed) linear
//   reference: e3nn FullyConnectedTensorProduct(x, one_hot, out) at
//   src/matten/nn/conv.py:59-61,77-79,84-86 and e3nn.o3.Linear at
//   src/matten/nn/nodewise.py:111.  With one-hot attributes the "uvw" tensor product is
//   a per-species dense matrix per irrep type; nodes are grouped by species so each CTA
//   multiplies a [rows x mul_in] tile (rows = (node, m)) by ONE [mul_in x mul_out]
//   weight slice staged in shared memory (register-tiled 4x4 FMA micro-kernel).
// =========================================================================
constexpr int kLinMaxBlocks = 24;
constexpr int kLinStage = 4096;  // x elements staged per k-chunk: ROWS x KC
constexpr int kLinMaxRows = 1024;
constexpr int kLinMaxDim = 9;    // 2l+1 for l <= 4 takes the fast tables; larger dims fall back to divisions
// Tile shapes (256 threads, thread tile TM rows x 4 columns):
//   COLS  TM  ROWS  KC     stage elems 2 x (KC x (ROWS+4) + KC x COLS)     output tile ROWS x (COLS+1)
//    64    8   128  32     12544                                            8320
//    32    8   256  16      9344                                            8448
//    16    8   512   8      8512                                            8704
//     8    4   512   8      8384                                            4608
//     4    4  1024   4      8256                                            5120
constexpr int kLinSmemElems = 12544;

template <typename T>
__device__ __forceinline__ void lin_ld4(const T* p, T (&v)[4]);
template <>
__device__ __forceinline__ void lin_ld4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void lin_ld4<double>(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// one element global -> shared without passing through registers (LDGSTS); !valid zero-fills the destination
template <typename T>
__device__ __forceinline__ void lin_cp_async(T* smem_dst, const T* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? (int)sizeof(T) : 0;
  if constexpr (sizeof(T) == 4)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void lin_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lin_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct LinParams {
  int num_blocks;
  int32_t in_off[kLinMaxBlocks], out_off[kLinMaxBlocks], mul_in[kLinMaxBlocks], mul_out[kLinMaxBlocks],
      dim[kLinMaxBlocks], w_off[kLinMaxBlocks];
  double scale[kLinMaxBlocks];
  int32_t cta_begin[kLinMaxBlocks + 1];  // prefix of CTA counts per block
  int32_t col_chunks[kLinMaxBlocks];     // ceil(mul_out / cols)
  int32_t nodes_per_tile[kLinMaxBlocks];
  int32_t tile_cols[kLinMaxBlocks];      // 64 / 32 / 16 / 8 / 4
  int in_dim, out_dim, S;
  const void* x;
  const void* weight;
  const int32_t* sperm;
  const int32_t* sptr;
  int accumulate;
  int transpose;  // 1: apply the transposed weight slices (backward w.r.t. x): blocks come with in/out swapped
  void* out;
  int64_t N;
};

struct LinTile {
  int b, s, c0, ncols, tn, mi, mo, d, COLS, lc, ROWS, KC, XS;
};

// k loop + epilogue of one tile with a TM x 4 register tile per thread.  The k-chunks are double buffered in shared
// memory and filled by cp.async one chunk ahead (round 1's kernel staged through registers with an integer division
// per element and exposed the global latency: ncu r2_step_full, 38 % ALU pipe, 45 % long-scoreboard stalls, 8 % of
// the instructions useful FMAs).  All index arithmetic in the loops is additive; the two (q / d, q % d) maps come from
// tables filled once per CTA.
template <typename T, int TM>
__device__ __forceinline__ void lin_tile_run(const LinParams& p, const LinTile& t, T* smem, const int* s_nodes,
                                             const int* qmap, const int* omap) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int COLS = t.COLS, KC = t.KC, XS = t.XS, d = t.d, mi = t.mi;
  const int stage_elems = KC * XS + KC * COLS;
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ W = static_cast<const T*>(p.weight);
  T* __restrict__ OUT = static_cast<T*>(p.out);
  const int tcn = COLS >> 2;
  const int tc = tid % tcn, tr = tid / tcn;
  using P = typename pair_of<T>::type;  // column pairs: one FFMA2 per pair in fp32
  P acc[TM][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[i][j].x = T(0); acc[i][j].y = T(0); }

  const int nchunks = (mi + KC - 1) / KC;
  const int kcd = KC * d;
  auto issue = [&](int c) {
    T* xsT = smem + (size_t)(c & 1) * stage_elems;  // [KC][XS]
    T* ws = xsT + KC * XS;                          // [KC][COLS]
    const int u0 = c * KC;
    const int ku = min(KC, mi - u0);
    const int seg = ku * d;
    // x rows: warps over nodes, lanes over the node's contiguous [u0 * d, (u0 + KC) * d) segment (coalesced)
    for (int j = warp; j < t.tn; j += 8) {
      const T* src = X + (size_t)s_nodes[j] * p.in_dim + p.in_off[t.b] + u0 * d;
      T* dst = xsT + j * d;
      for (int q = lane; q < kcd; q += 32) lin_cp_async<T>(dst + qmap[q], src + (q < seg ? q : 0), q < seg);
    }
    // weight slice ws[uu][c] = W[w_off + ((u0+uu)*S + s)*mo + c0 + c]  (transposed: roles of u and c swapped)
    for (int e = tid; e < KC * COLS; e += 256) {
      const int uu = e >> t.lc, c = e & (COLS - 1);
      const bool ok = uu < ku && c < t.ncols;
      const size_t off = p.transpose ? (size_t)p.w_off[t.b] + ((size_t)(t.c0 + c) * p.S + t.s) * mi + (u0 + uu)
                                     : (size_t)p.w_off[t.b] + ((size_t)(u0 + uu) * p.S + t.s) * t.mo + t.c0 + c;
      lin_cp_async<T>(ws + e, W + (ok ? off : 0), ok);
    }
    lin_cp_commit();
  };

  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      issue(c + 1);
      lin_cp_wait<1>();
    } else {
      lin_cp_wait<0>();
    }
    __syncthreads();
    const T* xsT = smem + (size_t)(c & 1) * stage_elems;
    const T* ws = xsT + KC * XS;
    const T* ap = xsT + tr * TM;
    const T* bp = ws + tc * 4;
#pragma unroll 4
    for (int uu = 0; uu < KC; ++uu) {
      T a[TM], bb[4];
#pragma unroll
      for (int i = 0; i < TM; i += 4) lin_ld4<T>(ap + uu * XS + i, *reinterpret_cast<T(*)[4]>(&a[i]));
      lin_ld4<T>(bp + uu * COLS, bb);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        fma_pair(a[i], bb[0], bb[1], acc[i][0]);
        fma_pair(a[i], bb[2], bb[3], acc[i][1]);
      }
    }
    __syncthreads();  // the buffer is refilled two chunks later
  }
  const T scale = T(p.scale[t.b]);
  const int OS = COLS + 1;
  T* ot = smem;  // [ROWS][COLS + 1], aliases the stages (all reads are behind the barrier above)
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ot[(tr * TM + i) * OS + tc * 4 + 2 * j] = acc[i][j].x * scale;
      ot[(tr * TM + i) * OS + tc * 4 + 2 * j + 1] = acc[i][j].y * scale;
    }
  __syncthreads();
  // coalesced write: per node the (w, m) range is contiguous
  const int span = t.ncols * d;
  for (int j = warp; j < t.tn; j += 8) {
    T* dst = OUT + (size_t)s_nodes[j] * p.out_dim + p.out_off[t.b] + t.c0 * d;
    const T* src = ot + j * d * OS;
    for (int q = lane; q < span; q += 32) {
      const T v = src[omap[q]];
      dst[q] = p.accumulate ? (dst[q] + v) : v;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) linear_fwd_kernel(const LinParams p) {
  extern __shared__ __align__(16) unsigned char lin_smem[];
  __shared__ int s_nodes[kLinMaxRows];
  __shared__ int s_qmap[32 * kLinMaxDim];  // q -> (q / d) * XS + q % d        (x staging, q < KC * d)
  __shared__ int s_omap[64 * kLinMaxDim];  // q -> (q % d) * (COLS + 1) + q / d (output, q < ncols * d)

  // which block / node tile / column chunk am I?
  int b = 0;
  while (b + 1 < p.num_blocks && (int)blockIdx.x >= p.cta_begin[b + 1]) ++b;
  int local = blockIdx.x - p.cta_begin[b];
  const int cc = local % p.col_chunks[b];
  int tile = local / p.col_chunks[b];
  const int TN = p.nodes_per_tile[b];
  LinTile t;
  t.b = b;
  t.mi = p.mul_in[b]; t.mo = p.mul_out[b]; t.d = p.dim[b];
  t.COLS = p.tile_cols[b];
  t.lc = 31 - __clz(t.COLS);
  const int TM = t.COLS >= 16 ? 8 : 4;
  t.ROWS = (256 / (t.COLS >> 2)) * TM;
  t.KC = kLinStage / t.ROWS;
  t.XS = t.ROWS + 4;  // row stride of the transposed x stage
  // locate (species, tile-in-species)
  int s = 0;
  int64_t begin = 0, end = 0;
  if (p.S == 1 && p.sptr == nullptr) {
    begin = (int64_t)tile * TN;
    end = imin64(begin + TN, p.N);
    if (begin >= p.N) return;
  } else {
    bool found = false;
    for (s = 0; s < p.S; ++s) {
      int cnt = p.sptr[s + 1] - p.sptr[s];
      int nt = (cnt + TN - 1) / TN;
      if (tile < nt) {
        begin = p.sptr[s] + (int64_t)tile * TN;
        end = imin64(begin + TN, (int64_t)p.sptr[s + 1]);
        found = true;
        break;
      }
      tile -= nt;
    }
    if (!found) return;
  }
  t.s = s;
  t.tn = (int)(end - begin);
  t.c0 = cc * t.COLS;
  t.ncols = min(t.COLS, t.mo - t.c0);
  const int tid = threadIdx.x;
  const int d = t.d;
  for (int i = tid; i < t.tn; i += blockDim.x) s_nodes[i] = p.sperm ? p.sperm[begin + i] : (int)(begin + i);
  for (int q = tid; q < t.KC * d; q += blockDim.x) s_qmap[q] = (q / d) * t.XS + q % d;
  for (int q = tid; q < t.ncols * d; q += blockDim.x) s_omap[q] = (q % d) * (t.COLS + 1) + q / d;
  __syncthreads();

  if (t.mi == 0) {  // irreps with no incoming path: zeros
    if (!p.accumulate) {
      T* __restrict__ OUT = static_cast<T*>(p.out);
      const int span = t.ncols * d;
      for (int j = tid >> 5; j < t.tn; j += 8)
        for (int q = tid & 31; q < span; q += 32)
          OUT[(size_t)s_nodes[j] * p.out_dim + p.out_off[b] + t.c0 * d + q] = T(0);
    }
    return;
  }
  T* smem = reinterpret_cast<T*>(lin_smem);
  if (TM == 8) lin_tile_run<T, 8>(p, t, smem, s_nodes, s_qmap, s_omap);
  else lin_tile_run<T, 4>(p, t, smem, s_nodes, s_qmap, s_omap);
}

// =========================================================================
// Node-streaming linear: one CTA owns a tile of nodes of ONE species and walks ALL irrep blocks over it.
//   * the species' weight slices of every block sit in shared memory for the CTA's lifetime;
//   * each block's input columns of the tile's rows arrive by cp.async (16 B per lane, coalesced along the row), one
//     block ahead of the arithmetic, so x is read from HBM exactly once and never passes through registers;
//   * lanes <-> output channels w, NB nodes per lane in registers: per 4 input channels a lane issues 4 LDS.32 of
//     weights (conflict free), NB x D LDS.128 of x (warp-wide broadcast) and NB x 4 x D FMAs.
// The tiled kernel above spends 152 k thread instructions per node on lin2 of the last layer for 14.8 k MACs (rows
// of 1392 floats cut into 80-byte slivers per k-chunk, every chunk paying the DRAM latency): ncu r2_step_full.
// =========================================================================
constexpr int kStreamSmemBytes = 200 * 1024;

struct LinStream {
  int tnode;                        // nodes per CTA tile
  int rs;                           // row stride (elements) of the staged x rows: max over blocks, multiple of 4, + 4
  int w_smem_off[kLinMaxBlocks];    // element offset of block b's [mi4][mo] slice
  int w_total;                      // elements of all slices
  int vec_ok[kLinMaxBlocks];        // rows of this block can be copied 16 bytes at a time
  int w_vec[kLinMaxBlocks];         // ... and so can the rows of its weight slice
  int tiles_per_cta;                // consecutive tiles of one species per CTA (weights staged once for all of them)
};

// Lanes of a warp = (output channel w: lpn lanes) x (k-split ks: 32 / lpn lanes): every warp works on NB nodes
// whatever mul_out is; narrow outputs (mul_out 16 or 4 for l >= 1) split the input channels over the spare lanes and
// reduce with shuffles at the end.
template <typename T, int D, int NB>
__device__ __forceinline__ void lin_stream_block(const T* __restrict__ ws, const T* __restrict__ xs, int RS, int mi4,
                                                 int mo, int tn, const int* __restrict__ s_nodes,
                                                 T* __restrict__ OUT, int out_dim, int out_off, T scale,
                                                 int accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lpn = 32;  // lanes per channel group: the smallest of 32/16/8/4 covering mul_out (32 when mul_out > 16)
  while (lpn > 4 && (lpn >> 1) >= mo) lpn >>= 1;
  const int subs = 32 / lpn, ks = lane / lpn, wl = lane - ks * lpn;
  const int ustep = 4 * subs;
  for (int w0 = 0; w0 < mo; w0 += lpn) {
    const int w = w0 + wl;
    const bool wok = w < mo;
    for (int n0 = warp * NB; n0 < tn; n0 += 8 * NB) {
      T acc[NB][D];
      const T* xq[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
        for (int m = 0; m < D; ++m) acc[nb][m] = T(0);
        // rows past the tile repeat the last one; never stored
        xq[nb] = xs + (size_t)min(n0 + nb, tn - 1) * RS + ks * 4 * D;
      }
      const T* wq = ws + (wok ? w : 0) + ks * 4 * mo;
      for (int u = ks * 4; u < mi4; u += ustep) {
        T wv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wv[k] = wq[k * mo];
        wq += ustep * mo;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          T xv[4 * D];
#pragma unroll
          for (int i = 0; i < D; ++i) lin_ld4<T>(xq[nb] + 4 * i, *reinterpret_cast<T(*)[4]>(&xv[4 * i]));
          xq[nb] += ustep * D;
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int m = 0; m < D; ++m) acc[nb][m] = fma(wv[k], xv[k * D + m], acc[nb][m]);
        }
      }
      for (int off = lpn; off < 32; off <<= 1) {  // sum the k-splits (fixed order: deterministic)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int m = 0; m < D; ++m) acc[nb][m] += __shfl_xor_sync(0xffffffffu, acc[nb][m], off);
      }
      if (wok && ks == 0) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          if (n0 + nb < tn) {
            T* dst = OUT + (size_t)s_nodes[n0 + nb] * out_dim + out_off + w * D;
#pragma unroll
            for (int m = 0; m < D; ++m) dst[m] = accumulate ? (dst[m] + acc[nb][m] * scale) : acc[nb][m] * scale;
          }
        }
      }
    }
  }
}

// CTA-synchronous variant (all warps walk the blocks together over a tile of 16 or 32 nodes, NB = 2 / 4 nodes per
// lane): better for the short rows of lin1 / sc, where amortising the weight loads over more nodes matters most.
constexpr int kStreamTiles = 8;  // at most: consecutive tiles of one species per CTA (weight slices staged once)

template <typename T>
__global__ void __launch_bounds__(256) linear_sync_kernel(const LinParams p, const LinStream e) {
  extern __shared__ __align__(16) unsigned char lin_smem[];
  __shared__ int s_nodes[kStreamTiles * 32];
  T* wsm = reinterpret_cast<T*>(lin_smem);
  T* xbuf = wsm + ((e.w_total + 3) & ~3);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int super = e.tnode * e.tiles_per_cta;
  // locate (species, super-tile-in-species)
  int tile = blockIdx.x, s = 0;
  int64_t begin = 0, end = 0;
  if (p.S == 1 && p.sptr == nullptr) {
    begin = (int64_t)tile * super;
    end = imin64(begin + super, p.N);
    if (begin >= p.N) return;
  } else {
    bool found = false;
    for (s = 0; s < p.S; ++s) {
      const int cnt = p.sptr[s + 1] - p.sptr[s];
      const int nt = (cnt + super - 1) / super;
      if (tile < nt) {
        begin = p.sptr[s] + (int64_t)tile * super;
        end = imin64(begin + super, (int64_t)p.sptr[s + 1]);
        found = true;
        break;
      }
      tile -= nt;
    }
    if (!found) return;
  }
  const int tn_all = (int)(end - begin);
  const int ntiles = (tn_all + e.tnode - 1) / e.tnode;
  for (int i = tid; i < tn_all; i += 256) s_nodes[i] = p.sperm ? p.sperm[begin + i] : (int)(begin + i);
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ W = static_cast<const T*>(p.weight);
  T* __restrict__ OUT = static_cast<T*>(p.out);
  constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte copy
  // this species' weight slices, [mi4][mo] per block (rows mi..mi4 zero); asynchronous copies, all in flight at once
  // (the first group the pipeline below waits for).  Rows of mo contiguous elements: 16 bytes per copy when mo allows.
  for (int b = 0; b < p.num_blocks; ++b) {
    const int mi = p.mul_in[b], mo = p.mul_out[b], mi4 = (mi + 3) & ~3;
    T* dst = wsm + e.w_smem_off[b];
    if (!p.transpose && e.w_vec[b]) {
      const int rowv = mo / V;  // vectors per row
      const int nvec = mi4 * rowv;
      int u = tid / rowv, c = tid - u * rowv;          // one division per thread and block; then additive
      const int du = 256 / rowv, dc = 256 - du * rowv;
      for (int v = tid; v < nvec; v += 256) {
        const bool ok = u < mi;
        const T* src = W + (ok ? (size_t)p.w_off[b] + ((size_t)u * p.S + s) * mo + c * V : 0);
        const uint32_t da = (uint32_t)__cvta_generic_to_shared(dst + u * mo + c * V);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(src), "r"(ok ? 16 : 0) : "memory");
        u += du; c += dc;
        if (c >= rowv) { c -= rowv; ++u; }
      }
    } else {
      for (int u = warp; u < mi4; u += 8)
        for (int w = lane; w < mo; w += 32) {
          const bool ok = u < mi;
          const size_t off = !ok ? 0
                             : p.transpose ? (size_t)p.w_off[b] + ((size_t)w * p.S + s) * mi + u
                                           : (size_t)p.w_off[b] + ((size_t)u * p.S + s) * mo + w;
          lin_cp_async<T>(dst + u * mo + w, W + off, ok);
        }
    }
  }
  lin_cp_commit();
  __syncthreads();  // s_nodes
  const int nb_ = p.num_blocks;
  const int nsteps = ntiles * nb_;
  const uint32_t xbuf_s = (uint32_t)__cvta_generic_to_shared(xbuf);
  auto issue = [&](int step) {
    const int t = step / nb_, b = step - t * nb_;
    const int mi = p.mul_in[b], d = p.dim[b];
    const int seg = mi * d, segp = (((mi + 3) & ~3) * d + 3) & ~3;  // copied + zero-filled up to the padded length
    const int n0 = t * e.tnode, tn = min(e.tnode, tn_all - n0);
    const uint32_t buf_s = xbuf_s + (uint32_t)((step & 1) * e.tnode * e.rs * (int)sizeof(T));
    if (mi > 0) {
      for (int j = warp; j < tn; j += 8) {
        const T* src = X + (size_t)s_nodes[n0 + j] * p.in_dim + p.in_off[b];
        const uint32_t drow = buf_s + (uint32_t)(j * e.rs * (int)sizeof(T));
        if (e.vec_ok[b]) {
          for (int q = lane * V; q < segp; q += 32 * V) {
            const int n = min(V, seg - q);  // elements really there
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + q * (int)sizeof(T)),
                         "l"(src + (n > 0 ? q : 0)), "r"(n > 0 ? n * (int)sizeof(T) : 0)
                         : "memory");
          }
        } else {
          T* dst = xbuf + (size_t)(step & 1) * e.tnode * e.rs + (size_t)j * e.rs;
          for (int q = lane; q < segp; q += 32) lin_cp_async<T>(dst + q, src + (q < seg ? q : 0), q < seg);
        }
      }
    }
    lin_cp_commit();
  };
  issue(0);
  for (int step = 0; step < nsteps; ++step) {
    if (step + 1 < nsteps) {
      issue(step + 1);
      lin_cp_wait<1>();
    } else {
      lin_cp_wait<0>();
    }
    __syncthreads();
    const int t = step / nb_, b = step - t * nb_;
    const int n0 = t * e.tnode, tn = min(e.tnode, tn_all - n0);
    const int mi = p.mul_in[b], mo = p.mul_out[b], d = p.dim[b];
    const T* xs = xbuf + (size_t)(step & 1) * e.tnode * e.rs;
    const int* nodes = s_nodes + n0;
    if (mi == 0) {
      if (!p.accumulate) {
        const int span = mo * d;
        for (int j = warp; j < tn; j += 8)
          for (int q = lane; q < span; q += 32) OUT[(size_t)nodes[j] * p.out_dim + p.out_off[b] + q] = T(0);
      }
    } else {
      const T* ws = wsm + e.w_smem_off[b];
      const int mi4 = (mi + 3) & ~3;
      const T sc = T(p.scale[b]);
      // nodes per lane: 8 warps x NB covers the tile (NB = 4 for 32-node tiles, 2 for 16-node tiles), fewer for wide irreps
#define MT_LIN_BLOCK(DD, NBB) \
  lin_stream_block<T, DD, NBB>(ws, xs, e.rs, mi4, mo, tn, nodes, OUT, p.out_dim, p.out_off[b], sc, p.accumulate)
      if (e.tnode > 16) {
        switch (d) {
          case 1: MT_LIN_BLOCK(1, 4); break;
          case 3: MT_LIN_BLOCK(3, 4); break;
          case 5: MT_LIN_BLOCK(5, 2); break;
          case 7: MT_LIN_BLOCK(7, 2); break;
          default: MT_LIN_BLOCK(9, 1); break;
        }
      } else {
        switch (d) {
          case 1: MT_LIN_BLOCK(1, 2); break;
          case 3: MT_LIN_BLOCK(3, 2); break;
          case 5: MT_LIN_BLOCK(5, 2); break;
          case 7: MT_LIN_BLOCK(7, 1); break;
          default: MT_LIN_BLOCK(9, 1); break;
        }
      }
#undef MT_LIN_BLOCK
    }
    __syncthreads();  // the buffer is refilled by step + 2
  }
}

// Warp-private variant of the same arithmetic: a warp owns NB nodes at a time and walks ALL blocks over them with
// its own double-buffered rows and its own cp.async groups -- no CTA barrier after the weight slices are in place, so
// the eight warps drift apart and one warp's copies run under another's FMAs.  (The CTA-synchronous version spent its
// time at the two barriers per block: 5.5 k of 16 k stall samples on lin2, ncu r2_lin_full.)
constexpr int kWarpNB = 2;       // nodes per warp and step
constexpr int kWarpGroupsMax = 4;  // node groups per warp: a CTA covers 8 x kWarpNB x groups nodes of one species

template <typename T, int D, int NB>
__device__ __forceinline__ void lin_warp_block(const T* __restrict__ ws, const T* __restrict__ xs, int RS, int mi4,
                                               int mo, int ncount, const int* __restrict__ nodes,
                                               T* __restrict__ OUT, int out_dim, int out_off, T scale, int accumulate) {
  const int lane = threadIdx.x & 31;
  int lpn = 32;  // lanes per channel group: the smallest of 32/16/8/4 covering mul_out (32 when mul_out > 16)
  while (lpn > 4 && (lpn >> 1) >= mo) lpn >>= 1;
  const int subs = 32 / lpn, ks = lane / lpn, wl = lane - ks * lpn;
  const int ustep = 4 * subs;
  for (int w0 = 0; w0 < mo; w0 += lpn) {
    const int w = w0 + wl;
    const bool wok = w < mo;
    T acc[NB][D];
    const T* xq[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
      for (int m = 0; m < D; ++m) acc[nb][m] = T(0);
      xq[nb] = xs + (size_t)nb * RS + ks * 4 * D;  // rows past ncount hold stale finite data; never stored
    }
    const T* wq = ws + (wok ? w : 0) + ks * 4 * mo;
    for (int u = ks * 4; u < mi4; u += ustep) {
      T wv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) wv[k] = wq[k * mo];
      wq += ustep * mo;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        T xv[4 * D];
#pragma unroll
        for (int i = 0; i < D; ++i) lin_ld4<T>(xq[nb] + 4 * i, *reinterpret_cast<T(*)[4]>(&xv[4 * i]));
        xq[nb] += ustep * D;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int m = 0; m < D; ++m) acc[nb][m] = fma(wv[k], xv[k * D + m], acc[nb][m]);
      }
    }
    for (int off = lpn; off < 32; off <<= 1) {  // sum the k-splits (fixed order: deterministic)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int m = 0; m < D; ++m) acc[nb][m] += __shfl_xor_sync(0xffffffffu, acc[nb][m], off);
    }
    if (wok && ks == 0) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        if (nb < ncount) {
          T* dst = OUT + (size_t)nodes[nb] * out_dim + out_off + w * D;
#pragma unroll
          for (int m = 0; m < D; ++m) dst[m] = accumulate ? (dst[m] + acc[nb][m] * scale) : acc[nb][m] * scale;
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) linear_stream_kernel(const LinParams p, const LinStream e) {
  extern __shared__ __align__(16) unsigned char lin_smem[];
  __shared__ int s_nodes[8 * kWarpNB * kWarpGroupsMax];
  T* wsm = reinterpret_cast<T*>(lin_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int super = e.tnode;  // 8 warps x kWarpNB x groups
  // locate (species, tile-in-species)
  int tile = blockIdx.x, s = 0;
  int64_t begin = 0, end = 0;
  if (p.S == 1 && p.sptr == nullptr) {
    begin = (int64_t)tile * super;
    end = imin64(begin + super, p.N);
    if (begin >= p.N) return;
  } else {
    bool found = false;
    for (s = 0; s < p.S; ++s) {
      const int cnt = p.sptr[s + 1] - p.sptr[s];
      const int nt = (cnt + super - 1) / super;
      if (tile < nt) {
        begin = p.sptr[s] + (int64_t)tile * super;
        end = imin64(begin + super, (int64_t)p.sptr[s + 1]);
        found = true;
        break;
      }
      tile -= nt;
    }
    if (!found) return;
  }
  const int tn_all = (int)(end - begin);
  for (int i = tid; i < tn_all; i += 256) s_nodes[i] = p.sperm ? p.sperm[begin + i] : (int)(begin + i);
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ W = static_cast<const T*>(p.weight);
  T* __restrict__ OUT = static_cast<T*>(p.out);
  constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte copy
  // this species' weight slices, [mi4][mo] per block (rows mi..mi4 zero); asynchronous copies, all in flight at once.
  // Rows of mo contiguous elements: 16 bytes per copy when mo allows.
  for (int b = 0; b < p.num_blocks; ++b) {
    const int mi = p.mul_in[b], mo = p.mul_out[b], mi4 = (mi + 3) & ~3;
    T* dst = wsm + e.w_smem_off[b];
    if (!p.transpose && e.w_vec[b]) {
      const int rowv = mo / V;  // vectors per row
      const int nvec = mi4 * rowv;
      int u = tid / rowv, c = tid - u * rowv;          // one division per thread and block; then additive
      const int du = 256 / rowv, dc = 256 - du * rowv;
      for (int v = tid; v < nvec; v += 256) {
        const bool ok = u < mi;
        const T* src = W + (ok ? (size_t)p.w_off[b] + ((size_t)u * p.S + s) * mo + c * V : 0);
        const uint32_t da = (uint32_t)__cvta_generic_to_shared(dst + u * mo + c * V);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(src), "r"(ok ? 16 : 0) : "memory");
        u += du; c += dc;
        if (c >= rowv) { c -= rowv; ++u; }
      }
    } else {
      for (int u = warp; u < mi4; u += 8)
        for (int w = lane; w < mo; w += 32) {
          const bool ok = u < mi;
          const size_t off = !ok ? 0
                             : p.transpose ? (size_t)p.w_off[b] + ((size_t)w * p.S + s) * mi + u
                                           : (size_t)p.w_off[b] + ((size_t)u * p.S + s) * mo + w;
          lin_cp_async<T>(dst + u * mo + w, W + off, ok);
        }
    }
  }
  lin_cp_commit();
  lin_cp_wait<0>();
  __syncthreads();  // weights and s_nodes in place: the only CTA-wide barrier

  // ---- from here on every warp runs on its own
  const int nb_ = p.num_blocks;
  const int ngroups = (tn_all + 8 * kWarpNB - 1) / (8 * kWarpNB);
  // my groups: node offsets (g * 8 + warp) * kWarpNB
  int mygroups = 0;
  for (int g = 0; g < ngroups; ++g)
    if ((g * 8 + warp) * kWarpNB < tn_all) ++mygroups;
  if (mygroups == 0) return;
  const int nsteps = mygroups * nb_;
  T* xw = wsm + ((e.w_total + 3) & ~3) + (size_t)warp * 2 * kWarpNB * e.rs;  // [2][kWarpNB][rs]
  const uint32_t xw_s = (uint32_t)__cvta_generic_to_shared(xw);
  auto issue = [&](int step) {
    const int g = step / nb_, b = step - g * nb_;
    const int mi = p.mul_in[b], d = p.dim[b];
    const int seg = mi * d, segp = (((mi + 3) & ~3) * d + 3) & ~3;  // copied + zero-filled up to the padded length
    const int n0 = (g * 8 + warp) * kWarpNB;
    const uint32_t buf_s = xw_s + (uint32_t)((step & 1) * kWarpNB * e.rs * (int)sizeof(T));
    if (mi > 0) {
#pragma unroll
      for (int nb = 0; nb < kWarpNB; ++nb) {
        if (n0 + nb >= tn_all) break;
        const T* src = X + (size_t)s_nodes[n0 + nb] * p.in_dim + p.in_off[b];
        const uint32_t drow = buf_s + (uint32_t)(nb * e.rs * (int)sizeof(T));
        if (e.vec_ok[b]) {
          for (int q = lane * V; q < segp; q += 32 * V) {
            const int n = min(V, seg - q);  // elements really there
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + q * (int)sizeof(T)),
                         "l"(src + (n > 0 ? q : 0)), "r"(n > 0 ? n * (int)sizeof(T) : 0)
                         : "memory");
          }
        } else {
          T* dst = xw + (size_t)(step & 1) * kWarpNB * e.rs + (size_t)nb * e.rs;
          for (int q = lane; q < segp; q += 32) lin_cp_async<T>(dst + q, src + (q < seg ? q : 0), q < seg);
        }
      }
    }
    lin_cp_commit();
  };
  // rows of a partial last group are never copied: make them finite once (they are multiplied, never stored)
  for (int i = lane; i < 2 * kWarpNB * e.rs; i += 32) xw[i] = T(0);
  __syncwarp();
  issue(0);
  for (int step = 0; step < nsteps; ++step) {
    if (step + 1 < nsteps) {
      issue(step + 1);
      lin_cp_wait<1>();
    } else {
      lin_cp_wait<0>();
    }
    __syncwarp();  // every lane's copies of this step have landed
    const int g = step / nb_, b = step - g * nb_;
    const int n0 = (g * 8 + warp) * kWarpNB;
    const int ncount = min(kWarpNB, tn_all - n0);
    const int mi = p.mul_in[b], mo = p.mul_out[b], d = p.dim[b];
    const T* xs = xw + (size_t)(step & 1) * kWarpNB * e.rs;
    const int* nodes = s_nodes + n0;
    if (mi == 0) {
      if (!p.accumulate) {
        const int span = mo * d;
        for (int j = 0; j < ncount; ++j)
          for (int q = lane; q < span; q += 32) OUT[(size_t)nodes[j] * p.out_dim + p.out_off[b] + q] = T(0);
      }
    } else {
      const T* ws = wsm + e.w_smem_off[b];
      const int mi4 = (mi + 3) & ~3;
      const T sc = T(p.scale[b]);
#define MT_LIN_BLOCK(DD) \
  lin_warp_block<T, DD, kWarpNB>(ws, xs, e.rs, mi4, mo, ncount, nodes, OUT, p.out_dim, p.out_off[b], sc, p.accumulate)
      switch (d) {
        case 1: MT_LIN_BLOCK(1); break;
        case 3: MT_LIN_BLOCK(3); break;
        case 5: MT_LIN_BLOCK(5); break;
        case 7: MT_LIN_BLOCK(7); break;
        default: MT_LIN_BLOCK(9); break;
      }
#undef MT_LIN_BLOCK
    }
    __syncwarp();  // the buffer is refilled by step + 2
  }
}

// =========================================================================
// Gate (+ folded BatchNorm affine)      reference src/matten/nn/utils.py:134-140,418
// =========================================================================
template <typename T>
__global__ void __launch_bounds__(256) gate_fwd_kernel(const T* __restrict__ x, int in_dim, int out_dim,
                                                       const int32_t* __restrict__ src_idx,
                                                       const int32_t* __restrict__ gate_idx,
                                                       const int32_t* __restrict__ act_id,
                                                       const T* __restrict__ act_cst,
                                                       const T* __restrict__ aff_a,
                                                       const T* __restrict__ aff_b, T* __restrict__ out,
                                                       int64_t N) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * out_dim) return;
  int64_t n = t / out_dim;
  int j = (int)(t - n * out_dim);
  const T* xr = x + n * in_dim;
  T v = xr[src_idx[j]];
  int g = gate_idx[j];
  T y;
  if (g < 0) y = apply_act<T>(act_id[j], v) * act_cst[j];
  else y = v * (apply_act<T>(act_id[j], xr[g]) * act_cst[j]);
  if (aff_a) y = y * aff_a[j] + aff_b[j];
  out[t] = y;
}

// =========================================================================
// segmented pooling                     reference src/matten/nn/nodewise.py:142-148
// =========================================================================
template <typename T>
__global__ void __launch_bounds__(256) segment_reduce_kernel(const T* __restrict__ x,
                                                             const int32_t* __restrict__ ptr, int dim,
                                                             int64_t B, int mode, T* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= B * dim) return;
  int64_t b = t / dim;
  int j = (int)(t - b * dim);
  int n0 = ptr[b], n1 = ptr[b + 1];
  T acc = T(0);
  if (n1 > n0) {
    acc = x[(size_t)n0 * dim + j];
    for (int n = n0 + 1; n < n1; ++n) {
      T v = x[(size_t)n * dim + j];
      if (mode <= 1) acc += v;
      else if (mode == 2) acc = v < acc ? v : acc;
      else acc = v > acc ? v : acc;
    }
    if (mode == 1) acc = acc / T(n1 - n0);
  }
  out[t] = acc;
}

}  // namespace mt

using namespace mt;

extern "C" {

static int linear_fwd_round(int dtype, const mt_lin_block* const* blocks, int num_blocks, int in_dim,
                            int out_dim, int num_species, const void* x, const void* weight,
                            const int32_t* species_perm, const int32_t* species_ptr, int accumulate, void* out,
                            int64_t N, cudaStream_t st, int transpose = 0) {
  LinParams p;
  memset(&p, 0, sizeof(p));
  p.transpose = transpose;
  p.num_blocks = num_blocks;
  int64_t total = 0;
  // ---- node-streaming kernel whenever the block dims are 2l+1 <= 9 and the tile fits in shared memory
  {
    LinStream e;
    memset(&e, 0, sizeof(e));
    const size_t es = dtype == MT_F64 ? 8 : 4;
    const int V = 16 / (int)es;
    bool ok = true;
    int rs = 0;
    for (int b = 0; b < num_blocks; ++b) {
      const mt_lin_block& k = *blocks[b];
      p.in_off[b] = k.in_off; p.out_off[b] = k.out_off; p.mul_in[b] = k.mul_in; p.mul_out[b] = k.mul_out;
      p.dim[b] = k.dim; p.w_off[b] = k.w_off; p.scale[b] = k.scale;
      if (!(k.dim == 1 || k.dim == 3 || k.dim == 5 || k.dim == 7 || k.dim == 9)) ok = false;
      const int mi4 = (k.mul_in + 3) & ~3;
      e.w_smem_off[b] = e.w_total;
      e.w_total += mi4 * k.mul_out;
      e.w_total = (e.w_total + 3) & ~3;
      const int segp = (mi4 * k.dim + 3) & ~3;
      if (segp > rs) rs = segp;
      e.vec_ok[b] = (k.in_off % V == 0) && (in_dim % V == 0) && ((uintptr_t)x % 16 == 0);
      e.w_vec[b] = (k.mul_out % V == 0) && (k.w_off % V == 0) && ((uintptr_t)weight % 16 == 0) && k.mul_out / V <= 256;
    }
    e.rs = rs + 4;
    if (in_dim < 512) {  // short rows: CTA-synchronous tiles
    int tnode = 32;
      auto need = [&](int tn) { return ((size_t)((e.w_total + 3) & ~3) + 2 * (size_t)tn * e.rs) * es; };
      // two resident CTAs per SM (16 warps, one CTA's copies under the other's arithmetic) beat one big tile
      while (tnode > 8 && need(tnode) > (size_t)(kStreamSmemBytes / 2 - 8 * 1024)) tnode >>= 1;
      if (ok && need(tnode) <= (size_t)kStreamSmemBytes) {
        e.tnode = tnode;
        p.in_dim = in_dim; p.out_dim = out_dim; p.S = num_species;
        p.x = x; p.weight = weight; p.sperm = species_perm; p.sptr = species_ptr;
        p.accumulate = accumulate; p.out = out; p.N = N;
        // several tiles per CTA only when that still leaves >= 4 CTAs per SM
        int tpc = (int)(N / ((int64_t)tnode * 4 * kNumSMs));
        tpc = tpc < 1 ? 1 : (tpc > kStreamTiles ? kStreamTiles : tpc);
        e.tiles_per_cta = tpc;
        const int64_t tiles = ceil_div<int64_t>(N, (int64_t)tnode * tpc) + (species_ptr ? num_species : 0);
        MT_REQUIRE(tiles < (int64_t)2147483647, "grid too large");
        const size_t smem = need(tnode);
        MT_DISPATCH_DTYPE(dtype, {
          static thread_local bool configured = false;
          if (!configured) {
            MT_CUDA_OK(cudaFuncSetAttribute(linear_sync_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kStreamSmemBytes));
            configured = true;
          }
          linear_sync_kernel<T><<<(unsigned)tiles, 256, smem, st>>>(p, e);
        });
        MT_LAUNCH_OK();
        return MT_OK;
      }
  
    }
    // a CTA covers 8 warps x kWarpNB nodes x groups of one species; more groups amortise the weight staging, fewer
    // keep the grid above ~4 CTAs per SM
    int groups = kWarpGroupsMax;
    while (groups > 1 && N / (8 * kWarpNB * groups) < 4 * kNumSMs) groups >>= 1;
    const int tnode = 8 * kWarpNB * groups;
    auto need = [&](int) { return ((size_t)((e.w_total + 3) & ~3) + 8 * 2 * (size_t)kWarpNB * e.rs) * es; };
    if (ok && need(tnode) <= (size_t)kStreamSmemBytes) {
      e.tnode = tnode;
      p.in_dim = in_dim; p.out_dim = out_dim; p.S = num_species;
      p.x = x; p.weight = weight; p.sperm = species_perm; p.sptr = species_ptr;
      p.accumulate = accumulate; p.out = out; p.N = N;
      e.tiles_per_cta = 1;
      const int64_t tiles = ceil_div<int64_t>(N, (int64_t)tnode) + (species_ptr ? num_species : 0);
      MT_REQUIRE(tiles < (int64_t)2147483647, "grid too large");
      const size_t smem = need(tnode);
      MT_DISPATCH_DTYPE(dtype, {
        static thread_local bool configured = false;
        if (!configured) {
          MT_CUDA_OK(cudaFuncSetAttribute(linear_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kStreamSmemBytes));
          configured = true;
        }
        linear_stream_kernel<T><<<(unsigned)tiles, 256, smem, st>>>(p, e);
      });
      MT_LAUNCH_OK();
      return MT_OK;
    }
  }
  for (int b = 0; b < num_blocks; ++b) {
    const mt_lin_block& k = *blocks[b];
    p.in_off[b] = k.in_off; p.out_off[b] = k.out_off; p.mul_in[b] = k.mul_in; p.mul_out[b] = k.mul_out;
    p.dim[b] = k.dim; p.w_off[b] = k.w_off; p.scale[b] = k.scale;
    int cols = 64;
    while (cols > 4 && cols / 2 >= k.mul_out) cols /= 2;  // smallest of 64/32/16/8/4 that covers mul_out (<= 64)
    const int rows = (256 / (cols / 4)) * (cols >= 16 ? 8 : 4);
    int tn = rows / k.dim;
    if (tn < 1) tn = 1;
    p.tile_cols[b] = cols;
    p.nodes_per_tile[b] = tn;
    p.col_chunks[b] = ceil_div<int>(k.mul_out, cols);
    int64_t tiles = ceil_div<int64_t>(N, tn) + (species_ptr ? num_species : 0);
    p.cta_begin[b] = (int32_t)total;
    total += tiles * p.col_chunks[b];
    MT_REQUIRE(total < (int64_t)2147483647, "grid too large");
  }
  p.cta_begin[num_blocks] = (int32_t)total;
  p.in_dim = in_dim; p.out_dim = out_dim; p.S = num_species;
  p.x = x; p.weight = weight; p.sperm = species_perm; p.sptr = species_ptr;
  p.accumulate = accumulate; p.out = out; p.N = N;
  MT_DISPATCH_DTYPE(dtype, {
    const size_t smem = (size_t)kLinSmemElems * sizeof(T);
    static thread_local bool configured = false;
    if (!configured) {
      MT_CUDA_OK(cudaFuncSetAttribute(linear_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    linear_fwd_kernel<T><<<(unsigned)total, 256, smem, st>>>(p);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

static int linear_check_blocks(const mt_lin_block* blocks, int num_blocks, int in_dim, int out_dim,
                               const void* weight) {
  for (int b = 0; b < num_blocks; ++b) {
    const mt_lin_block& k = blocks[b];
    MT_REQUIRE(k.dim >= 1 && k.dim <= kLinMaxDim && k.mul_out > 0 && k.mul_in >= 0, "bad linear block %d", b);
    MT_REQUIRE(k.out_off >= 0 && k.out_off + k.mul_out * k.dim <= out_dim, "block %d exceeds out_dim", b);
    MT_REQUIRE(k.mul_in == 0 || (k.in_off >= 0 && k.in_off + k.mul_in * k.dim <= in_dim), "block %d exceeds in_dim", b);
    MT_REQUIRE(k.mul_in == 0 || weight != nullptr, "null weight");
  }
  return MT_OK;
}

// Blocks that write the same output range (several input irreps of one type feeding one output,
// e3nn sums them) must not race: round r holds the r-th block of every distinct output range;
// rounds after the first accumulate.  Simplified irreps (every matten model) need one round.
static int linear_rounds(int dtype, const mt_lin_block* blocks, int num_blocks, int in_dim, int out_dim,
                         int num_species, const void* x, const void* weight, const int32_t* species_perm,
                         const int32_t* species_ptr, int accumulate, void* out, int64_t N, cudaStream_t st,
                         int transpose) {
  int round_of[4 * kLinMaxBlocks];
  int max_round = 0;
  for (int b = 0; b < num_blocks; ++b) {
    int r = 0;
    for (int c = 0; c < b; ++c)
      if (blocks[c].out_off == blocks[b].out_off) ++r;
    round_of[b] = r;
    if (r > max_round) max_round = r;
  }
  for (int r = 0; r <= max_round; ++r) {
    const mt_lin_block* sel[kLinMaxBlocks];
    int n = 0;
    for (int b = 0; b < num_blocks; ++b) {
      if (round_of[b] != r) continue;
      MT_REQUIRE(n < kLinMaxBlocks, "more than %d linear blocks in one round", kLinMaxBlocks);
      sel[n++] = &blocks[b];
    }
    int rc = linear_fwd_round(dtype, sel, n, in_dim, out_dim, num_species, x, weight, species_perm, species_ptr,
                              (accumulate || r > 0) ? 1 : 0, out, N, st, transpose);
    if (rc != MT_OK) return rc;
  }
  return MT_OK;
}

// Backward of the linear w.r.t. x: the transposed weight slices applied to grad_out (train_ops.cu).
int linear_transposed(int dtype, const mt_lin_block* blocks, int num_blocks, int in_dim, int out_dim,
                      int num_species, const void* grad_out, const void* weight, const int32_t* species_perm,
                      const int32_t* species_ptr, int accumulate, void* grad_x, int64_t N, cudaStream_t st) {
  mt_lin_block tb[4 * kLinMaxBlocks];
  int n = 0;
  for (int b = 0; b < num_blocks; ++b) {
    const mt_lin_block& k = blocks[b];
    if (k.mul_in == 0) continue;  // zero-fill blocks carry no gradient
    tb[n] = k;
    tb[n].in_off = k.out_off; tb[n].out_off = k.in_off; tb[n].mul_in = k.mul_out; tb[n].mul_out = k.mul_in;
    ++n;
  }
  if (!accumulate) {
    const size_t es = dtype == MT_F64 ? 8 : 4;
    MT_CUDA_OK(cudaMemsetAsync(grad_x, 0, (size_t)N * in_dim * es, st));
  }
  if (n == 0) return MT_OK;
  return linear_rounds(dtype, tb, n, out_dim, in_dim, num_species, grad_out, weight, species_perm, species_ptr, 1,
                       grad_x, N, st, 1);
}

int mt_linear_fwd(int dtype, const mt_lin_block* blocks, int num_blocks, int in_dim, int out_dim,
                  int num_species, const void* x, const void* weight, const int32_t* species_perm,
                  const int32_t* species_ptr, int accumulate, void* out, int64_t N, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(blocks && num_blocks > 0 && num_blocks <= 4 * kLinMaxBlocks, "num_blocks %d not in 1..%d",
             num_blocks, 4 * kLinMaxBlocks);
  MT_REQUIRE(in_dim > 0 && out_dim > 0 && num_species >= 1, "bad dims");
  MT_REQUIRE((species_perm == nullptr) == (species_ptr == nullptr), "species_perm/ptr must be given together");
  MT_REQUIRE(num_species == 1 || species_ptr != nullptr, "species grouping required when num_species > 1");
  if (N == 0) return MT_OK;
  MT_REQUIRE(x && out, "null pointer");
  int rc = linear_check_blocks(blocks, num_blocks, in_dim, out_dim, weight);
  if (rc != MT_OK) return rc;
  return linear_rounds(dtype, blocks, num_blocks, in_dim, out_dim, num_species, x, weight, species_perm,
                       species_ptr, accumulate, out, N, as_stream(stream), 0);
}

int mt_gate_fwd(int dtype, const void* x, int in_dim, int out_dim, const int32_t* src_idx,
                const int32_t* gate_idx, const int32_t* act_id, const void* act_cst, const void* affine_a,
                const void* affine_b, void* out, int64_t N, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(in_dim > 0 && out_dim > 0, "bad dims");
  MT_REQUIRE((affine_a == nullptr) == (affine_b == nullptr), "affine_a/b must be given together");
  if (N == 0) return MT_OK;
  MT_REQUIRE(x && out && src_idx && gate_idx && act_id && act_cst, "null pointer");
  int64_t total = N * out_dim;
  MT_DISPATCH_DTYPE(dtype, {
    gate_fwd_kernel<T><<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        (const T*)x, in_dim, out_dim, src_idx, gate_idx, act_id, (const T*)act_cst, (const T*)affine_a,
        (const T*)affine_b, (T*)out, N);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

int mt_segment_reduce(int dtype, const void* x, const int32_t* ptr, int dim, int64_t B, int mode, void* out,
                      mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0 && mode >= 0 && mode <= 3, "bad arguments");
  if (B == 0) return MT_OK;
  MT_REQUIRE(x && ptr && out, "null pointer");
  int64_t total = B * dim;
  MT_DISPATCH_DTYPE(dtype, {
    segment_reduce_kernel<T><<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        (const T*)x, ptr, dim, B, mode, (T*)out);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

}  // extern "C"
