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
constexpr int kLinOut = 4096;    // outputs per CTA tile: rows x cols, 256 threads x (4 x 4)
constexpr int kLinStage = 8192;  // staged x elements per k-chunk (rows x KC)
constexpr int kLinMaxRows = 1024;
// shared memory (elements of T): max over the tile shapes of staging (KC x (rows + 4) + KC x cols) and of the
// output tile (rows x (cols + 1)) that aliases it
constexpr int kLinSmemElems = 8832;

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

struct LinParams {
  int num_blocks;
  int32_t in_off[kLinMaxBlocks], out_off[kLinMaxBlocks], mul_in[kLinMaxBlocks], mul_out[kLinMaxBlocks],
      dim[kLinMaxBlocks], w_off[kLinMaxBlocks];
  double scale[kLinMaxBlocks];
  int32_t cta_begin[kLinMaxBlocks + 1];  // prefix of CTA counts per block
  int32_t col_chunks[kLinMaxBlocks];     // ceil(mul_out / cols)
  int32_t nodes_per_tile[kLinMaxBlocks];
  int32_t tile_cols[kLinMaxBlocks];      // 64 / 32 / 16 / 8 / 4: the tile is (4096 / cols) rows x cols
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

// Tile shape per block: narrow outputs (mul_out = 4 for the l = 2 irreps, 16 for l = 1) take tall tiles, so the
// 4 x 4 register micro-kernel never multiplies padding (a fixed 64 x 64 tile spent 97 % of its instructions on
// it for the l >= 1 blocks of lin2: ncu r1, 330 M instructions for 18 M useful FMAs).
template <typename T>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const LinParams p) {
  extern __shared__ __align__(16) unsigned char lin_smem[];
  __shared__ int s_nodes[kLinMaxRows];

  // which block / node tile / column chunk am I?
  int b = 0;
  while (b + 1 < p.num_blocks && (int)blockIdx.x >= p.cta_begin[b + 1]) ++b;
  int local = blockIdx.x - p.cta_begin[b];
  const int cc = local % p.col_chunks[b];
  int tile = local / p.col_chunks[b];
  const int TN = p.nodes_per_tile[b];
  const int mi = p.mul_in[b], mo = p.mul_out[b], d = p.dim[b];
  const int COLS = p.tile_cols[b], ROWS = kLinOut / COLS;
  const int KC = min(32, kLinStage / ROWS);
  const int XS = ROWS + 4;  // row stride of the transposed x stage
  T* xsT = reinterpret_cast<T*>(lin_smem);              // [KC][XS]
  T* ws = xsT + (size_t)KC * XS;                        // [KC][COLS]
  T* ot = reinterpret_cast<T*>(lin_smem);               // [ROWS][COLS + 1], aliases the stage after the k loop
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
  const int tn = (int)(end - begin);
  const int tid = threadIdx.x;
  for (int t = tid; t < tn; t += blockDim.x) s_nodes[t] = p.sperm ? p.sperm[begin + t] : (int)(begin + t);
  __syncthreads();

  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ W = static_cast<const T*>(p.weight);
  T* __restrict__ OUT = static_cast<T*>(p.out);
  const int c0 = cc * COLS;
  const int ncols = min(COLS, mo - c0);

  if (mi == 0) {  // irreps with no incoming path: zeros
    if (!p.accumulate) {
      for (int t = tid; t < tn * ncols * d; t += blockDim.x) {
        int j = t / (ncols * d), q = t - j * (ncols * d);
        OUT[(size_t)s_nodes[j] * p.out_dim + p.out_off[b] + c0 * d + q] = T(0);
      }
    }
    return;
  }

  const int tcn = COLS >> 2;  // thread columns
  const int tc = tid % tcn, tr = tid / tcn;
  using P = typename pair_of<T>::type;  // column pairs: one FFMA2 per pair in fp32
  P acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[i][j].x = T(0); acc[i][j].y = T(0); }

  for (int u0 = 0; u0 < mi; u0 += KC) {
    const int ku = min(KC, mi - u0);
    // stage weights  ws[uu][c] = W[w_off + ((u0+uu)*S + s)*mo + c0 + c]
    for (int t = tid; t < KC * COLS; t += blockDim.x) {
      int uu = t / COLS, c = t - uu * COLS;
      T v = T(0);
      if (uu < ku && c < ncols)
        v = p.transpose ? W[(size_t)p.w_off[b] + ((size_t)(c0 + c) * p.S + s) * mi + (u0 + uu)]
                        : W[(size_t)p.w_off[b] + ((size_t)(u0 + uu) * p.S + s) * mo + c0 + c];
      ws[t] = v;
    }
    // stage x transposed  xsT[uu][j*d+m] = X[node_j, in_off + (u0+uu)*d + m]; rows beyond tn*d and uu >= ku are zero
    const int seg = ku * d;
    const int R = tn * d;
    if (ku < KC || R < ROWS) {  // partial chunk / tile only: the staging below overwrites everything else
      for (int t = tid; t < KC * ROWS; t += blockDim.x) {
        const int uu = t / ROWS, r = t - uu * ROWS;
        if (uu >= ku || r >= R) xsT[uu * XS + r] = T(0);
      }
    }
    for (int t = tid; t < tn * seg; t += blockDim.x) {
      int j = t / seg, q = t - j * seg;
      int uu = q / d, m = q - uu * d;
      xsT[uu * XS + j * d + m] = X[(size_t)s_nodes[j] * p.in_dim + p.in_off[b] + u0 * d + q];
    }
    __syncthreads();
#pragma unroll 4
    for (int uu = 0; uu < KC; ++uu) {
      T a[4], bb[4];
      lin_ld4<T>(xsT + uu * XS + tr * 4, a);   // 16-byte aligned: XS and COLS are multiples of 4
      lin_ld4<T>(ws + uu * COLS + tc * 4, bb);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        fma_pair(a[i], bb[0], bb[1], acc[i][0]);
        fma_pair(a[i], bb[2], bb[3], acc[i][1]);
      }
    }
    __syncthreads();
  }
  const T scale = T(p.scale[b]);
  const int OS = COLS + 1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ot[(tr * 4 + i) * OS + tc * 4 + 2 * j] = acc[i][j].x * scale;
      ot[(tr * 4 + i) * OS + tc * 4 + 2 * j + 1] = acc[i][j].y * scale;
    }
  __syncthreads();
  // coalesced write: per node the (w, m) range is contiguous
  const int span = ncols * d;
  for (int t = tid; t < tn * span; t += blockDim.x) {
    int j = t / span, q = t - j * span;
    int w = q / d, m = q - w * d;
    size_t o = (size_t)s_nodes[j] * p.out_dim + p.out_off[b] + c0 * d + q;
    T v = ot[(j * d + m) * OS + w];
    OUT[o] = p.accumulate ? (OUT[o] + v) : v;
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
  for (int b = 0; b < num_blocks; ++b) {
    const mt_lin_block& k = *blocks[b];
    p.in_off[b] = k.in_off; p.out_off[b] = k.out_off; p.mul_in[b] = k.mul_in; p.mul_out[b] = k.mul_out;
    p.dim[b] = k.dim; p.w_off[b] = k.w_off; p.scale[b] = k.scale;
    int cols = 64;
    while (cols > 4 && cols / 2 >= k.mul_out) cols /= 2;  // smallest of 64/32/16/8/4 that covers mul_out (<= 64)
    const int rows = kLinOut / cols;
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
    MT_REQUIRE(k.dim >= 1 && k.dim <= 64 && k.mul_out > 0 && k.mul_in >= 0, "bad linear block %d", b);
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
