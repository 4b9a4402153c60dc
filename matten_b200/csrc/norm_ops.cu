// Normalisation rows of SURVEY section 8(f4): the graph-wise InstanceNorm, e3nn's NormActivation, the target
// normalisers of the data pipeline, and the arg-based backward of min/max pooling.
// See include/matten_b200.h for the contract of every entry point.  Deterministic: every reduction is a serial
// fixed-order loop owned by one thread.
#include "common.cuh"

namespace mt {

// ------------------------------------------------------------------------------------------------------------
// Graph InstanceNorm                      reference src/matten/nn/utils.py:448-588
// thread <-> (graph, channel): the channel's 2l+1 columns of the graph's nodes are walked three times (mean,
// squared norm, apply).  A crystal has tens of nodes, so the walks hit L1/L2; x is read from HBM once.
// ------------------------------------------------------------------------------------------------------------
// The statistics and the per-element map are evaluated in double whatever T is: centring subtracts nearly equal
// numbers (a two-node crystal with similar scalars), where an fp32 mean costs 1e-5 of the normalised value.
template <typename T>
__global__ void __launch_bounds__(128) instance_norm_fwd_kernel(
    const T* __restrict__ x, const int32_t* __restrict__ gptr, int dim, int nf, const int32_t* __restrict__ first,
    const int32_t* __restrict__ cdim, const int32_t* __restrict__ scal, const T* __restrict__ weight,
    const T* __restrict__ bias, double eps, int reduce, int normalization, T* __restrict__ out,
    double* __restrict__ save_mean, double* __restrict__ save_rstd, int32_t* __restrict__ save_arg) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int g = blockIdx.x;
  if (c >= nf) return;
  const int n0 = gptr[g], n1 = gptr[g + 1];
  const int d = cdim[c], f = first[c], s = scal[c];
  const double cnt = double(n1 - n0 > 0 ? n1 - n0 : 1);
  double mean = 0.0;
  if (s >= 0) {
    for (int n = n0; n < n1; ++n) mean += double(x[(size_t)n * dim + f]);
    mean = mean / cnt;
  }
  const double dd = normalization == 0 ? double(d) : 1.0;
  double v = 0.0;
  int arg = n0;
  for (int n = n0; n < n1; ++n) {
    double ss = 0.0;
    for (int m = 0; m < d; ++m) {
      double xc = double(x[(size_t)n * dim + f + m]) - mean;
      ss += xc * xc;
    }
    ss = ss / dd;
    if (reduce == 0) v += ss;
    else if (n == n0 || ss > v) { v = ss; arg = n; }
  }
  if (reduce == 0) v = v / cnt;
  const double rstd = 1.0 / sqrt(v + eps);
  const double a = weight ? rstd * double(weight[c]) : rstd;
  const double b = (bias && s >= 0) ? double(bias[s]) : 0.0;
  for (int n = n0; n < n1; ++n)
    for (int m = 0; m < d; ++m) {
      size_t o = (size_t)n * dim + f + m;
      out[o] = T((double(x[o]) - mean) * a + b);
    }
  const size_t k = (size_t)g * nf + c;
  save_mean[k] = mean;
  save_rstd[k] = rstd;
  save_arg[k] = arg;
}

// With xc = x - mean, S = sum_{n,m} g xc:   d/dxc = a g - [mean: 1/cnt | max: (n == arg)] a rstd^2 xc S / dd,
// and l = 0 channels subtract the node mean of that (centring).  Per-graph partial parameter gradients are written
// to [G, nf] buffers and summed over graphs by mt_col_reduce on the host side (fixed order).
template <typename T>
__global__ void __launch_bounds__(128) instance_norm_bwd_kernel(
    const T* __restrict__ x, const T* __restrict__ go, const int32_t* __restrict__ gptr, int dim, int nf,
    const int32_t* __restrict__ first, const int32_t* __restrict__ cdim, const int32_t* __restrict__ scal,
    const T* __restrict__ weight, int reduce, int normalization, const double* __restrict__ save_mean,
    const double* __restrict__ save_rstd, const int32_t* __restrict__ save_arg, T* __restrict__ gx,
    T* __restrict__ gw_part, T* __restrict__ gb_part) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int g = blockIdx.x;
  if (c >= nf) return;
  const int n0 = gptr[g], n1 = gptr[g + 1];
  const int d = cdim[c], f = first[c], s = scal[c];
  const size_t k = (size_t)g * nf + c;
  const double mean = save_mean[k], rstd = save_rstd[k];
  const int arg = save_arg[k];
  const double cnt = double(n1 - n0 > 0 ? n1 - n0 : 1);
  const double dd = normalization == 0 ? double(d) : 1.0;
  const double a = weight ? rstd * double(weight[c]) : rstd;
  double S = 0.0, sg = 0.0;
  for (int n = n0; n < n1; ++n)
    for (int m = 0; m < d; ++m) {
      size_t o = (size_t)n * dim + f + m;
      double gv = double(go[o]);
      S += gv * (double(x[o]) - mean);
      sg += gv;
    }
  const double coef = a * rstd * rstd * S / dd;  // multiplies xc on the selected rows
  double shift = 0.0;                            // node mean of d/dxc for centred channels
  if (s >= 0 && n1 > n0) {
    shift = a * sg / cnt;
    if (reduce != 0) shift -= coef * (double(x[(size_t)arg * dim + f]) - mean) / cnt;
    // reduce == mean: sum_n xc = 0, the second term vanishes
  }
  for (int n = n0; n < n1; ++n) {
    const double sel = reduce == 0 ? 1.0 / cnt : (n == arg ? 1.0 : 0.0);
    for (int m = 0; m < d; ++m) {
      size_t o = (size_t)n * dim + f + m;
      gx[o] = T(a * double(go[o]) - sel * coef * (double(x[o]) - mean) - shift);
    }
  }
  if (gw_part) gw_part[k] = T(S * rstd);
  if (gb_part) gb_part[k] = T(s >= 0 ? sg : 0.0);  // [G, nf]; the host keeps the l = 0 channels
}

// ------------------------------------------------------------------------------------------------------------
// NormActivation (normalize=True, epsilon=1e-8, bias=False)      reference src/matten/nn/utils.py:142-150
//   n = sqrt(max(sum_m x_m^2, eps^2));  y_m = x_m f(n) / n          thread <-> (node, channel)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const T* __restrict__ x, int dim, int nf,
                                                           const int32_t* __restrict__ first,
                                                           const int32_t* __restrict__ cdim, int act, T eps2,
                                                           T* __restrict__ out, int64_t N) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * nf) return;
  const int64_t n = t / nf;
  const int c = (int)(t - n * nf);
  const int d = cdim[c];
  const T* xr = x + n * dim + first[c];
  T* yr = out + n * dim + first[c];
  T ss = T(0);
  for (int m = 0; m < d; ++m) ss += xr[m] * xr[m];
  if (ss < eps2) ss = eps2;
  const T nn = sqrt(ss);
  const T sc = apply_act<T>(act, nn) / nn;
  for (int m = 0; m < d; ++m) yr[m] = sc * xr[m];
}

template <typename T>
__global__ void __launch_bounds__(256) norm_act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ go, int dim,
                                                           int nf, const int32_t* __restrict__ first,
                                                           const int32_t* __restrict__ cdim, int act, T eps2,
                                                           T* __restrict__ gx, int64_t N) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * nf) return;
  const int64_t n = t / nf;
  const int c = (int)(t - n * nf);
  const int d = cdim[c];
  const size_t o = (size_t)n * dim + first[c];
  T ss = T(0), gxs = T(0);
  for (int m = 0; m < d; ++m) {
    ss += x[o + m] * x[o + m];
    gxs += go[o + m] * x[o + m];
  }
  const bool clamped = ss < eps2;  // the clamped norm is a constant: no gradient through it
  if (clamped) ss = eps2;
  const T nn = sqrt(ss);
  const T fv = apply_act<T>(act, nn);
  const T sc = fv / nn;
  // d(f(n)/n)/dx_m = (f'(n) n - f(n)) / n^2 * x_m / n
  const T k = clamped ? T(0) : (apply_act_grad<T>(act, nn) * nn - fv) / (ss * nn) * gxs;
  for (int m = 0; m < d; ++m) gx[o + m] = sc * go[o + m] + k * x[o + m];
}

// ------------------------------------------------------------------------------------------------------------
// Target normalisers                      reference src/matten/data/transform.py:116-133, 265-279
//   forward: (data - mean) / (norm * scale)      inverse: data * (norm * scale) + mean   (separate roundings, as
//   the reference's elementwise expressions evaluate them)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

template <typename T>
__global__ void __launch_bounds__(256) normalize_kernel(const T* __restrict__ data, const T* __restrict__ mean,
                                                        const T* __restrict__ norm, T scale, int inverse,
                                                        T* __restrict__ out, int64_t N, int dim) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * dim) return;
  const int j = (int)(t % dim);
  const T s = mul_rn<T>(norm[j], scale);
  out[t] = inverse ? add_rn<T>(mul_rn<T>(data[t], s), mean[j]) : (data[t] - mean[j]) / s;
}

// ------------------------------------------------------------------------------------------------------------
// min / max pooling backward              reference src/matten/nn/nodewise.py:142-148 (torch_scatter min/max)
// thread <-> (segment, column): the gradient goes to the first row holding the extreme value.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) segment_extreme_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g,
                                                                  const int32_t* __restrict__ ptr, int dim, int64_t B,
                                                                  int mode, T* __restrict__ dx) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= B * dim) return;
  const int64_t b = t / dim;
  const int j = (int)(t - b * dim);
  const int n0 = ptr[b], n1 = ptr[b + 1];
  if (n1 <= n0) return;
  int arg = n0;
  T best = x[(size_t)n0 * dim + j];
  for (int n = n0 + 1; n < n1; ++n) {
    T v = x[(size_t)n * dim + j];
    if (mode == 2 ? v < best : v > best) { best = v; arg = n; }
  }
  for (int n = n0; n < n1; ++n) dx[(size_t)n * dim + j] = n == arg ? g[t] : T(0);
}

}  // namespace mt

using namespace mt;

extern "C" {

int mt_instance_norm_fwd(int dtype, const void* x, const int32_t* graph_ptr, int64_t G, int dim, int num_channels,
                         const int32_t* chan_first, const int32_t* chan_dim, const int32_t* chan_scalar,
                         const void* weight, const void* bias, double eps, int reduce, int normalization, void* out,
                         void* save_mean, void* save_rstd, int32_t* save_arg, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0 && num_channels > 0, "instance_norm: empty irreps");
  MT_REQUIRE(reduce == 0 || reduce == 1, "instance_norm: reduce must be 0 (mean) or 1 (max)");
  MT_REQUIRE(normalization == 0 || normalization == 1, "instance_norm: normalization must be 0 (component) or 1 (norm)");
  if (G == 0) return MT_OK;
  MT_REQUIRE(x && graph_ptr && chan_first && chan_dim && chan_scalar && out && save_mean && save_rstd && save_arg,
             "null pointer");
  MT_REQUIRE(G <= 0x7fffffff, "instance_norm: too many graphs");
  dim3 grid((unsigned)G, (unsigned)ceil_div<int>(num_channels, 128));
  MT_DISPATCH_DTYPE(dtype, {
    instance_norm_fwd_kernel<T><<<grid, 128, 0, as_stream(stream)>>>(
        (const T*)x, graph_ptr, dim, num_channels, chan_first, chan_dim, chan_scalar, (const T*)weight, (const T*)bias,
        eps, reduce, normalization, (T*)out, (double*)save_mean, (double*)save_rstd, save_arg);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

int mt_instance_norm_bwd(int dtype, const void* x, const void* grad_out, const int32_t* graph_ptr, int64_t G, int dim,
                         int num_channels, const int32_t* chan_first, const int32_t* chan_dim,
                         const int32_t* chan_scalar, const void* weight, int reduce, int normalization,
                         const void* save_mean, const void* save_rstd, const int32_t* save_arg, void* grad_x,
                         void* grad_weight_part, void* grad_bias_part, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0 && num_channels > 0, "instance_norm: empty irreps");
  MT_REQUIRE((reduce == 0 || reduce == 1) && (normalization == 0 || normalization == 1), "instance_norm: bad mode");
  if (G == 0) return MT_OK;
  MT_REQUIRE(x && grad_out && graph_ptr && chan_first && chan_dim && chan_scalar && save_mean && save_rstd &&
                 save_arg && grad_x,
             "null pointer");
  MT_REQUIRE(G <= 0x7fffffff, "instance_norm: too many graphs");
  dim3 grid((unsigned)G, (unsigned)ceil_div<int>(num_channels, 128));
  MT_DISPATCH_DTYPE(dtype, {
    instance_norm_bwd_kernel<T><<<grid, 128, 0, as_stream(stream)>>>(
        (const T*)x, (const T*)grad_out, graph_ptr, dim, num_channels, chan_first, chan_dim, chan_scalar,
        (const T*)weight, reduce, normalization, (const double*)save_mean, (const double*)save_rstd, save_arg,
        (T*)grad_x,
        (T*)grad_weight_part, (T*)grad_bias_part);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

int mt_norm_act_fwd(int dtype, const void* x, int dim, int num_channels, const int32_t* chan_first,
                    const int32_t* chan_dim, int act_id, double epsilon, void* out, int64_t N, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0 && num_channels > 0 && epsilon > 0, "norm_act: bad arguments");
  if (N == 0) return MT_OK;
  MT_REQUIRE(x && chan_first && chan_dim && out, "null pointer");
  const int64_t total = N * num_channels;
  MT_DISPATCH_DTYPE(dtype, {
    norm_act_fwd_kernel<T><<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        (const T*)x, dim, num_channels, chan_first, chan_dim, act_id, T(epsilon * epsilon), (T*)out, N);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

int mt_norm_act_bwd(int dtype, const void* x, const void* grad_out, int dim, int num_channels,
                    const int32_t* chan_first, const int32_t* chan_dim, int act_id, double epsilon, void* grad_x,
                    int64_t N, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0 && num_channels > 0 && epsilon > 0, "norm_act: bad arguments");
  if (N == 0) return MT_OK;
  MT_REQUIRE(x && grad_out && chan_first && chan_dim && grad_x, "null pointer");
  const int64_t total = N * num_channels;
  MT_DISPATCH_DTYPE(dtype, {
    norm_act_bwd_kernel<T><<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        (const T*)x, (const T*)grad_out, dim, num_channels, chan_first, chan_dim, act_id, T(epsilon * epsilon),
        (T*)grad_x, N);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

int mt_normalize(int dtype, const void* data, const void* mean, const void* norm, double scale, int inverse,
                 void* out, int64_t N, int dim, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0, "normalize: dim must be positive");
  if (N == 0) return MT_OK;
  MT_REQUIRE(data && mean && norm && out, "null pointer");
  const int64_t total = N * dim;
  MT_DISPATCH_DTYPE(dtype, {
    normalize_kernel<T><<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        (const T*)data, (const T*)mean, (const T*)norm, T(scale), inverse, (T*)out, N, dim);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

int mt_segment_extreme_bwd(int dtype, const void* x, const void* grad_out, const int32_t* ptr, int dim, int64_t B,
                           int mode, void* grad_x, mt_stream stream) {
  MT_ENTRY_GUARD();
  MT_REQUIRE(dim > 0 && (mode == 2 || mode == 3), "segment_extreme_bwd handles min (2) and max (3)");
  if (B == 0) return MT_OK;
  MT_REQUIRE(x && grad_out && ptr && grad_x, "null pointer");
  const int64_t total = B * dim;
  MT_DISPATCH_DTYPE(dtype, {
    segment_extreme_bwd_kernel<T><<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        (const T*)x, (const T*)grad_out, ptr, dim, B, mode, (T*)grad_x);
  });
  MT_LAUNCH_OK();
  return MT_OK;
}

}  // extern "C"
