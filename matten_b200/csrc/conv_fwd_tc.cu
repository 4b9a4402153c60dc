// Host launcher of the tcgen05 convolution forward (see conv_fwd_tc.cuh).
#include "conv_fwd_tc.cuh"

namespace mt {

// returns MT_OK and sets *used = 1 when the tensor-core path ran; *used = 0 when the plan/shape does not
// qualify (caller falls back to the FMA-pipe kernel)
int conv_fwd_tc_try(const mt_conv_plan* plan, const void* x, const void* sh, const void* emb,
                    const void* const* mlp_weights, const int32_t* rowptr, const int32_t* perm,
                    const int32_t* src_sorted, double avg, const void* num_neigh, void* out, int64_t N, int64_t E,
                    cudaStream_t st, int* used) {
  *used = 0;
  if (plan->tc_num_tiles <= 0 || plan->tc_num_tiles > kTcMaxTiles) return MT_OK;
  if (plan->tc_num_sub <= 0 || plan->tc_num_sub > kTcMaxSub) return MT_OK;
  if (!plan->tc_row_wcol || !plan->tc_sub_hdr || !plan->tc_sub_slot || !plan->tc_q_list) return MT_OK;
  for (int i = 0; i < plan->mlp_num_layers; ++i)
    if (plan->mlp_sizes[i] > kTcK) return MT_OK;
  if (E >= (int64_t)2147483647 || N >= (int64_t)2147483647) return MT_OK;
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.x_dim = plan->x_dim;
  p.y_dim = plan->y_dim;
  p.out_dim = plan->out_dim;
  p.num_tiles = plan->tc_num_tiles;
  p.num_sub = plan->tc_num_sub;
  p.row_wcol = plan->tc_row_wcol;
  p.sub_hdr = plan->tc_sub_hdr;
  p.sub_slot = plan->tc_sub_slot;
  p.q_list = plan->tc_q_list;
  for (int q = 0; q < 4; ++q) p.q_count[q] = plan->tc_q_count[q];
  p.nl = plan->mlp_num_layers;
  for (int i = 0; i <= p.nl; ++i) p.sizes[i] = plan->mlp_sizes[i];
  for (int i = 0; i < p.nl; ++i) p.w[i] = static_cast<const float*>(mlp_weights[i]);
  p.act = plan->mlp_act;
  p.act_cst = (float)plan->mlp_act_cst;
  p.x = static_cast<const float*>(x);
  p.sh = static_cast<const float*>(sh);
  p.emb = static_cast<const float*>(emb);
  p.rowptr = rowptr;
  p.perm = perm;
  p.src = src_sorted;
  p.avg = (float)avg;
  p.num_neigh = static_cast<const float*>(num_neigh);
  p.out = static_cast<float*>(out);
  p.N = N;
  p.E = E;
  p.xs_stride = p.x_dim;
  const TcSmemLayout L = tc_smem_layout(p.num_tiles, p.xs_stride, p.y_dim);
  if (L.total > (size_t)227 * 1024 - 256) return MT_OK;  // does not fit: fall back
  static thread_local size_t configured = 0;
  if (L.total > configured) {
    MT_CUDA_OK(cudaFuncSetAttribute(conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    configured = L.total;
  }
  int64_t grid = kNumSMs;
  if (grid > N) grid = N;
  if (grid < 1) grid = 1;
  conv_fwd_tc_kernel<<<(unsigned)grid, kTcThreads, L.total, st>>>(p);
  MT_LAUNCH_OK();
  *used = 1;
  return MT_OK;
}

}  // namespace mt
