// Host launcher of the tcgen05 convolution forward (see conv_fwd_tc.cuh).
#include <stdlib.h>

#include "conv_fwd_tc.cuh"

namespace mt {

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// widest chunk (MMA N) whose staging buffers fit the shared memory; 0: none
static int tc_pick_ne(const mt_conv_plan* plan) {
  const int y_pad = (plan->y_dim + 3) & ~3;
  for (int ne = tc_chunk_cols(plan->tc_num_tiles); ne >= 64; ne >>= 1) {
    const TcSmemLayout L = tc_smem_layout(plan->tc_num_tiles, plan->x_dim, y_pad, plan->tc_num_sub, ne);
    if (L.total + 512 <= (size_t)227 * 1024) return ne;  // dynamic + static shared memory must fit 227 KB
  }
  return 0;
}

static bool tc_plan_qualifies(const mt_conv_plan* plan) {
  if (plan->tc_num_tiles <= 0 || plan->tc_num_tiles > kTcMaxTiles) return false;
  if (plan->tc_num_sub <= 0 || plan->tc_num_sub > kTcMaxSub) return false;
  if (!plan->tc_row_wcol || !plan->tc_sub_hdr || !plan->tc_sub_slot || !plan->tc_q_list) return false;
  if (plan->mlp_num_layers > 3) return false;  // at most two hidden layers in the preparation kernel
  for (int i = 0; i < plan->mlp_num_layers; ++i)
    if (plan->mlp_sizes[i] > kTcK) return false;
  if (plan->x_dim > 65535 || plan->y_dim > 252 || (plan->x_dim & 3) != 0) return false;  // 16-byte bulk copies
  return tc_pick_ne(plan) > 0;
}

size_t conv_fwd_tc_workspace_bytes(const mt_conv_plan* plan, int64_t E) {
  if (!tc_plan_qualifies(plan) || E <= 0) return 0;
  const int y_pad = (plan->y_dim + 3) & ~3;
  return align256((size_t)E * 192) + align256((size_t)E * y_pad * 4) + 256;
}

// returns MT_OK and sets *used = 1 when the tensor-core path ran; *used = 0 when the plan/shape does not
// qualify (caller falls back to the FMA-pipe kernel)
int conv_fwd_tc_try(const mt_conv_plan* plan, const void* x, const void* sh, const void* emb,
                    const void* const* mlp_weights, const int32_t* rowptr, const int32_t* perm,
                    const int32_t* src_sorted, double avg, const void* num_neigh, void* out, void* workspace,
                    size_t workspace_bytes, int64_t N, int64_t E, cudaStream_t st, int* used) {
  *used = 0;
  if (!tc_plan_qualifies(plan)) return MT_OK;
  if (E >= (int64_t)2147483647 || N >= (int64_t)2147483647) return MT_OK;
  const size_t need = conv_fwd_tc_workspace_bytes(plan, E);
  if (E > 0 && (workspace == nullptr || workspace_bytes < need)) return MT_OK;
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
  p.y_pad = (p.y_dim + 3) & ~3;
  {
    // workspace: 256-byte aligned sub-buffers
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
    p.hplanes = reinterpret_cast<__nv_bfloat16*>(base);
    p.ysorted = reinterpret_cast<float*>(base + align256((size_t)E * 192));
  }
  const char* dbg = getenv("MT_CONV_TC_DEBUG");
  p.dbg = (dbg && *dbg) ? reinterpret_cast<long long*>(strtoull(dbg, nullptr, 10)) : nullptr;
  const int NE = tc_pick_ne(plan);
  const TcSmemLayout L = tc_smem_layout(p.num_tiles, p.x_dim, p.y_pad, p.num_sub, NE);
  if (E > 0) {
    int64_t g = ceil_div<int64_t>(E, kPrepThreads);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    edge_prepare_kernel<<<(unsigned)g, kPrepThreads, 0, st>>>(p);
    MT_LAUNCH_OK();
  }
  int64_t grid = kNumSMs;
  if (grid > N) grid = N;
  if (grid < 1) grid = 1;
  static thread_local size_t configured[3] = {0, 0, 0};
  const int vi = NE == 256 ? 0 : (NE == 128 ? 1 : 2);
  if (L.total > configured[vi]) {
    if (NE == 256)
      MT_CUDA_OK(cudaFuncSetAttribute(conv_fwd_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    else if (NE == 128)
      MT_CUDA_OK(cudaFuncSetAttribute(conv_fwd_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    else
      MT_CUDA_OK(cudaFuncSetAttribute(conv_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    configured[vi] = L.total;
  }
  if (NE == 256) conv_fwd_tc_kernel<256><<<(unsigned)grid, tc_threads(256), L.total, st>>>(p);
  else if (NE == 128) conv_fwd_tc_kernel<128><<<(unsigned)grid, tc_threads(128), L.total, st>>>(p);
  else conv_fwd_tc_kernel<64><<<(unsigned)grid, tc_threads(64), L.total, st>>>(p);
  MT_LAUNCH_OK();
  *used = 1;
  return MT_OK;
}

}  // namespace mt
