// Experiment-only code (compiled with -DGNN_TUNE by tools/build_tune.sh; never part of libgnn_b200.so).
// Included in the middle of gnn_kernels.cu, so it sees its helpers (kFull, cdiv, GNN_LAUNCH_CHECK, ...).
#ifdef GNN_TUNE
int gnn_debug_spmm_hub(const int *rowptr, const int *colidx, const float *vals, int M, int nnz, int D, const float *X, int ldx,
                       float *Y, int ldy, const short *hubslot, const int *hubcols, int H, int K, int ranges, int unr,
                       gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int slabs = (int)cdiv(D, 32);
  const size_t smem = (size_t)H * 32 * sizeof(float) + (((size_t)K * 2 + 15) / 16) * 16;
  const unsigned grid = (unsigned)(slabs * ranges);
#define HB(U_)                                                                                                    \
  do {                                                                                                            \
    GNN_CUDA(cudaFuncSetAttribute(spmm_hub_proto_kernel<U_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    spmm_hub_proto_kernel<U_><<<grid, 1024, smem, st>>>(rowptr, colidx, vals, M, nnz, D, X, ldx, Y, ldy, hubslot, hubcols, H, K, ranges); \
  } while (0)
  if (unr == 1) HB(1); else if (unr == 2) HB(2); else if (unr == 4) HB(4); else HB(8);
#undef HB
  GNN_LAUNCH_CHECK();
  return 0;
}
#endif

#ifdef GNN_TUNE
int gnn_debug_gather_roof(const float *X, int ldx, int K, const int *colidx, int nnz, int nv, int u, int warps, int per_warp,
                          float *sink, gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)cdiv(warps, 8);
#define GR(NV_, U_) gather_roof_kernel<NV_, U_><<<grid, 256, 0, st>>>(X, ldx, K, colidx, nnz, per_warp, sink)
  if (nv == 8 && u == 1) GR(8, 1); else if (nv == 8 && u == 2) GR(8, 2);
  else if (nv == 4 && u == 2) GR(4, 2); else if (nv == 4 && u == 4) GR(4, 4);
  else if (nv == 2 && u == 4) GR(2, 4); else if (nv == 2 && u == 8) GR(2, 8);
  else if (nv == 1 && u == 8) GR(1, 8); else if (nv == 1 && u == 16) GR(1, 16);
  else return GNN_E_BADARG;
#undef GR
  GNN_LAUNCH_CHECK();
  return 0;
}
#endif

