/*
 * l3d_b200.h — C ABI of libl3d_b200.so, the B200 (sm_100a) drop-in for learning3d's
 * data-parallel hot path (pairwise distance / kNN / grouping, Chamfer, EMD, 3x3 Kabsch).
 *
 * Conventions:
 *   - plain pointers + sizes, no torch types;  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream);  every *_dev pointer is DEVICE memory, contiguous,
 *     fp32 / int32 / int64 exactly as named;  outputs are CALLER-allocated
 *     (the convention of the reference's `cd` and `pointnet2_cuda` modules).
 *   - return 0 on success, a positive cudaError_t on a CUDA failure, a negative L3D_ERR_*
 *     on a bad argument.  Nothing prints, nothing calls exit() (the reference printf()s /
 *     exit(-1)s: losses/cuda/chamfer_distance/chamfer_distance.cu:152-154,
 *     utils/lib/src/ball_query_gpu.cu:62-66).  No hidden synchronisation: kernels are
 *     enqueued on `stream` and the call returns.
 *   - *_host entry points take HOST buffers, do H2D -> kernel -> D2H on an internal stream
 *     and synchronise before returning (the end-to-end path bench.py reports as `e2e`).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * learning3d checkout).
 */
#ifndef L3D_B200_H_
#define L3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L3D_OK 0
#define L3D_ERR_INVALID (-1)
#define L3D_ERR_UNSUPPORTED (-2)

/* Largest candidate cloud the selection kernels keep resident in shared memory (the fast path).  Larger clouds are
 * streamed through a shared-memory tile with the same arithmetic and ordering (any N; k <= 2048 there). */
#define L3D_KNN_MAX_N 8192

/* ---- library ------------------------------------------------------------------- */
/* ABI version (bumped when a signature changes). */
int l3d_abi_version(void);
/* Human-readable text for a return code of any function in this header. */
const char* l3d_error_string(int code);
/* Number of kernels this library has launched in the calling process (all entry points).
 * bench.py differences it around the timed region to report `gpu_launches`. */
uint64_t l3d_launch_count(void);
/* Testing hook: nonzero forces the selection kernels onto their exact k-round slow path
 * (normally taken only when a row overflows the candidate buffer, e.g. duplicate points). */
void l3d_debug_force_slow_path(int on);
/* Testing hook (per host thread): which kernel knn() on xyz clouds takes when both apply.  0 = automatic (the
 * thread-per-row kernel of knn_tpr.cu when k <= 24, N % 32 == 0, 128 <= N <= 2048 and B*N/32 >= 512, else the
 * warp-per-row-pair kernel of knn.cu), 1 = never the thread-per-row kernel, 2 = the thread-per-row kernel whenever
 * the shape is eligible, whatever the batch size.  Both kernels return bit-identical results. */
void l3d_debug_knn_path(int path);

/* ---- kNN family ------------------------------------------------------------------ */
/*
 * knn() of utils/model_common_utils.py:3-9 for C == 3 (the DGCNN / DCP graph on xyz).
 *   x_dev   [B,3,N] fp32  ("bcn", exactly what the reference passes)
 *   idx_dev [B,N,k] int64, nearest first (self normally rank 0)
 *   val_dev optional [B,N,k] fp32: the reference's negated expansion distance
 *           pd = ((-|x_j|^2) + 2 x_i.x_j) - |x_i|^2   (model_common_utils.py:5-7)
 * Arithmetic: x_i.x_j = fma(z,z', fma(y,y', x*x')) — the K=3 GEMM accumulation order of
 * torch.matmul (verified against MKL, oracle/README.md); ties -> lower index first.
 * Requires 1 <= k <= N.  N <= L3D_KNN_MAX_N runs the resident kernel, larger clouds the streamed selection
 * (same results, k <= 2048).
 */
int l3d_knn_expansion(const float* x_dev, int B, int N, int k, int64_t* idx_dev,
                      float* val_dev, void* stream);

/*
 * knn() + get_graph_feature() of utils/model_common_utils.py:3-9,132-155 in ONE launch for the xyz graph
 * (C == 3): x_dev [B,3,N] -> idx_dev [B,N,k] int64 (required: autograd needs it) and
 * feat_dev [B,6,N,k] = cat(x[:, :, idx], x[:, :, n] repeated k).  The neighbour coordinates are still in
 * shared memory when a row's selection is final, so the separate gather launch and its re-read of the
 * indices disappear (SURVEY.md §8d "kNN+graph-feature fused").  Same results, bit for bit, as
 * l3d_knn_expansion followed by l3d_graph_feature.
 */
int l3d_knn_graph_feature(const float* x_dev, int B, int N, int k, int64_t* idx_dev, float* feat_dev,
                          void* stream);

/*
 * knn() of utils/model_common_utils.py:3-9 for any channel count C (the dynamic feature-space graphs of
 * PRNet's DGCNN, models/prnet.py:78-90: C = 64 / 64 / 128).  x_dev [B,C,N] -> idx_dev [B,N,k] int64.
 * Three launches on `stream`: |x_n|^2; the Gram matrix on the tensor cores (tcgen05, 3xTF32: fp32-class
 * accuracy, NOT bit-identical to a cuBLAS / MKL SGEMM — neither are those to each other) fused with
 * pd = ((-|x_j|^2) + 2 x_i.x_j) - |x_i|^2 into ws_dev; top-k per row (lower index first on exact ties).
 * ws_dev: l3d_knn_features_ws_bytes(B,C,N) bytes of device scratch (16-byte aligned), no initialisation needed.
 */
size_t l3d_knn_features_ws_bytes(int B, int C, int N);
int l3d_knn_features(const float* x_dev, int B, int C, int N, int k, int64_t* idx_dev, void* ws_dev,
                     void* stream);

/* Same, HOST buffers (pinned or pageable); copies + kernel + copy back, then syncs. */
int l3d_knn_expansion_host(const float* x_host, int B, int N, int k, int64_t* idx_host);

/*
 * get_graph_feature() of utils/model_common_utils.py:132-155 given the kNN indices:
 *   x_dev [B,C,N] fp32, idx_dev [B,N,k] int64 (values in [0,N)),
 *   out_dev [B,2C,N,k] fp32 = cat(x[:, :, idx], x[:, :, n] repeated k)   (:149-154)
 */
int l3d_graph_feature(const float* x_dev, const int64_t* idx_dev, int B, int C, int N, int k,
                      float* out_dev, void* stream);
/* Backward of the gather above: grad_x[B,C,N] (+)= scatter of grad_out[B,2C,N,k].
 * grad_x_dev must be zero-initialised by the caller (atomicAdd scatter). */
int l3d_graph_feature_grad(const float* grad_out_dev, const int64_t* idx_dev, int B, int C, int N,
                           int k, float* grad_x_dev, void* stream);

/*
 * knn_point(k, pos1, pos2) of utils/model_common_utils.py:84-100 (direct-difference kNN):
 *   data_dev [B,N,3] (pos1), query_dev [B,M,3] (pos2)
 *   val_dev [B,M,k] fp32 = sqrt(d2) nearest first, idx_dev [B,M,k] int64.
 * d2 = (dx*dx + dy*dy) + dz*dz, every operation rounded (torch elementwise order).
 */
int l3d_knn_point(const float* data_dev, const float* query_dev, int B, int N, int M, int k,
                  float* val_dev, int64_t* idx_dev, void* stream);

/*
 * knn_point(nsample, xyz, new_xyz) of utils/pointconv_util.py:107-118:
 * the nsample smallest entries of square_distance(new_xyz, xyz) (expansion form,
 * pointconv_util.py:18-39).  The reference asks topk(sorted=False); we return the same SET
 * in ascending-distance order.  idx_dev [B,S,nsample] int64.
 */
int l3d_knn_sqdist(const float* xyz_dev, const float* new_xyz_dev, int B, int N, int S,
                   int nsample, int64_t* idx_dev, void* stream);

/*
 * pointnet2_cuda.knn_wrapper(b,n,m,k,unknown,known,dist2,idx)
 * (utils/lib/src/pointnet2_api.cpp:22, interpolate_gpu.cu:9-57; called from
 * utils/lib/pointnet2_utils.py:96 and models/flownet3d.py:157,222):
 *   unknown_dev [b,n,3] queries, known_dev [b,m,3] data,
 *   dist2_dev [b,n,k] fp32 squared distance ascending, idx_dev [b,n,k] int32.
 * d2 = fma(dz,dz, fma(dy,dy, dx*dx))  (nvcc's contraction of the reference expression).
 * k <= 128 runs the threshold + sorting-network selection; larger k (the reference allows 200) the exact k-round scan.
 */
int l3d_pn2_knn(int b, int n, int m, int k, const float* unknown_dev, const float* known_dev,
                float* dist2_dev, int32_t* idx_dev, void* stream);
/* pointnet2_cuda.three_nn_wrapper (interpolate_gpu.cu:81-124): the k = 3 case. */
int l3d_pn2_three_nn(int b, int n, int m, const float* unknown_dev, const float* known_dev,
                     float* dist2_dev, int32_t* idx_dev, void* stream);

/* ---- Chamfer distance ---------------------------------------------------------------- */
/*
 * cd.forward_cuda(xyz1, xyz2, dist1, dist2, idx1, idx2)
 * (losses/cuda/chamfer_distance/chamfer_distance.cpp:27-40,180-185; kernel .cu:6-155; called from
 * losses/cuda/chamfer_distance/chamfer_distance.py:34):
 *   xyz1_dev [B,n,3], xyz2_dev [B,m,3] fp32 -> dist1_dev [B,n], dist2_dev [B,m] SQUARED nearest
 *   distance, idx1_dev/idx2_dev int32 arg-min (lowest index on ties).
 * Arithmetic is that of the reference's CPU nnsearch (chamfer_distance.cpp:59-87):
 * d = (dx*dx + dy*dy) + dz*dz, no fma — results are bit-identical to cd.forward.
 */
int l3d_chamfer_forward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m,
                        float* dist1_dev, float* dist2_dev, int32_t* idx1_dev, int32_t* idx2_dev,
                        void* stream);
/*
 * cd.backward_cuda(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
 * (chamfer_distance.cpp:42-57; kernel .cu:158-209; called from chamfer_distance.py:57).
 * Gather-form, deterministic, bit-identical to the CPU loops chamfer_distance.cpp:141-176;
 * the outputs are fully overwritten (no memset needed, unlike .cu:201-202).
 */
int l3d_chamfer_backward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m,
                         const float* graddist1_dev, const float* graddist2_dev,
                         const int32_t* idx1_dev, const int32_t* idx2_dev, float* gradxyz1_dev,
                         float* gradxyz2_dev, void* stream);
/*
 * Fused ChamferDistanceLoss (losses/chamfer_distance.py:34-51):
 *   loss = (mean(sqrt(dist1)) + mean(sqrt(dist2))) / 2 in the same launch as the NN search.
 * ws_dev: l3d_chamfer_ws_bytes(B,n,m) bytes of device scratch, zero-filled ONCE by the caller
 * (the arrival counter resets itself); one workspace per concurrently running stream.
 * loss_dev: device scalar.  dist/idx outputs as in l3d_chamfer_forward (kept for backward).
 */
size_t l3d_chamfer_ws_bytes(int B, int n, int m);
int l3d_chamfer_loss_forward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m,
                             float* dist1_dev, float* dist2_dev, int32_t* idx1_dev,
                             int32_t* idx2_dev, float* loss_dev, void* ws_dev, void* stream);
/* Backward of the fused loss: grad_loss_dev is the upstream scalar gradient ON THE DEVICE (no
 * host sync); chain rule of /2, mean, sqrt and the Chamfer gather in one launch. */
int l3d_chamfer_loss_backward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m,
                              const float* dist1_dev, const float* dist2_dev,
                              const int32_t* idx1_dev, const int32_t* idx2_dev,
                              const float* grad_loss_dev, float* gradxyz1_dev, float* gradxyz2_dev,
                              void* stream);

/* Host-buffer convenience call (like l3d_knn_expansion_host): Chamfer loss of losses/chamfer_distance.py:34-43 and
 * its gradients w.r.t. both clouds from HOST arrays xyz1_host [B,n,3], xyz2_host [B,m,3] -> loss_host [1],
 * grad1_host [B,n,3], grad2_host [B,m,3] (both NULL = forward only).  Copies, two launches, copies back, one sync. */
int l3d_chamfer_loss_fwd_bwd_host(const float* xyz1_host, const float* xyz2_host, int B, int n, int m,
                                  float* loss_host, float* grad1_host, float* grad2_host);

/* ---- pointnet2_cuda replacements (utils/lib/src/pointnet2_api.cpp:10-25) ------------------ */
/* Same argument order, caller-allocated outputs and int32 indices as the reference wrappers. */

/*
 * ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx)   (ball_query.cpp / _gpu.cu:9-45,
 * called from utils/lib/pointnet2_utils.py:247): for each of the m centres new_xyz_dev [b,m,3] the
 * first nsample indices k (ascending) of xyz_dev [b,n,3] with d2 < radius*radius, padded with the
 * first hit; a row with no hit is all 0 (the reference relies on the caller's .zero_(); here the
 * kernel writes every slot).  idx_dev [b,m,nsample] int32.
 */
int l3d_pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz_dev,
                       const float* xyz_dev, int32_t* idx_dev, void* stream);
/* group_points_wrapper(b,c,n,npoints,nsample,points,idx,out)  (group_points_gpu.cu:47-66;
 * pointnet2_utils.py:202): out[b,c,p,s] = points[b,c,idx[b,p,s]]. */
int l3d_pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points_dev,
                         const int32_t* idx_dev, float* out_dev, void* stream);
/* group_points_grad_wrapper (group_points_gpu.cu:8-25; pointnet2_utils.py:222): atomicAdd scatter
 * into grad_points_dev [b,c,n], which the caller zero-fills (as the reference does). */
int l3d_pn2_group_points_grad(int b, int c, int n, int npoints, int nsample,
                              const float* grad_out_dev, const int32_t* idx_dev,
                              float* grad_points_dev, void* stream);
/* gather_points_wrapper(b,c,n,npoints,points,idx,out)  (sampling_gpu.cu:8-24; pointnet2_utils.py:56):
 * out[b,c,j] = points[b,c,idx[b,j]]. */
int l3d_pn2_gather_points(int b, int c, int n, int npoints, const float* points_dev,
                          const int32_t* idx_dev, float* out_dev, void* stream);
/* gather_points_grad_wrapper (sampling_gpu.cu:46-63; pointnet2_utils.py:68). */
int l3d_pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out_dev,
                               const int32_t* idx_dev, float* grad_points_dev, void* stream);
/* furthest_point_sampling_wrapper(b,n,m,points,temp,idx)  (sampling_gpu.cu:93-246;
 * pointnet2_utils.py:28): temp_dev [b,n] must hold 1e10 on entry (caller-filled, as in the
 * reference) and holds the final min-distances on return; idxs_dev [b,m] int32, idxs[:,0] = 0.
 * Index-exact including the reference's tie rule.  n <= 8192 runs with the
 * cloud and the minima in registers / shared memory; larger clouds keep the minima in temp_dev. */
int l3d_pn2_furthest_point_sampling(int b, int n, int m, const float* dataset_dev, float* temp_dev,
                                    int32_t* idxs_dev, void* stream);
/* three_interpolate_wrapper(b,c,m,n,points,idx,weight,out)  (interpolate_gpu.cu:149-169;
 * pointnet2_utils.py:160) and its grad (interpolate_gpu.cu:192-214; pointnet2_utils.py:182). */
int l3d_pn2_three_interpolate(int b, int c, int m, int n, const float* points_dev,
                              const int32_t* idx_dev, const float* weight_dev, float* out_dev,
                              void* stream);
int l3d_pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out_dev,
                                   const int32_t* idx_dev, const float* weight_dev,
                                   float* grad_points_dev, void* stream);

/* ---- pure-torch grouping helpers ------------------------------------------------------------ */
/*
 * query_ball_point() in its three variants (utils/model_common_utils.py:102-130 [get_cnt],
 * utils/pointconv_util.py:85-105, utils/ppfnet_util.py:96-131 [itself_indices]):
 *   xyz_dev [B,N,3], new_xyz_dev [B,S,3]; radius2 = float32(radius**2) (the scalar the reference
 *   compares against); keeps expansion-form d2 <= radius2, first nsample in index order, padded
 *   with the first hit (or with itself_indices[b,s] when given, which is also excluded from the
 *   hits); a row without hits is all N (what the reference's sort leaves).  group_idx_dev
 *   [B,S,nsample] int64; cnt_dev optional [B,S] int64 = number of hits before truncation.
 */
int l3d_query_ball_point(const float* xyz_dev, const float* new_xyz_dev, int B, int N, int S,
                         float radius2, int nsample, const int64_t* itself_indices_dev,
                         int64_t* group_idx_dev, int64_t* cnt_dev, void* stream);
/*
 * farthest_point_sample() (model_common_utils.py:58-82, pointconv_util.py:60-83,
 * ppfnet_util.py:71-93): start_dev optional [B] int64 first indices (NULL = start at 0, the
 * pointconv / start_with_first_point variant); centroids_dev [B,npoint] int64.  N > 8192 takes a
 * stream-ordered scratch array (cudaMallocAsync) for the running minima.
 */
int l3d_farthest_point_sample(const float* xyz_dev, int B, int N, int npoint,
                              const int64_t* start_dev, int64_t* centroids_dev, void* stream);
/* square_distance(src, dst) (model_common_utils.py:19-38 and copies): out_dev [B,N,M]. */
int l3d_square_distance(const float* src_dev, const float* dst_dev, int B, int N, int M,
                        float* out_dev, void* stream);
/* index_points(points, idx) (model_common_utils.py:40-56 and copies): points_dev [B,N,C],
 * idx_dev [B,R] int64 (R = product of idx's trailing dims) -> out_dev [B,R,C]; and its backward
 * (atomicAdd into a caller-zeroed grad_points_dev [B,N,C]). */
int l3d_index_points(const float* points_dev, const int64_t* idx_dev, int B, int N, int64_t R, int C,
                     float* out_dev, void* stream);
int l3d_index_points_grad(const float* grad_out_dev, const int64_t* idx_dev, int B, int N, int64_t R,
                          int C, float* grad_points_dev, void* stream);
/* compute_density(xyz, bandwidth) (pointconv_util.py:199-209) as a fused row reduction:
 * density[b,i] = mean_j( exp(-d2_ij / two_bw2) / norm ), two_bw2 = float32(2*bw*bw),
 * norm = float32(2.5*bw). */
int l3d_compute_density(const float* xyz_dev, int B, int N, float two_bw2, float norm,
                        float* density_dev, void* stream);

/* ---- approximate EMD (losses/cuda/emd_torch) ---------------------------------------------- */
/*
 * emd_forward(xyz1, xyz2) -> [cost, match]   (pkg/include/emd.h:25-33, pkg/src/cuda/emd.cu:8-43;
 * kernels approxmatch emd.cuh:6-185 + matchcost emd.cuh:201-244; called from
 * pkg/layer/emd_loss_layer.py:10):  xyz1_dev [B,n,3], xyz2_dev [B,m,3] -> cost_dev [B] and,
 * optionally (NULL to skip), match_dev [B,n,m] in the reference's memory order
 * match[b*n*m + l*n + k] (l indexes xyz2, k indexes xyz1; emd.cuh:158).
 * The reference allocates its outputs and `temp`; here the caller passes them, plus ws_dev with
 * l3d_emd_forward_ws_bytes(B,n,m) bytes of scratch (no initialisation needed).  No
 * cudaDeviceSynchronize (the reference syncs inside, emd.cuh:197).  fp32; agreement with the
 * sequential restatement ~1e-6 relative.
 */
size_t l3d_emd_forward_ws_bytes(int B, int n, int m);
int l3d_emd_forward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m, float* cost_dev,
                    float* match_dev, void* ws_dev, void* stream);
/*
 * emd_backward(xyz1, xyz2, match) -> [grad_xyz1, grad_xyz2]   (pkg/include/emd.h:35-46,
 * emd.cu:45-70; kernels matchcostgrad1/2 emd.cuh:258-323): gradients of cost with the matching
 * held constant.  ws_dev: l3d_emd_backward_ws_bytes(B,n,m) bytes.  Deterministic (fixed-order
 * partial sums instead of one serial thread per point).
 */
size_t l3d_emd_backward_ws_bytes(int B, int n, int m);
int l3d_emd_backward(const float* xyz1_dev, const float* xyz2_dev, const float* match_dev, int B, int n,
                     int m, float* grad1_dev, float* grad2_dev, void* ws_dev, void* stream);

/* ---- DCP SVD head tail (utils/svd.py:29-58) -------------------------------------------------- */
/*
 * Given H = src_centered * src_corr_centered^T (H_dev [B,3,3]) and the two means (src_mean_dev,
 * corr_mean_dev [B,3]): R = V U^T of the SVD of H with the determinant fix of svd.py:40-45, and
 * t = -R*mean(src) + mean(src_corr) (:58).  Replaces the per-item torch.svd / torch.det loop and
 * its host synchronisation.  R_dev [B,3,3], t_dev [B,3].
 */
int l3d_kabsch3x3_batched(const float* H_dev, const float* src_mean_dev, const float* corr_mean_dev,
                          int B, float* R_dev, float* t_dev, void* stream);
/* Same, fused with the centring and the H reduction: src_dev, src_corr_dev [B,3,N] (svd.py:29-33). */
int l3d_svd_head_tail(const float* src_dev, const float* src_corr_dev, int B, int N, float* R_dev,
                      float* t_dev, void* stream);
/* Backward of l3d_svd_head_tail (what autograd derives through torch.svd / torch.det in the reference's
 * training scripts, examples/train_dcp.py): grad_R_dev [B,3,3], grad_t_dev [B,3] ->
 * grad_src_dev, grad_src_corr_dev [B,3,N].  Closed-form SVD differential, fp64 inside. */
int l3d_svd_head_tail_backward(const float* src_dev, const float* src_corr_dev, const float* grad_R_dev,
                               const float* grad_t_dev, int B, int N, float* grad_src_dev,
                               float* grad_src_corr_dev, void* stream);

/*
 * Front half of SVDHead.forward (utils/svd.py:23-28), fused, forward only:
 *   scores = softmax(src_emb^T . tgt_emb / sqrt(D), dim=2);  src_corr = tgt_xyz . scores^T
 * src_emb_dev [B,D,Ns], tgt_emb_dev [B,D,Nt], tgt_xyz_dev [B,3,Nt] -> src_corr_dev [B,3,Ns], all fp32.
 * The [B,Ns,Nt] score matrix is never written: tcgen05 3xTF32 GEMM tiles in TMEM + online softmax.
 * Tolerance vs the reference's fp32 matmul/softmax: 2e-5 absolute on src_corr (tests/test_gpu_softcorr.py).
 */
int l3d_soft_correspondence(const float* src_emb_dev, const float* tgt_emb_dev, const float* tgt_xyz_dev,
                            int B, int D, int Ns, int Nt, float* src_corr_dev, void* stream);
/* Test/debug aid: synchronises the device and returns the pipeline status word of
 * l3d_soft_correspondence (0 = ok, 1/2/3 = a bounded mbarrier wait of the epilogue / producer /
 * MMA-issuer role ran out), or a CUDA error code. */
int l3d_soft_correspondence_status(void);
/* Testing hook for the operand pipeline of l3d_soft_correspondence / l3d_knn_features:
 *   0 = automatic (TMA + CTA pairs when eligible: Ns, Nt multiples of 4, 16-byte aligned embeddings, Ns > 128),
 *   1 = force the generic (LDG producer, any shape) pipeline,  2 = TMA pipeline without CTA pairs. */
void l3d_debug_soft_correspondence_force_generic(int on);
/* Testing hook for the target-range split of l3d_soft_correspondence / l3d_knn_features (small batches
 * spread one row block's target tiles over several CTAs and merge the partial softmax states):
 * -1 = never split, 0 = automatic, n > 0 = force n splits. */
void l3d_debug_soft_correspondence_split(int n);
/* Debug aid: shared-memory image (4 x 4096 floats: A_hi, A_lo, B_hi, B_lo) of the first pipeline stage of
 * CTA (0,0) in the last l3d_debug_soft_correspondence_scores launch on the TMA path -> host_out. */
int l3d_debug_soft_correspondence_tiles(float* host_out);
/* Debug variant of l3d_soft_correspondence that also dumps the raw score accumulators
 * (src_emb^T . tgt_emb, before the 1/sqrt(D) scaling) to scores_dev [B,Ns,Nt]; used by the GEMM parity test. */
int l3d_debug_soft_correspondence_scores(const float* src_emb_dev, const float* tgt_emb_dev,
                                         const float* tgt_xyz_dev, int B, int D, int Ns, int Nt,
                                         float* src_corr_dev, float* scores_dev, void* stream);

/* ---- RPMNet matching tail (models/rpmnet.py:130-254) ----------------------------------------
 *
 * l3d_feature_square_distance: square_distance (utils/ppfnet_util.py:29-48) on C-dimensional features, i.e.
 * match_features(feat_src, feat_ref, 'l2') (rpmnet.py:130-154): src_dev [B,N,C], dst_dev [B,M,C] -> out_dev [B,N,M]
 * = |s|^2 + |d|^2 - 2 s.d, the Gram matrix on tcgen05 (3xTF32).  With beta_dev / alpha_dev ([B] each, both or
 * neither) the epilogue writes RPMNet's affinity -beta[b] * (dist - alpha[b]) (rpmnet.py:266-272) instead.
 * ws_dev: l3d_feature_square_distance_ws_bytes(B, N, M) bytes.  Toleranced like any fp32 GEMM.
 *
 * l3d_sinkhorn: sinkhorn(log_alpha, n_iters, slack) (rpmnet.py:157-218, eps <= 0): log_alpha_dev [B,J,K] ->
 * out_dev [B,J,K], the log of the (near) doubly stochastic matrix; slack != 0 adds the zero-padded slack row and
 * column.  The input is never rewritten: row / column potentials + one finishing pass.
 * ws_dev: l3d_sinkhorn_ws_bytes(B, J, K) bytes.
 *
 * l3d_rpm_match_tail: the same iterations fused with RPMNet.spam's tail (rpmnet.py:283-287):
 * perm = exp(sinkhorn(affinity)) -> perm_out_dev [B,J,K] (optional), rowsum_out_dev [B,J] = sum_k perm (optional),
 * weighted_out_dev [B,J,3] = perm @ xyz_ref / (rowsum + eps).  xyz_ref_dev [B,K,3].
 *
 * l3d_weighted_rigid_transform: compute_rigid_transform(a, b, weights) (rpmnet.py:221-254): a_dev, b_dev [B,M,3],
 * w_dev [B,M] -> T_dev [B,3,4] = [R | t], weighted Kabsch with the reference's determinant rule (third right
 * singular vector negated); eps = the reference's _EPS (1e-5, rpmnet.py:11) in w / (sum w + eps). */
/* Row-wise top-k of a key matrix keys_dev [rows, N] already in HBM: idx_dev [rows, k] int64, largest key first,
 * lower index on ties (the selection stage of l3d_knn_features).  knn_point() on C != 3 features
 * (model_common_utils.py:84-100) = l3d_feature_square_distance + this on the negated distances. */
int l3d_topk_rows(const float* keys_dev, long long rows, int N, int k, int64_t* idx_dev, void* stream);
size_t l3d_feature_square_distance_ws_bytes(int B, int N, int M);
int l3d_feature_square_distance(const float* src_dev, const float* dst_dev, int B, int N, int M, int C,
                                const float* beta_dev, const float* alpha_dev, float* out_dev, void* ws_dev,
                                void* stream);
size_t l3d_sinkhorn_ws_bytes(int B, int J, int K);
int l3d_sinkhorn(const float* log_alpha_dev, int B, int J, int K, int n_iters, int slack, float* out_dev,
                 void* ws_dev, void* stream);
int l3d_rpm_match_tail(const float* affinity_dev, const float* xyz_ref_dev, int B, int J, int K, int n_iters,
                       int slack, float eps, float* perm_out_dev, float* weighted_out_dev, float* rowsum_out_dev,
                       void* ws_dev, void* stream);
int l3d_weighted_rigid_transform(const float* a_dev, const float* b_dev, const float* w_dev, int B, int M, float eps,
                                 float* T_dev, void* stream);

/* Testing hook for the forward path of l3d_emd_forward: 0 = default (ONE cooperative persistent launch runs all 20
 * sweeps with per-item barriers), 1 = multi-launch (21 kernels), 2 = cooperative, 3 / 4 = one thread-block cluster
 * of 16 / 8 CTAs per item (hardware cluster barriers).  All paths produce the same vectors. */
void l3d_debug_emd_force_multilaunch(int mode);

/* ---- EdgeConv stack of DGCNN (models/dgcnn.py:32-48, eval mode) ---------------------
 *
 * l3d_edgeconv_layer1: conv1 (6 -> 64, 1x1, bias-free) + folded BatchNorm + ReLU computed straight from the kNN
 * indices — the [B,6,N,k] tensor of get_graph_feature (model_common_utils.py:132-155) is never written:
 *   h1[b,c,n*k+j] = relu(scale[c] * (w[c,0:3] . x[:, idx[b,n,j]] + w[c,3:6] . x[:, n]) + shift[c])
 * x_dev [B,3,N], idx_dev [B,N,k] int64 (l3d_knn_expansion's output); w_host [C1,6], scale_host, shift_host [C1]
 * are HOST arrays (2 KB, passed to the kernel by value); C1 must be 64 (L3D_ERR_UNSUPPORTED otherwise).
 * h1_dev [B,C1,N*k] and/or pool_dev (max over the k neighbours, written at
 * pool_dev[b*pool_bstride + (pool_coff+c)*N + n]) may be NULL.
 *
 * l3d_conv1x1_bn_relu_maxk: one 1x1 convolution layer as a tensor-core GEMM (tcgen05 3xTF32, fp32 accumulate
 * in TMEM), fused with folded BatchNorm, optional ReLU and the max over every group of G consecutive positions:
 *   y[b,m,p]    = act(scale[m] * sum_k wt[k,m] * x[b,k,p] + shift[m])            -> h_out_dev [B,M,P]   (optional)
 *   pool[b,m,n] = max_{j<G} y[b,m,n*G+j]   -> pool_out_dev[b*pool_bstride + (pool_coff+m)*(P/G) + n]   (optional)
 * wt_dev [K,M] is the TRANSPOSED weight (conv.weight[M,K].t().contiguous()), x_dev [B,K,P].  EdgeConv layers
 * 2-4 use G = k (P = N*k); conv5 uses G = 1, pool_out_dev = NULL.  Requires P % 4 == 0, M % 4 == 0, 16-byte
 * aligned wt/x (L3D_ERR_UNSUPPORTED otherwise) and P % G == 0, G <= 256 when pooling.
 * Accuracy: that of an fp32 GEMM (|err| <~ 2^-21 sum|w||x|).  No device synchronisation; a pipeline time-out
 * (never observed) writes NaN to the outputs and is reported by l3d_edgeconv_status(). */
int l3d_edgeconv_layer1(const float* x_dev, const int64_t* idx_dev, const float* w_host, const float* scale_host,
                        const float* shift_host, int B, int N, int k, int C1, float* h1_dev, float* pool_dev,
                        long long pool_bstride, int pool_coff, void* stream);
int l3d_conv1x1_bn_relu_maxk(const float* wt_dev, const float* x_dev, const float* scale_dev,
                             const float* shift_dev, int B, int M, int K, int P, int G, int relu, float* h_out_dev,
                             float* pool_out_dev, long long pool_bstride, int pool_coff, void* stream);
/* ---- DCP's transformer (utils/transformer.py), activations channel-major [B, d_model, N] -------------------------
 *
 * l3d_linear_cm: nn.Linear on channel-major activations, tcgen05 3xTF32:
 *   out[b,m,p] = act(sum_k wt[k,m] * x[b,k,p] + bias[m]) (+ residual[b,m,p])     wt_dev = linear.weight.t() [K,M]
 * bias_dev / residual_dev / col_div_dev may be NULL; col_div_dev [B,P] divides every output column (after bias and
 * activation, before the residual).  w_heads = h > 0 selects "one weight head per item" (the P.V product of
 * attention): x_dev holds B = batch*h items, wt_dev is [batch, K, h*M], item b uses columns (b % h)*M.. of weight
 * batch b / h, M must be a multiple of 128.  Same shape requirements as l3d_conv1x1_bn_relu_maxk.
 *
 * l3d_attention_stats / l3d_attention_probs_t: softmax(q^T k / sqrt(D)) of transformer.py:17-23 in two passes of
 * the tcgen05 score pipeline (scores never written): q_dev [BH,D,Nq], k_dev [BH,D,Nk] (BH = batch*heads, the
 * [B, h*d_k, N] projections viewed per head) -> stats_dev [BH,Nq,2] (row max in log2 units, row sum) and then
 * probs_t_dev [BH,Nk,Nq], the probabilities TRANSPOSED.  Fast protocol: stats with precise = 0 (ONE TF32 pass: the
 * max only serves as exponent reference), probs with normalized = 0 (writes 2^(s - max) and stores the exact row
 * sums into stats_dev[..,1]), then l3d_linear_cm(..., col_div_dev = those sums).  The result is — exactly the activation operand l3d_linear_cm(w_heads)
 * needs for out[bh, d_v, q] = sum_k v^T[k, d_v] p^T[k, q].
 *
 * l3d_layernorm_cm: LayerNorm of transformer.py:128-137 over the channel axis of x_dev [B,D,N]:
 *   a2 * (x - mean) / (std_unbiased + eps) + b2. */
/* Backward of l3d_soft_correspondence w.r.t. the scaled scores (DCP training, utils/svd.py:23-28 under autograd):
 * dS[i,j] = P[i,j] (g_i . tgt_j - g_i . src_corr_i) / sqrt(D), P = softmax(src_emb^T tgt_emb / sqrt(D)) recomputed on
 * tcgen05 from stats_dev (l3d_attention_stats); written plain (ds_dev [B,Ns,Nt]) and / or transposed
 * (ds_t_dev [B,Nt,Ns]).  The embedding gradients follow as two l3d_linear_cm calls with per-item weights:
 * d src_emb = tgt_emb . dS^T, d tgt_emb = src_emb . dS. */
int l3d_soft_correspondence_dscores(const float* src_emb_dev, const float* tgt_emb_dev, const float* tgt_xyz_dev,
                                    const float* stats_dev, const float* grad_corr_dev, const float* corr_dev, int B,
                                    int D, int Ns, int Nt, float* ds_dev, float* ds_t_dev, void* stream);
int l3d_linear_cm(const float* wt_dev, const float* x_dev, const float* bias_dev, const float* residual_dev,
                  const float* col_div_dev, int B, int M, int K, int P, int relu, int w_heads, float* out_dev,
                  void* stream);
/* l3d_linear_cm with the output transposed: out_dev [B, P, M] (positions major) — the attention value projection
 * writes v^T [B, N_k, h*d_v], the per-head weight operand of the p.v product, directly. */
int l3d_linear_cm_t(const float* wt_dev, const float* x_dev, const float* bias_dev, int B, int M, int K, int P, int relu,
                    float* out_dev, void* stream);
int l3d_attention_stats(const float* q_dev, const float* k_dev, int BH, int D, int Nq, int Nk, int precise,
                        float* stats_dev, void* stream);
int l3d_attention_probs_t(const float* q_dev, const float* k_dev, float* stats_dev, int BH, int D, int Nq, int Nk,
                          int normalized, float* probs_t_dev, void* stream);
/* Bound-referenced softmax: l3d_attention_bounds writes stats_dev[(bh,i),0] = |q_i| max_j|k_j| log2(e)
 * / sqrt(D) — an upper bound of the row's scaled scores, a valid exponent reference for l3d_attention_probs_t(normalized
 * = 0) whenever it is <= 40 — and raises the flag in the LAST int of ws_dev (l3d_attention_bounds_ws_bytes) otherwise;
 * l3d_attention_stats_if is l3d_attention_stats(precise = 0) that runs only when that flag is set (device-side
 * decision, no host round trip). */
size_t l3d_attention_bounds_ws_bytes(int BH);
int l3d_attention_bounds(const float* q_dev, const float* k_dev, int BH, int D, int Nq, int Nk, float* stats_dev,
                         void* ws_dev, void* stream);
int l3d_attention_stats_if(const float* q_dev, const float* k_dev, int BH, int D, int Nq, int Nk, const int* flag_dev,
                           float* stats_dev, void* stream);
int l3d_layernorm_cm(const float* x_dev, const float* a2_dev, const float* b2_dev, float eps, int B, int D, int N,
                     float* out_dev, void* stream);
/* Synchronises the device; returns and clears the pipeline error word of l3d_conv1x1_bn_relu_maxk (0 = ok). */
int l3d_edgeconv_status(void);

#ifdef __cplusplus
}
#endif
#endif /* L3D_B200_H_ */
