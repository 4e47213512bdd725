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

/* Largest candidate cloud the selection kernels keep resident in shared memory. */
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

/* ---- kNN family ------------------------------------------------------------------ */
/*
 * knn() of utils/model_common_utils.py:3-9 for C == 3 (the DGCNN / DCP graph on xyz).
 *   x_dev   [B,3,N] fp32  ("bcn", exactly what the reference passes)
 *   idx_dev [B,N,k] int64, nearest first (self normally rank 0)
 *   val_dev optional [B,N,k] fp32: the reference's negated expansion distance
 *           pd = ((-|x_j|^2) + 2 x_i.x_j) - |x_i|^2   (model_common_utils.py:5-7)
 * Arithmetic: x_i.x_j = fma(z,z', fma(y,y', x*x')) — the K=3 GEMM accumulation order of
 * torch.matmul (verified against MKL, oracle/README.md); ties -> lower index first.
 * Requires 1 <= k <= N <= L3D_KNN_MAX_N.
 */
int l3d_knn_expansion(const float* x_dev, int B, int N, int k, int64_t* idx_dev,
                      float* val_dev, void* stream);

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
 * k <= 128 here (the reference allows 200).
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

#ifdef __cplusplus
}
#endif
#endif /* L3D_B200_H_ */
