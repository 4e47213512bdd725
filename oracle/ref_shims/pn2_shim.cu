// extern "C" entry points around the reference's own pointnet2 CUDA launchers
// (utils/lib/src/*_gpu.cu, compiled IN PLACE from /root/reference by oracle/build_ref.py).
// TEST INFRASTRUCTURE ONLY: lets the GPU tests run the reference kernels on the B200 to pin
// oracle/l3d_oracle_group.c.  Only prototypes are declared here; no reference code is copied.
#include <cuda_runtime.h>

// prototypes of the launchers defined in ball_query_gpu.cu:48, group_points_gpu.cu:27,68,
// sampling_gpu.cu:26,65,211, interpolate_gpu.cu:60,127,172,217
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                     const float* new_xyz, const float* xyz, int* idx, cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                       const float* points, const int* idx, float* out, cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                            const float* grad_out, const int* idx, float* grad_points,
                                            cudaStream_t stream);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float* points,
                                        const int* idx, float* out, cudaStream_t stream);
void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float* dataset, float* temp,
                                             int* idxs, cudaStream_t stream);
void knn_kernel_launcher_fast(int b, int n, int m, int k, const float* unknown, const float* known,
                              float* dist2, int* idx, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float* unknown, const float* known,
                                   float* dist2, int* idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float* points,
                                            const int* idx, const float* weight, float* out,
                                            cudaStream_t stream);

extern "C" {
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                    const float* xyz, int* idx, void* s) {
  ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)s);
}
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                      const int* idx, float* out, void* s) {
  group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out, (cudaStream_t)s);
}
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                           const int* idx, float* grad_points, void* s) {
  group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, grad_out, idx, grad_points,
                                         (cudaStream_t)s);
}
void ref_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                       float* out, void* s) {
  gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out, (cudaStream_t)s);
}
void ref_fps(int b, int n, int m, const float* dataset, float* temp, int* idxs, void* s) {
  furthest_point_sampling_kernel_launcher(b, n, m, dataset, temp, idxs, (cudaStream_t)s);
}
void ref_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2,
             int* idx, void* s) {
  knn_kernel_launcher_fast(b, n, m, k, unknown, known, dist2, idx, (cudaStream_t)s);
}
void ref_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                  int* idx, void* s) {
  three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)s);
}
void ref_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                           const float* weight, float* out, void* s) {
  three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out, (cudaStream_t)s);
}
}
