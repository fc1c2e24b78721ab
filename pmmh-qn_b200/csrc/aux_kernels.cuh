// aux_kernels.cuh -- host-side launchers of aux_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pmmh {

inline size_t sv_align_aux(size_t x) { return (x + 255) & ~(size_t)255; }

cudaError_t launch_transpose(const double* in, double* out, long long rows, int cols, int batch,
                             long long in_stride, long long out_stride, cudaStream_t st);
cudaError_t launch_copy_head(const double* in, double* out, int n, int batch, long long in_stride,
                             long long out_stride, cudaStream_t st);
cudaError_t launch_norm_cdf(const double* in, double* out, long long n, cudaStream_t st);
cudaError_t launch_crank_nicolson(const double* u, const double* xi, double* out, long long n,
                                  double a, double b, unsigned long long seed,
                                  unsigned long long offset, cudaStream_t st);
cudaError_t launch_importance_discrete(const double* obs, long long obs_stride, const double* params,
                                       const double* rvr, const double* rvp, int nobs, int n, int batch,
                                       double* filt, double* ll, double* traj, double* grad,
                                       int* traj_idx, cudaStream_t st);
size_t subsample_ws_bytes(int m);
cudaError_t launch_subsample_indices(const double* u, int m, int n, int apply_cdf, int* idx,
                                     double* sorted, void* ws, cudaStream_t st);
size_t logistic_ws_bytes(int m, int d, int hess);
cudaError_t launch_logistic(const double* x, const double* y, const int* idx, int m, int d,
                            long long row_begin, long long row_end, const double* beta, int hess,
                            double* out, void* ws, cudaStream_t st);

}  // namespace pmmh
