// hair_gen.cuh — launchers of the generation / skinning kernels (see hair_gen.cu).
#pragma once
#include <cuda_runtime.h>

namespace bh {

cudaError_t launch_expand_strands(const float* root_pos3, const float* root_nrm3, const float* random_value,
                                  long long nstrands, int nverts, float scaleOffset, float4* pos, float4* vel,
                                  cudaStream_t stream);
cudaError_t launch_sphere_roots(const float* rowtab, const float* coltab, int rows, int cols, int column_major, long long first,
                                long long count, float* root_pos3, float* root_nrm3, cudaStream_t stream);
cudaError_t launch_patch_indices(const int* tri, long long nfaces, int nverts, int* out, cudaStream_t stream);
cudaError_t launch_skin_roots_dq(const float* rest_pos3, const int* joints4, const float* weights3, const float* dq,
                                 long long nstrands, int nverts, float4* pos, cudaStream_t stream);

// hair_tess.cu: tess-stream stage; out holds npatches * ninstances * nlines * nsub * 2 float4
cudaError_t launch_tess_stream(const float4* pos, const float4* tan, const int* patch, long long npatches, int nverts, float scale,
                               int ninstances, int nlines, int nsub, unsigned seed, float4* out, cudaStream_t stream);

}  // namespace bh
