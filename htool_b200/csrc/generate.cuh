// htool_b200/csrc/generate.cuh — launch interface of the device leaf assembly (generate.cu).
#ifndef HTB_GENERATE_CUH
#define HTB_GENERATE_CUH

#include "store.hpp"
#include <cuda_runtime.h>
#include <htool_b200.h>

namespace htb {
bool kernel_is_complex(int kernel);
// Evaluates the built-in kernel function at the points of every task's unit, straight into side 0's stream.
// target_points / source_points: 3 doubles per index of the root block, CLUSTER numbering, on the device.
cudaError_t launch_generate_dense(int kernel, const DenseTask *tasks, long long n_tasks, unsigned char *stream, const double *target_points, const double *source_points, double wavenumber, cudaStream_t st);
// Device assembly: the stage headers of a side, packed back to back on the host (Packer::fill_headers), copied to their places
// in the (zeroed) stream. One warp per stage; header st = compact[hdr_off[st], hdr_off[st + 1]) -> stream + stages[st].byte_off.
cudaError_t launch_scatter_headers(const StageDesc *stages, const unsigned long long *hdr_off, long long n_stages, const unsigned char *compact, unsigned char *stream, cudaStream_t st);
} // namespace htb
#endif
