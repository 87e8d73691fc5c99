#include "cuda_runtime.h"
namespace uba_emu { thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim; }
