// uba_vo.h — glue between uba_vo.cu (pose-only mode) and the handle of uba_host.cu.
#ifndef UBA_VO_H_INCLUDED
#define UBA_VO_H_INCLUDED
#include <cuda_runtime.h>

#include "../../include/uba.h"

struct uba_vo_state;
uba_vo_state* uba_vo_new();
void uba_vo_free(uba_vo_state* s);
// implemented in uba_host.cu (they need the handle's layout)
uba_vo_state* uba_vo_get(uba_handle* h, bool create);   // selects the handle's device
cudaStream_t uba_vo_stream(uba_handle* h);
int uba_vo_fail(uba_handle* h, int code, const char* what, const char* detail);
void uba_vo_count(uba_handle* h, int kernels);
#endif
