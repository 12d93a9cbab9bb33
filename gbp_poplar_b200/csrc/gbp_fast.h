// Interface of the opt-in contracted-FMA build of the sweep kernel (gbp_fast.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <cstring>

size_t gbp_fast_graph_bytes();  // sizeof(DeviceGraph) as that translation unit sees it (must equal the caller's)
// k_sweep_tma<PREP, true, UPPER> with contracted arithmetic; returns a cudaError_t
int gbp_fast_launch_sweep(const void* graph_bytes, const void* maps_bytes, int prep, int upper, unsigned grid, cudaStream_t s);
