// Tile scheduler / pixel post-processing launchers (tiles.cu).
#pragma once
#include "gemm.cuh"

namespace ir {

int tile_gather_launch(const float* src, float* dst, const int* coords, int ntiles, int N, int C, int H, int W, int th,
                       int tw, int scale, cudaStream_t s);
int tile_blend_launch(const float* tiles, const int* coords, int ntiles, float* out, int N, int C, int H, int W, int th,
                      int tw, int scale, cudaStream_t s);
size_t wavelet_workspace_bytes(int N, int C, int H, int W);
int wavelet_reconstruction_launch(const float* content, const float* style, float* out, int N, int C, int H, int W,
                                  void* workspace, size_t workspace_bytes, cudaStream_t s);
int adain_launch(const float* content, const float* style, float* out, int N, int C, int HW, cudaStream_t s);
int to_uint8_launch(const float* img, uint8_t* out, int N, int C, int H, int W, cudaStream_t s);

}  // namespace ir
