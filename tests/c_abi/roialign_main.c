/* A plain-C caller of the drop-in boundary: no Python, no torch, only the CUDA runtime for memory.
 * usage: roialign_main <in.bin> <out.bin>
 * in.bin : int32 B,N,C,H,W,PH,PW; float32 spatial_scale; data[B*C*H*W]; rois[B*N*4]
 * out.bin: out[B*N*C*PH*PW], argmax_x[...], argmax_y[...]                                  */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "simpledet_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "cuda: %s\n", cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc != 3) return 64;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 66;
  int hdr[7];
  float scale;
  if (fread(hdr, 4, 7, f) != 7 || fread(&scale, 4, 1, f) != 1) return 65;
  const int B = hdr[0], N = hdr[1], C = hdr[2], H = hdr[3], W = hdr[4], PH = hdr[5], PW = hdr[6];
  const size_t nd = (size_t)B * C * H * W, nr = (size_t)B * N * 4, no = (size_t)B * N * C * PH * PW;
  float* hd = (float*)malloc(nd * 4);
  float* hr = (float*)malloc(nr * 4);
  float* ho = (float*)malloc(no * 4 * 3);
  if (fread(hd, 4, nd, f) != nd || fread(hr, 4, nr, f) != nr) return 65;
  fclose(f);
  float *dd, *dr, *dout, *dax, *day;
  void* ws;
  const size_t wsb = sdet_roi_align_v2_workspace(B, N);
  CK(cudaMalloc((void**)&dd, nd * 4)); CK(cudaMalloc((void**)&dr, nr * 4));
  CK(cudaMalloc((void**)&dout, no * 4)); CK(cudaMalloc((void**)&dax, no * 4)); CK(cudaMalloc((void**)&day, no * 4));
  CK(cudaMalloc(&ws, wsb));
  CK(cudaMemcpy(dd, hd, nd * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr, hr, nr * 4, cudaMemcpyHostToDevice));
  if (sdet_abi_version() != 4) { fprintf(stderr, "unexpected ABI version %d\n", sdet_abi_version()); return 3; }
  /* an invalid call must fail loudly and leave a message */
  if (sdet_roi_align_v2_forward(dd, dr, dout, dax, NULL, B, N, C, H, W, PH, PW, scale, ws, wsb, NULL) == SDET_OK) return 4;
  if (!sdet_last_error() || !sdet_last_error()[0]) return 5;
  const int rc = sdet_roi_align_v2_forward(dd, dr, dout, dax, day, B, N, C, H, W, PH, PW, scale, ws, wsb, NULL);
  if (rc != SDET_OK) { fprintf(stderr, "sdet: %s\n", sdet_last_error()); return 6; }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(ho, dout, no * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho + no, dax, no * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho + 2 * no, day, no * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[2], "wb");
  if (!f || fwrite(ho, 4, no * 3, f) != no * 3) return 73;
  fclose(f);
  printf("launches=%llu\n", (unsigned long long)sdet_launch_count());
  return 0;
}
