/* The reference's own C ABI, called exactly as gpu_nms.pyx:13-31 would after linking this library instead of
 * nms_kernel.cu: only the declaration from gpu_nms.hpp (copied below as a prototype, it is a 2-line interface),
 * host buffers, no CUDA calls in the caller.
 * usage: nms_main <in.bin> <out.bin>      in.bin: int32 n, dim; float32 thresh; boxes[n*dim] (score-sorted)
 *                                          out.bin: int32 num_out, keep[num_out]                            */
#include <stdio.h>
#include <stdlib.h>

void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
          float nms_overlap_thresh, int device_id);

int main(int argc, char** argv) {
  if (argc != 3) return 64;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 66;
  int hdr[2];
  float thresh;
  if (fread(hdr, 4, 2, f) != 2 || fread(&thresh, 4, 1, f) != 1) return 65;
  const int n = hdr[0], dim = hdr[1];
  float* boxes = (float*)malloc((size_t)n * dim * 4);
  int* keep = (int*)malloc((size_t)n * 4);
  if (fread(boxes, 4, (size_t)n * dim, f) != (size_t)n * dim) return 65;
  fclose(f);
  int num_out = -1;
  _nms(keep, &num_out, boxes, n, dim, thresh, 0);
  f = fopen(argv[2], "wb");
  if (!f || fwrite(&num_out, 4, 1, f) != 1 || fwrite(keep, 4, (size_t)(num_out > 0 ? num_out : 0), f) != (size_t)(num_out > 0 ? num_out : 0)) return 73;
  fclose(f);
  return 0;
}
