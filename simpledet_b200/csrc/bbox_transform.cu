// Device versions of operator_py/bbox_transform.py and operator_py/cython/{bbox,bbox_self}.pyx — the
// box utilities the reference's loader threads and test-time code call on the CPU.
//   overlaps:            float32 with the Cython-generated double promotions (bit-exact)
//   encode / decode:     float64 like numpy (nonlinear_transform, nonlinear_pred + clip_boxes, iou_pred)
// All are one thread per output element; they are latency-sized ops whose value is staying on the
// device between the kernels of this library.
#include "common.cuh"

namespace {

__device__ __forceinline__ float fminr(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float fmaxr(float a, float b) { return a < b ? b : a; }

// mode 0: bbox_overlaps_cython (bbox.pyx:32-73) IoU; mode 1: bbox_selfoverlaps_cython (bbox_self.pyx:32-75)
__global__ void __launch_bounds__(256) overlaps_kernel(const float4* __restrict__ boxes, const float4* __restrict__ query,
                                                       float* __restrict__ out, const int N, const int K,
                                                       const int mode) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i - (size_t)n * K);
  const float4 b = __ldg(boxes + n), q = __ldg(query + k);
  float v = 0.f;
  const float iw = (float)((double)__fsub_rn(fminr(b.z, q.z), fmaxr(b.x, q.x)) + 1.0);
  if (iw > 0) {
    const float ih = (float)((double)__fsub_rn(fminr(b.w, q.w), fmaxr(b.y, q.y)) + 1.0);
    if (ih > 0) {
      const float inter = __fmul_rn(iw, ih);
      const double barea = ((double)__fsub_rn(b.z, b.x) + 1.0) * ((double)__fsub_rn(b.w, b.y) + 1.0);
      if (mode == 0) {
        const float qarea = (float)(((double)__fsub_rn(q.z, q.x) + 1.0) * ((double)__fsub_rn(q.w, q.y) + 1.0));
        const float ua = (float)((barea + (double)qarea) - (double)inter);
        v = __fdiv_rn(inter, ua);
      } else {
        v = __fdiv_rn(inter, (float)barea);
      }
    }
  }
  out[i] = v;
}

// nonlinear_transform (bbox_transform.py:52-78), float64
__global__ void __launch_bounds__(256) encode_kernel(const double* __restrict__ ex, const double* __restrict__ gt,
                                                     double* __restrict__ out, const int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double* e = ex + (size_t)n * 4;
  const double* g = gt + (size_t)n * 4;
  const double ew = e[2] - e[0] + 1.0, eh = e[3] - e[1] + 1.0;
  const double ecx = e[0] + 0.5 * (ew - 1.0), ecy = e[1] + 0.5 * (eh - 1.0);
  const double gw = g[2] - g[0] + 1.0, gh = g[3] - g[1] + 1.0;
  const double gcx = g[0] + 0.5 * (gw - 1.0), gcy = g[1] + 0.5 * (gh - 1.0);
  double* o = out + (size_t)n * 4;
  o[0] = (gcx - ecx) / (ew + 1e-14);
  o[1] = (gcy - ecy) / (eh + 1e-14);
  o[2] = log(gw / ew);
  o[3] = log(gh / eh);
}

// nonlinear_pred (:81-120) / iou_pred (:129-161), optionally followed by clip_boxes (:34-49).
// boxes (N,4) float32 (cast to float64 like `.astype(np.float)`), deltas/out (N,4*K) float64.
__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ boxes, const double* __restrict__ deltas,
                                                     double* __restrict__ out, const int N, const int K,
                                                     const int iou, const int clip, const double im_h,
                                                     const double im_w) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * K) return;
  const int n = (int)(i / K);
  const float* b = boxes + (size_t)n * 4;
  const double x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
  const double* d = deltas + i * 4;
  double o0, o1, o2, o3;
  if (iou) {
    o0 = d[0] + x1; o1 = d[1] + y1; o2 = d[2] + x2; o3 = d[3] + y2;
  } else {
    const double w = x2 - x1 + 1.0, h = y2 - y1 + 1.0;
    const double cx = x1 + 0.5 * (w - 1.0), cy = y1 + 0.5 * (h - 1.0);
    const double clipv = 4.135166556742356;  // np.log(1000. / 16.), bbox_transform.py:5
    const double dw = d[2] < clipv ? d[2] : clipv, dh = d[3] < clipv ? d[3] : clipv;  // np.minimum
    const double pcx = d[0] * w + cx, pcy = d[1] * h + cy;
    const double pw = exp(dw) * w, ph = exp(dh) * h;
    o0 = pcx - 0.5 * (pw - 1.0); o1 = pcy - 0.5 * (ph - 1.0);
    o2 = pcx + 0.5 * (pw - 1.0); o3 = pcy + 0.5 * (ph - 1.0);
  }
  if (clip) {  // np.maximum(np.minimum(v, im - 1), 0)
    const double mx = im_w - 1, my = im_h - 1;
    o0 = fmax(fmin(o0, mx), 0.0); o1 = fmax(fmin(o1, my), 0.0);
    o2 = fmax(fmin(o2, mx), 0.0); o3 = fmax(fmin(o3, my), 0.0);
  }
  double* o = out + i * 4;
  o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
}

// flip_boxes (bbox_transform.py:164-169): x1' = W - x2 - 1, x2' = W - x1 - 1, in the boxes' own dtype
template <typename T>
__global__ void __launch_bounds__(256) flip_kernel(const T* __restrict__ boxes, T* __restrict__ out, const size_t nbox,
                                                   const T im_width) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbox) return;
  const T x1 = boxes[i * 4], y1 = boxes[i * 4 + 1], x2 = boxes[i * 4 + 2], y2 = boxes[i * 4 + 3];
  out[i * 4] = im_width - x2 - (T)1;
  out[i * 4 + 1] = y1;
  out[i * 4 + 2] = im_width - x1 - (T)1;
  out[i * 4 + 3] = y2;
}

// box_voting (bbox_transform.py:172-221): one warp per top box sweeps all_dets; float32 throughout
// (np.average of float32 boxes with float32 weights stays float32).  method: 0 ID, 1 TEMP_AVG, 2 AVG,
// 3 IOU_AVG, 4 GENERALIZED_AVG, 5 QUASI_SUM.
__global__ void __launch_bounds__(256) box_voting_kernel(const float* __restrict__ top, const float* __restrict__ all,
                                                         float* __restrict__ out, const int T, const int N,
                                                         const float thresh, const int method, const float beta) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= T) return;
  const float4 b = make_float4(top[k * 5], top[k * 5 + 1], top[k * 5 + 2], top[k * 5 + 3]);
  float sw = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, cnt = 0.f, aux = 0.f, aux2 = 0.f;
  for (int j = lane; j < N; j += 32) {
    const float* q = all + (size_t)j * 5;
    float v = 0.f;  // bbox_overlaps_cython with its double promotions
    const float iw = (float)((double)__fsub_rn(fminr(b.z, q[2]), fmaxr(b.x, q[0])) + 1.0);
    if (iw > 0) {
      const float ih = (float)((double)__fsub_rn(fminr(b.w, q[3]), fmaxr(b.y, q[1])) + 1.0);
      if (ih > 0) {
        const float inter = __fmul_rn(iw, ih);
        const double barea = ((double)__fsub_rn(b.z, b.x) + 1.0) * ((double)__fsub_rn(b.w, b.y) + 1.0);
        const float qarea = (float)(((double)__fsub_rn(q[2], q[0]) + 1.0) * ((double)__fsub_rn(q[3], q[1]) + 1.0));
        v = __fdiv_rn(inter, (float)((barea + (double)qarea) - (double)inter));
      }
    }
    if (!(v >= thresh)) continue;
    const float w = q[4];
    sw += w; s0 += w * q[0]; s1 += w * q[1]; s2 += w * q[2]; s3 += w * q[3];
    cnt += 1.f;
    if (method == 1) {         // softmax of (log P / beta) over {w, 1-w}, first component
      const float p1 = 1.0f - w, pm = fmaxf(w, p1);
      const float e0 = expf(logf(w / pm) / beta), e1 = expf(logf(p1 / pm) / beta);
      aux += e0 / (e0 + e1);
    } else if (method == 3) {  // IoU-weighted mean of the scores
      aux += w * v;
      aux2 += v;
    } else if (method == 4) {
      aux += powf(w, beta);
    }
  }
  for (int o = 16; o; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o); s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    aux += __shfl_xor_sync(0xffffffffu, aux, o); aux2 += __shfl_xor_sync(0xffffffffu, aux2, o);
  }
  if (lane == 0) {
    float* o = out + (size_t)k * 5;
    o[0] = s0 / sw; o[1] = s1 / sw; o[2] = s2 / sw; o[3] = s3 / sw;
    float sc = top[k * 5 + 4];
    if (method == 1) sc = aux / cnt;
    else if (method == 2) sc = sw / cnt;
    else if (method == 3) sc = aux / aux2;
    else if (method == 4) sc = powf(aux / cnt, 1.0f / beta);
    else if (method == 5) sc = sw / powf(cnt, beta);
    o[4] = sc;
  }
}

}  // namespace

extern "C" int sdet_bbox_flip(const void* boxes, void* out, size_t num_boxes, double im_width, int is_double,
                              void* stream) {
  if (num_boxes == 0) return SDET_OK;
  SDET_REQUIRE(boxes && out, "NULL argument");
  const unsigned grid = (unsigned)((num_boxes + 255) / 256);
  if (is_double)
    flip_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const double*>(boxes), static_cast<double*>(out), num_boxes, im_width);
  else
    flip_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(boxes), static_cast<float*>(out), num_boxes, (float)im_width);
  SDET_LAUNCH_CHECK("flip_kernel");
  return SDET_OK;
}

extern "C" int sdet_box_voting(const float* top_dets, const float* all_dets, float* out, int T, int N, float thresh,
                               int scoring_method, float beta, void* stream) {
  SDET_REQUIRE(T >= 0 && N >= 0 && scoring_method >= 0 && scoring_method <= 5, "bad argument");
  if (T == 0) return SDET_OK;
  SDET_REQUIRE(top_dets && all_dets && out, "NULL argument");
  box_voting_kernel<<<(unsigned)((T * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(top_dets, all_dets, out, T, N,
                                                                                      thresh, scoring_method, beta);
  SDET_LAUNCH_CHECK("box_voting_kernel");
  return SDET_OK;
}

extern "C" int sdet_bbox_overlaps(const float* boxes, const float* query_boxes, float* overlaps, int N, int K,
                                  int mode, void* stream) {
  SDET_REQUIRE(N >= 0 && K >= 0 && (mode == 0 || mode == 1), "bad argument");
  if (N == 0 || K == 0) return SDET_OK;
  SDET_REQUIRE(boxes && query_boxes && overlaps, "NULL argument");
  SDET_REQUIRE(((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(query_boxes)) & 15) == 0,
               "boxes must be 16-byte aligned");
  const size_t total = (size_t)N * K;
  overlaps_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), reinterpret_cast<const float4*>(query_boxes), overlaps, N, K, mode);
  SDET_LAUNCH_CHECK("overlaps_kernel");
  return SDET_OK;
}

extern "C" int sdet_bbox_nonlinear_transform(const double* ex_rois, const double* gt_rois, double* targets, int N,
                                             void* stream) {
  SDET_REQUIRE(N >= 0, "bad argument");
  if (N == 0) return SDET_OK;
  SDET_REQUIRE(ex_rois && gt_rois && targets, "NULL argument");
  encode_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ex_rois, gt_rois, targets, N);
  SDET_LAUNCH_CHECK("encode_kernel");
  return SDET_OK;
}

extern "C" int sdet_bbox_pred(const float* boxes, const double* box_deltas, double* pred_boxes, int N, int K,
                              int iou, int clip, double im_h, double im_w, void* stream) {
  SDET_REQUIRE(N >= 0 && K >= 0, "bad argument");
  if (N == 0 || K == 0) return SDET_OK;
  SDET_REQUIRE(boxes && box_deltas && pred_boxes, "NULL argument");
  const size_t total = (size_t)N * K;
  decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boxes, box_deltas, pred_boxes, N, K,
                                                                                  iou ? 1 : 0, clip ? 1 : 0, im_h, im_w);
  SDET_LAUNCH_CHECK("decode_kernel");
  return SDET_OK;
}
