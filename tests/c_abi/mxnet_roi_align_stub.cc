// The MXNet-side binding of INTEGRATION.md §1, as a real translation unit: it replaces the GPU compute functions of
// operator_cxx/contrib/roi_align_v2.cu (ROIAlignForward_v2<gpu> declared at roi_align_v2-inl.h:157-162 and
// ROIAlignBackward_v2<gpu> at :198-203) with one call each into include/simpledet_b200.h.
// tests/test_integration_stub.py compiles it (syntax + types) against the reference's own header and the
// MXNet stand-in of oracle/shim, so the documented snippet cannot drift from the C ABI.
#include "./roi_align_v2-inl.h"   // the reference's header: ROIAlignParam_v2, roialign_v2::k*

#include "simpledet_b200.h"

#define SDET_CALL(x)                                                                                  \
  do {                                                                                                \
    if (int rc = (x)) LOG(FATAL) << "simpledet_b200: " << sdet_last_error() << " (code " << rc << ")"; \
  } while (0)

namespace mxnet {
namespace op {

template <>
void ROIAlignForward_v2<gpu>(const nnvm::NodeAttrs& attrs, const OpContext& ctx, const std::vector<TBlob>& in,
                             const std::vector<OpReqType>& req, const std::vector<TBlob>& out) {
  const ROIAlignParam_v2& p = nnvm::get<ROIAlignParam_v2>(attrs.parsed);
  cudaStream_t st = mshadow::Stream<gpu>::GetStream(ctx.get_stream<gpu>());
  const bool train = ctx.is_train;  // argmax_x / argmax_y are only needed for the backward pass
  const int B = in[roialign_v2::kData].size(0), N = in[roialign_v2::kBox].size(1);
  // the plan / band-list scratch: ctx.requested[kTempSpace] sized by sdet_roi_align_v2_workspace(B, N)
  const size_t ws_bytes = sdet_roi_align_v2_workspace(B, N);
  mshadow::Tensor<gpu, 1, uint8_t> ws = ctx.requested[0].get_space_typed<gpu, 1, uint8_t>(
      mshadow::Shape1(static_cast<mshadow::index_t>(ws_bytes)), ctx.get_stream<gpu>());
  SDET_CALL(sdet_roi_align_v2_forward(
      in[roialign_v2::kData].dptr<float>(), in[roialign_v2::kBox].dptr<float>(), out[roialign_v2::kOut].dptr<float>(),
      train ? out[roialign_v2::kMaxIdx_x].dptr<float>() : nullptr,
      train ? out[roialign_v2::kMaxIdx_y].dptr<float>() : nullptr, B, N, in[roialign_v2::kData].size(1),
      in[roialign_v2::kData].size(2), in[roialign_v2::kData].size(3), p.pooled_size[0], p.pooled_size[1],
      p.spatial_scale, ws.dptr_, ws_bytes, st));
}

template <>
void ROIAlignBackward_v2<gpu>(const nnvm::NodeAttrs& attrs, const OpContext& ctx, const std::vector<TBlob>& inputs,
                              const std::vector<OpReqType>& req, const std::vector<TBlob>& outputs) {
  // inputs = {ograd, rois, argmax_x, argmax_y} (ROIAlignGrad_v2, roi_align_v2-inl.h:206-218); outputs = {grad_data, grad_rois}
  cudaStream_t st = mshadow::Stream<gpu>::GetStream(ctx.get_stream<gpu>());
  CHECK_NE(req[roialign_v2::kData], kWriteInplace) << "ROIAlign: Backward doesn't support kWriteInplace.";
  if (req[roialign_v2::kData] == kNullOp) return;
  SDET_CALL(sdet_roi_align_v2_backward(
      inputs[0].dptr<float>(), inputs[2].dptr<float>(), inputs[3].dptr<float>(), outputs[0].dptr<float>(),
      req[roialign_v2::kBox] == kWriteTo ? outputs[1].dptr<float>() : nullptr, inputs[1].size(0), inputs[1].size(1),
      outputs[0].size(1), outputs[0].size(2), outputs[0].size(3), inputs[0].size(3), inputs[0].size(4),
      /*accumulate=*/req[roialign_v2::kData] == kAddTo, st));
}

}  // namespace op
}  // namespace mxnet
