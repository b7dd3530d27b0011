#!/usr/bin/env python
"""Compile the reference's OWN C++/CUDA operator sources, unmodified and where they lie under
/root/reference/operator_cxx, into oracle/_ref/libref_cxx.so — test infrastructure only.

`operator_cxx/**` includes MXNet / mshadow / nnvm / dmlc headers that are not in this image (and the
reference's build is "drop the files into an MXNet checkout and build MXNet": not runnable here).  The
few types those files touch are supplied by oracle/shim/mxnet_shim.h (force-included; the shim tree also
answers their relative `#include "../mshadow_op.h"` etc.).  The `.cu` files are compiled as plain C++:
their device functors run on the host through the shim's serial `Kernel<OP, gpu>::Launch` and
`atomicAdd`.  proposal_target*.cc call std::random_shuffle -> rand(): the undefined `rand` symbol of
those two objects is renamed to `ref_shim_rand` (objcopy), which oracle/ref_harness.cc defines, so
the pin tests control the shuffles.  Nothing from the reference is copied into the repo; objects are
built in a temp dir and only the .so lands in oracle/_ref/ (git-ignored, travels to the GPU box).
"""
import os
import subprocess
import sys
import tempfile

REF = "/root/reference/operator_cxx"
HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "shim")
OUT = os.path.join(HERE, "_ref", "libref_cxx.so")

# (source under operator_cxx, needs the rand rename)
SOURCES = [
    ("contrib/roi_align_v2.cc", False),   # functor roi_align_v2-inl.h:61-153, CPU gather backward, registration
    ("contrib/roi_align_v2.cu", False),   # ROIAlignBackwardKernelGPU_v2 (:17-85) + driver (:88-143)
    ("roi_pooling_v1.cc", False, ["-DSHIM_GPU_DISPATCH"]),  # ROIPoolForward_v1 / ROIPoolBackwardAcc_v1 (:40-221) + op
    ("contrib/decodebbox.cc", False),     # BBoxTransformXYWH/XYXY (:34-133) + DecodeBBoxOp::Forward
    ("proposal_target.cc", True),         # SampleROI, BBoxOverlap, targets (:22-227) + ProposalTargetOp::Forward
    ("proposal_target_v2.cc", True),
    # ProposalMaskTarget: the operator against a stand-in maskApi.h (oracle/shim/coco_api: cocoapi is not in the tree)
    ("proposal_mask_target.cc", True),
    ("contrib/generate_anchor.cc", False, ["-DSHIM_GPU_DISPATCH"]),  # GenAnchorOp<cpu>::Forward + gen_anchor_utils (generate_anchor-inl.h:139-183)
    ("contrib/focal_loss.cc", False),     # FocalLossOp::Forward / Backward as mshadow expressions (focal_loss-inl.h:100-231)
    ("contrib/bbox_norm.cc", False),      # BBoxNormOp::Backward (bbox_norm-inl.h:99-129)
    # GPU-only operators.  Plain C++ cannot parse `kernel<<<grid, block, ...>>>(args)`, so the .cu files below pass
    # through ONE textual rewrite on their way to the compiler (a temporary copy, never written into the repository):
    #     kernel<<<cfg>>>(args)   ->   shim_launch(cfg).run([&](auto... a) { kernel(a...); }, args)
    # and nothing else; shim_launch (oracle/shim/mxnet_shim.h) runs every thread of every block in turn.
    # SigmoidCrossEntropy: the .cc says NotImplemented for the CPU; kernels + mshadow reductions :43-120.
    ("contrib/sigmoid_cross_entropy.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/sigmoid_cross_entropy.cu", False, [], "rewrite_launches"),
    # Proposal_v3: the GPU operator SimpleDet runs (the CPU twin in the .cc indexes its scores out of range);
    # thrust::stable_sort_by_key / cudaMemcpy come from oracle/shim/thrust, the shim's host stand-ins.
    ("contrib/proposal_v3.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/proposal_v3.cu", False, [], "rewrite_launches"),
    # the same treatment for the remaining proposal-family GPU operators
    ("contrib/proposal.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/proposal.cu", False, [], "rewrite_launches"),
    ("contrib/proposal_v2.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/proposal_v2.cu", False, [], "rewrite_launches"),
    ("contrib/nms.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/nms.cu", False, [], "rewrite_launches"),
    ("contrib/generate_proposal.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/generate_proposal.cu", False, [], "rewrite_launches"),
    ("contrib/generate_proposal_retina.cc", False, ["-DSHIM_GPU_DISPATCH"]),
    ("contrib/generate_proposal_retina.cu", False, [], "rewrite_launches"),
    # GPU twins of operators whose .cc is already pinned: the kernels the reference actually runs
    ("roi_pooling_v1.cu", False, [], "rewrite_launches"),
    ("contrib/generate_anchor.cu", False, [], "rewrite_launches"),
]
# -O2 without -march: like MXNet's x86-64 CPU build there is no FMA instruction to contract into;
# -ffp-contract=off makes that explicit.
CXXFLAGS = ["-std=c++14", "-O2", "-fPIC", "-ffp-contract=off", "-w"]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "ref_harness.cc"), os.path.join(SHIM, "mxnet_shim.h"), __file__]
    deps += [os.path.join(REF, e[0]) for e in SOURCES]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def main() -> int:
    if not os.path.isdir(REF):
        print("reference not present; keeping prebuilt oracle/_ref/libref_cxx.so", file=sys.stderr)
        return 0
    if not stale():
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    inc = ["-I", SHIM, "-I", os.path.join(SHIM, "l1"), "-I", os.path.join(SHIM, "l1", "l2")]
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for entry in SOURCES:
            src, rename_rand = entry[0], entry[1]
            extra = list(entry[2]) if len(entry) > 2 else []
            path = os.path.join(REF, src)
            if len(entry) > 3 and entry[3] == "rewrite_launches":
                import re

                text = re.sub(r"\b([A-Za-z_]\w*(?:<[\w:, ]+>)?)\s*<<<(.*?)>>>\s*\(",
                              lambda m: "shim_launch(%s).run([&](auto... a) { %s(a...); }, " % (m.group(2), m.group(1)),
                              open(path).read(), flags=re.S)
                path = os.path.join(tmp, os.path.basename(src) + ".cc")
                open(path, "w").write(text)
                extra += ["-iquote", os.path.dirname(os.path.join(REF, src)), "-DSHIM_CUDA_DEVICE_MATH"]
            obj = os.path.join(tmp, src.replace("/", "_").replace(".", "_") + ".o")
            cmd = ["g++", *CXXFLAGS, *extra, *inc, "-include", os.path.join(SHIM, "mxnet_shim.h"), "-x", "c++", "-c",
                   path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stderr[-4000:], file=sys.stderr)
                return 1
            if rename_rand:
                subprocess.run(["objcopy", "--redefine-sym", "rand=ref_shim_rand", obj], check=True)
            objs.append(obj)
        hobj = os.path.join(tmp, "ref_harness.o")
        r = subprocess.run(["g++", *CXXFLAGS, "-I", HERE, "-c", os.path.join(HERE, "ref_harness.cc"), "-o", hobj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-4000:], file=sys.stderr)
            return 1
        r = subprocess.run(["g++", "-shared", "-o", OUT, hobj, *objs], capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-4000:], file=sys.stderr)
            return 1
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
