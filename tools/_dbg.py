import numpy as np, torch, sys
sys.path.insert(0,'.')
from simpledet_b200 import _lib
from simpledet_b200._lib import check
cuda=torch.device('cuda:0')
for (stride,dilate,pad,dg,C) in [(1,1,1,4,8),(1,1,1,1,64)]:
    rng = np.random.default_rng(1)
    B,H,W=2,13,19
    data = rng.standard_normal((B, C, H, W)).astype(np.float32)
    Ho = (H + 2 * pad - (dilate * 2 + 1)) // stride + 1
    Wo = (W + 2 * pad - (dilate * 2 + 1)) // stride + 1
    offset = (rng.standard_normal((B, dg * 18, Ho, Wo)) * 3).astype(np.float32)
    d=torch.from_numpy(data).to(cuda); o=torch.from_numpy(offset).to(cuda)
    col = torch.empty((B, C * 9, Ho * Wo), device=cuda)
    check(_lib.lib().sdet_deformable_im2col(d.data_ptr(), o.data_ptr(), col.data_ptr(), B, C, H, W, 3,3,pad,pad,stride,stride,dilate,dilate,dg,None))
    x = d.permute(0, 2, 3, 1).contiguous()
    col_t = torch.full((B, Ho * Wo, 9, C), -77.0, device=cuda)
    check(_lib.lib().sdet_deformable_im2col_nhwc(x.data_ptr(), o.data_ptr(), col_t.data_ptr(), B, C, H, W, 3, 3, pad, pad, stride, stride, dilate, dilate, dg, None))
    torch.cuda.synchronize()
    a=col_t.permute(0, 3, 2, 1).reshape(B, C * 9, Ho * Wo)
    bad=(a!=col).nonzero()
    print(C,dg,'bad',bad.shape[0],'of',a.numel(), 'untouched', int((a==-77).sum()))
    for r in bad[:8].tolist():
        b,k,p=r; print(r,'c',k//9,'t',k%9,'p',p, float(a[b,k,p]), float(col[b,k,p]))
