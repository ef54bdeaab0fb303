import torch, numpy as np, warnings, sys
sys.path.insert(0, '.')
warnings.simplefilter('ignore')
from nerf_qa_b200 import synth, _lib
from oracle import reference_oracle as O
from tests._util import make_dists
ws, bs = synth.make_vgg_weights(1234)
torch.set_printoptions(precision=5, linewidth=200)
for prec, P in (("fp32", _lib.PREC_FP32), ("bf16", _lib.PREC_BF16)):
  m = make_dists(prec)
  for shape in [(1,1,1),(1,1,7),(1,3,5)]:
    b,h,w = shape
    g = torch.Generator().manual_seed(41)
    x = torch.rand(b,3,h,w,generator=g); y = (x + 0.1*torch.randn(b,3,h,w,generator=g)).clamp(0,1)
    with torch.no_grad():
        fx, fy = O.trunk(x, ws, bs), O.trunk(y, ws, bs)
        ctx = m._context(torch.device('cuda:0')); m._sync_alpha_beta(ctx)
        score, s1s2 = ctx.dists_forward(x.cuda(), y.cuda(), P, want_s1s2=True)
        s_ff, s1s2_ff = ctx.dists_from_feats([f.cuda() for f in fx], [f.cuda() for f in fy])
    s1, s2 = O.dists_level_stats(fx[0], fy[0])
    print(prec, shape, 'x', x.flatten()[:3], 'y', y.flatten()[:3])
    print('  L0 S1 ours', s1s2[0,0,:3].cpu(), 'from_feats', s1s2_ff[0,0,:3].cpu(), 'oracle', s1[0])
    print('  L0 S2 ours', s1s2[0,1,:3].cpu(), 'from_feats', s1s2_ff[0,1,:3].cpu(), 'oracle', s2[0])
    s1, s2 = O.dists_level_stats(fx[1], fy[1])
    print('  L1 S1 ours', s1s2[0,0,3:9].cpu(), 'from_feats', s1s2_ff[0,0,3:9].cpu(), 'oracle', s1[0,:6])
    print('  L1 S2 ours', s1s2[0,1,3:9].cpu(), 'from_feats', s1s2_ff[0,1,3:9].cpu(), 'oracle', s2[0,:6])
