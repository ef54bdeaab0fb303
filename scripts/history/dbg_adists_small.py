import torch, numpy as np, warnings, sys
sys.path.insert(0, '.')
warnings.simplefilter('ignore')
from nerf_qa_b200 import synth
from oracle import reference_oracle as O
from tests._util import make_adists
m = make_adists("fp32")
ws, bs = synth.make_vgg_weights(1234)
for shape in [(1,8,8),(1,12,30),(1,16,16),(1,32,32),(1,48,48),(1,64,64),(1,21,21),(1,22,22),(1,40,21)]:
    b,h,w = shape
    x,y = synth.make_frame_pairs(b,h,w,seed=43)
    with torch.no_grad():
        fx, fy = O.trunk(x, ws, bs), O.trunk(y, ws, bs)
        want = O.adists_from_feats(fx, fy).numpy()
        got_ff = m.forward_from_feats([f.cuda() for f in fx], [f.cuda() for f in fy]).cpu().numpy()
        got = m(x.cuda(), y.cuda(), as_loss=False).cpu().numpy()
    print(shape, 'oracle', want, 'from_feats', got_ff, 'full', got)
