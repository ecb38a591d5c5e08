import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import math, torch
import ranking_b200 as tfr
from ranking_b200 import _C
from oracle import losses_impl as OL, utils as OU
import test_parity_gpu as T

b, n, s_ = 5, 17, 3
scores, labels, item_w, list_w = T._batch(b, n, seed=71)
u = T._hash_uniforms((5 << 32) | 1, b * s_ * n).reshape(b, s_, n).double()
teacher = torch.empty(b * s_, n, device='cuda')
_C.check(_C.lib.tfr_gumbel_sample(None, _C.ptr(labels.cuda().contiguous()), b, n, s_, 1.0, (5 << 32) | 1, 1,
                                  _C.ptr(teacher), None, None, _C.stream()))
mask = labels >= 0
tl = torch.where(mask, labels, torch.full_like(labels, math.log(1e-10))).double()
g = -torch.log(-torch.log(u + 1e-20) + 1e-20)
ref_t = torch.log(torch.softmax(tl.unsqueeze(1) + g, -1) + 1e-10).reshape(b * s_, n)
print('teacher max abs diff', float((teacher.cpu().double() - ref_t).abs().max()))
print('order equal', torch.equal(torch.argsort(-teacher.cpu(), stable=True), torch.argsort(-ref_t, stable=True)))
lc = tfr.losses_impl.CoupledRankDistilLoss(sample_size=s_, temperature=0.8)
lc.seed(5)
l1, w1 = lc.compute_per_list(labels.cuda(), scores.cuda(), None)
lo = OL.CoupledRankDistilLoss(sample_size=s_, temperature=0.8)
lo.uniforms = u
l2, w2 = lo.compute_per_list(labels.double(), scores.double(), None)
print('cuda', l1.cpu().tolist(), w1.cpu().tolist())
print('orac', l2.tolist(), w2.tolist())
lc.seed(5)
l3, _ = lc._run(labels.cuda(), scores.cuda(), None, None, 0.8)
lo2 = OL.CoupledRankDistilLoss(sample_size=s_, temperature=0.8)
lo2.uniforms = u
l4, _ = lo2._compute_unreduced_loss_impl(labels.double(), scores.double() / 0.8, mask)
print('T=0.8 cuda', l3.cpu().tolist()); print('T=0.8 orac', l4.reshape(-1).tolist())
