"""debug helper (not a test): per-parameter gradient error of the line branch vs oracle autograd"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import oracle, synth_weights
import test_train_gpu as T

net, criterion, train, c5, targets = T.setup()
lb = train.LineBranch(synth_weights(), net.cfg)
logits, lines = lb.forward(c5)
cfg = dict(oracle.DEFAULT_CFG)
sd = {k: v.clone().float() for k, v in synth_weights().items() if v.is_floating_point()}
if os.environ.get("BF16W"):
    sd = {k: v.bfloat16().float() for k, v in sd.items()}
names = list(lb.index)
for k in names:
    sd[k].requires_grad_(True)
c5_ref = c5.float().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
lo, li = T.oracle_branch(sd, c5_ref, cfg)
total, losses, dc5 = lb.loss_and_grads(c5, targets, criterion)
tl = [t["lines"].cpu() for t in targets]
num_items = max(float(sum(len(t) for t in tl)), 1.0)
ref_total = 0.0
for s in range(lo.shape[0]):
    idx = criterion.matcher({"pred_logits": logits[s], "pred_lines": lines[s]}, targets)
    ce, l1 = oracle.set_losses(lo[s], li[s], tl, idx, num_items, 0.1)
    ref_total = ref_total + ce + 5.0 * l1
dlogits, dlines = lb.last_cotangents
ref_grads = torch.autograd.grad([lo, li], [sd[k] for k in names] + [c5_ref], [dlogits.cpu(), dlines.cpu()])
got = lb.grads()
for k, gr in zip(names, ref_grads):
    gg = got[k].double().cpu().reshape(-1); gr = gr.double().reshape(-1)
    print("%-60s |g| %.3e rel %.4f cos %.5f" % (k, gr.norm(), (gg - gr).norm() / gr.norm(), gg @ gr / (gg.norm() * gr.norm())))
