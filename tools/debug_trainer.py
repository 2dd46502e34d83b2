"""whole-model gradient check (debug): Trainer vs torch.autograd over the oracle, errors aggregated per module"""
import sys, os, time, collections, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import oracle, synth, synth_weights
import gwdepth_b200
from gwdepth_b200 import model as M
from gwdepth_b200.train_model import Trainer
def rel_l2(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 2, 128, 160
images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=0)
sd = synth_weights()
_, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
wd = crit[0].weight_dict
sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".bn" not in k and "downsample.1" not in k
           and not k.startswith(("backbone.0.body.conv1", "backbone.0.body.layer1")) else v) for k, v in sd.items()}
t0 = time.time()
trace = {}
ref = oracle.forward(sdr, images, trace=trace, grad=True)
tl = [t["lines"] for t in targets]
set_l, idx = oracle.set_criterion(ref, tl)
total = sum(v * wd[k] for k, v in set_l.items()) + sum(oracle.depth_losses(ref["pred_depth"], depth_gt)) + oracle.seg_loss(ref["pred_seg"], seg_gt)
total.backward()
print("oracle fwd+bwd %.1f s, total loss %.4f" % (time.time() - t0, float(total)))
tr = Trainer(sd, M.default_args().__dict__ and None)
pinned = {"line_ids": trace["line_ids"].cuda(), "sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}
tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
torch.cuda.synchronize(); t0 = time.time()
logits, lines, outs = tr.forward(images.cuda(), pinned)
g = tr.dense.loss_grads(outs, depth_gt.cuda(), seg_gt.cuda())
tr.backward_dense(*g)
stacked = idx[1:] + idx[:1]          # the oracle lists the final stage first, the stacked layout has it last
if os.environ.get("PIN_MATCH", "1") == "1":
    set_losses, dlogits, dlines = crit[0].cuda().forward_backward_stacked(logits, lines, tg, pinned_pairs=stacked)
else:
    set_losses, dlogits, dlines = crit[0].cuda().forward_backward_stacked(logits, lines, tg)
    mine = crit[0].matcher.pairs_from_raw()
    diff = sum(int(not (torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))) for sa, sb in zip(mine, stacked) for a, b in zip(sa, sb))
    print("assignment problems that differ from the oracle's: %d of %d" % (diff, len(stacked) * B))
tr.backward_line(dlogits, dlines)
torch.cuda.synchronize()
print("trainer fwd+bwd %.3f s" % (time.time() - t0))
tot = float(crit[0].last_total + tr.dense.losses().sum())
print("total loss cuda %.4f oracle %.4f; dense losses" % (tot, float(total)), tr.dense.losses().tolist())
grads = tr.grads()
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
missing, extra = [], []
for k, v in sdr.items():
    if not (isinstance(v, torch.Tensor) and v.requires_grad):
        continue
    if v.grad is None:
        if k in grads: extra.append(k)
        continue
    if k not in grads:
        missing.append(k); continue
    grp = ".".join(k.split(".")[:3]) if k.startswith(("backbone", "transformer")) else ".".join(k.split(".")[:2])
    d = (grads[k].double().cpu() - v.grad.double()); a = agg[grp]
    a[0] += float(d.pow(2).sum()); a[1] += float(v.grad.double().pow(2).sum()); a[2] += 1
for grp, (e, n, c) in sorted(agg.items()):
    print("%-50s %3d tensors  rel L2 %.3f  |ref| %.3e" % (grp, c, (e / max(n, 1e-300)) ** 0.5, n ** 0.5))
print("missing (oracle has a gradient, trainer does not):", missing[:20], len(missing))
print("extra (trainer stores, oracle has none):", extra[:20], len(extra))
print("trainable tensors with a gradient:", sum(a[2] for a in agg.values()))
