import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import oracle, synth, synth_weights
import gwdepth_b200
from gwdepth_b200.train_backbone import BODY, BackboneTrain
def rel_l2(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))
B, H, W = 1, 64, 96
sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith(BODY)}
images, _, _, _ = synth.synth_batch(B, H, W, seed=3)
for emulate in (False, True):
    sdr = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and k.endswith(".weight") and ("conv" in k or "downsample.0" in k)
                                                   and any(("layer%d." % i) in k for i in (2, 3, 4))) else v) for k, v in sd.items()}
    if emulate:   # the oracle with bf16-rounded activations after every bottleneck conv (straight-through): the yard-stick
        import torch.nn.functional as F
        q = lambda t: t + (t.bfloat16().float() - t).detach()
        def bottleneck(x, p, stride):
            out = q(F.relu(oracle.frozen_bn(F.conv2d(x, p["conv1.weight"]), p, "bn1")))
            out = q(F.relu(oracle.frozen_bn(F.conv2d(out, p["conv2.weight"], stride=stride, padding=1), p, "bn2")))
            out = oracle.frozen_bn(F.conv2d(out, p["conv3.weight"]), p, "bn3")
            if p.has("downsample.0.weight"):
                x = q(oracle.frozen_bn(F.conv2d(x, p["downsample.0.weight"], stride=stride), p, "downsample.1"))
            return q(F.relu(out + x))
        oracle.bottleneck = bottleneck
    feats = oracle.resnet50_features(images, oracle.P(sdr, BODY))
    g = torch.Generator().manual_seed(B)
    cots = [torch.randn(f.shape, generator=g) * (f > 0) for f in feats[1:]]
    sum((f * c).sum() for f, c in zip(feats[1:], cots)).backward()
    if not emulate:
        ref = {k: v.grad.clone() for k, v in sdr.items() if isinstance(v, torch.Tensor) and v.requires_grad}
        bb = BackboneTrain({k: v.cuda() for k, v in sd.items()}, lr=1e-5)
        c2 = bb.frozen_front(images.cuda())
        outs = bb.forward(c2)
        bb.backward(*[c.permute(0, 2, 3, 1).contiguous().bfloat16().cuda() for c in cots])
        grads = bb.grads()
    else:
        for k in ref:
            print("%-55s cuda %.3f   bf16-emulated oracle %.3f" % (k[len(BODY):], rel_l2(grads[k], ref[k]), rel_l2(sdr[k].grad, ref[k])))
