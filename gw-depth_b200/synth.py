"""Deterministic synthetic weights / inputs of the reference's shapes: what `bench.py` feeds the model (there
is no dataset and no checkpoint in the container) and what the oracle, the golden-vector generator and the tests regenerate from
seeds (`oracle/synth.py` re-exports this module).  Pure data generation: no model code.

The reference's own random init cannot be reproduced on the GPU box (the reference is not there), so weights are generated per
state-dict key from a hash-seeded generator.  The scales are chosen so that activations stay O(1) through ~100 layers and the
discrete selections (top-20 line logits, top-K uncertainty pixels) are not near-tied, which is what makes bf16-vs-fp32 parity
meaningful.
"""
import hashlib

import torch


def _gen(key, seed):
    h = hashlib.sha256(("%d:%s" % (seed, key)).encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)


def synth_tensor(key, shape, dtype=torch.float32, seed=0):
    g = _gen(key, seed)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if dtype in (torch.int64, torch.int32):
        raise ValueError("integer buffers are structural, not synthetic: " + key)
    if leaf == "running_var":
        return torch.rand(shape, generator=g) * 0.5 + 0.75
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    is_norm = any(t in key for t in (".bn", "norm", "downsample.1")) and len(shape) == 1
    if is_norm and leaf == "weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "bias" or leaf == "in_proj_bias":
        return 0.05 * torch.randn(shape, generator=g)
    if leaf in ("diff_logsigma", "border_logsigma"):
        return 0.1 * torch.randn(shape, generator=g)
    if leaf in ("diff_mu", "border_mu"):
        return 0.5 * torch.randn(shape, generator=g)
    if leaf == "relative_position_bias_table":
        return 0.2 * torch.randn(shape, generator=g)
    if key.endswith("query_embed.weight"):
        return 2.0 * torch.randn(shape, generator=g)
    if leaf in ("depth_token", "seg_token"):
        return torch.randn(shape, generator=g)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        gain = 1.0
        if key.startswith("backbone") and ("conv3" in key):
            gain = 0.5          # residual branch: keep the 16-block ResNet from blowing up
        elif key.startswith("backbone"):
            gain = 1.4          # ReLU layers
        if key.startswith("class_embed"):
            gain = 3.0          # well separated line logits
        if "lines_embed.layers.2" in key:
            gain = 8.0          # end points spread over the unit square like the targets: L1 matching costs
                                # are piecewise linear, so clustered predictions give exactly tied assignments
        if "depth_pred" in key and key.endswith(".1.weight"):
            gain = 1.5          # spread the coarse depth maps over (0,1) without saturating the sigmoid
        if key.endswith("multihead_attn.out_proj.weight"):
            gain = 1.5          # the cross-attention read-out is what makes line queries differ
        elif "encoder.layers" in key and key.endswith("out_proj.weight"):
            gain = 0.2          # keep encoder tokens distinct (random self-attention is an averaging filter)
        elif any(key.endswith(t) for t in ("out_proj.weight", "attn.proj.weight", "linear2.weight", "fc2.weight")):
            gain = 0.5          # sub-layer outputs stay below the residual stream, as in a trained network;
                                # otherwise random attention averages every token / query onto one vector
        if key.endswith("lastconv.2.weight"):
            gain = 3.0          # a selective anchor mixture in PointBasedPred
        w = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        if any(t in key for t in ("input_proj.weight", "proj_backbn")):
            w = w * 0.5         # backbone maps have std of a few units; bring their projections to O(1)
        return w
    return 0.05 * torch.randn(shape, generator=g)


def _resnet50_maps(img, w):
    """plain torch ResNet-50 with frozen batch norm (torchvision layout of the keys, eps inside the rsqrt) -> the four stage outputs;
    used ONLY to estimate the channel means below (a data-generation detail, float64 on the CPU)"""
    import torch.nn.functional as F

    def bn(x, name):
        scale = w[name + ".weight"] * (w[name + ".running_var"] + 1e-5).rsqrt()
        shift = w[name + ".bias"] - w[name + ".running_mean"] * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    def block(x, pre, stride):
        out = F.relu(bn(F.conv2d(x, w[pre + "conv1.weight"]), pre + "bn1"))
        out = F.relu(bn(F.conv2d(out, w[pre + "conv2.weight"], stride=stride, padding=1), pre + "bn2"))
        out = bn(F.conv2d(out, w[pre + "conv3.weight"]), pre + "bn3")
        if pre + "downsample.0.weight" in w:
            x = bn(F.conv2d(x, w[pre + "downsample.0.weight"], stride=stride), pre + "downsample.1")
        return F.relu(out + x)
    x = F.relu(bn(F.conv2d(img, w["conv1.weight"], stride=2, padding=3), "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    maps = []
    for li, nblocks in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nblocks):
            x = block(x, "layer%d.%d." % (li, bi), 2 if (li > 1 and bi == 0) else 1)
        maps.append(x)
    return maps


def _center_backbone_consumers(sd):
    """Give the layers that read post-ReLU backbone maps a bias that cancels the maps' per-channel mean
    (estimated in float64 on a small fixed synthetic image, so it is reproducible across machines).  Without
    it every token is dominated by one shared vector and the 100 line queries collapse onto each other, which
    makes top-k selections and Hungarian assignments ill-conditioned -- unlike a trained network."""
    img, _, _, _ = synth_batch(1, 128, 160, seed=12345)
    sd64 = {k[len("backbone.0.body."):]: v.double() for k, v in sd.items() if k.startswith("backbone.0.body.")}
    with torch.no_grad():
        feats = _resnet50_maps(img.double(), sd64)
    mu = [f.mean(dim=(0, 2, 3)) for f in feats]          # C2, C3, C4, C5
    for key, level in (("input_proj", 3), ("dense_input_proj", 3), ("dense_encoder.proj_backbn1.conv", 2),
                       ("dense_encoder.proj_backbn2.conv", 1), ("dense_encoder.proj_backbn3.conv", 0)):
        w = sd[key + ".weight"].double()
        sd[key + ".bias"] = sd[key + ".bias"] - torch.einsum("oikl,i->o", w, mu[level]).float()
    return sd


def synth_state_dict(spec, seed=0):
    """spec: iterable of (key, shape, dtype-string).  Integer buffers (relative_position_index) are skipped:
    they are structural constants that every implementation builds itself."""
    sd = {}
    for key, shape, dt in spec:
        if dt in ("int64", "int32"):
            continue
        sd[key] = synth_tensor(key, shape, seed=seed)
    if "input_proj.weight" in sd and "backbone.0.body.conv1.weight" in sd:
        _center_backbone_consumers(sd)
    return sd


def relative_position_index(ws=7):
    """multiscale_transformerr.py:236-246 (structural buffer of every window-attention module)"""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def add_structural_buffers(sd, spec):
    for key, shape, dt in spec:
        if key.endswith("relative_position_index"):
            sd[key] = relative_position_index(7)
    return sd


def synth_batch(B, H, W, seed=0):
    """images, per-image line targets, depth gt, seg gt  """
    g = torch.Generator().manual_seed(1000 + seed)
    # multi-scale random field: white noise alone is statistically identical at every 1/32 cell, which makes
    # all transformer tokens (and then all line queries) collapse; real images have large-scale structure
    import torch.nn.functional as F
    images = 0.25 * torch.randn(B, 3, H, W, generator=g)
    for div, amp in ((64, 2.0), (32, 1.5), (8, 0.5)):
        coarse = torch.randn(B, 3, max(H // div, 1), max(W // div, 1), generator=g) * amp
        images = images + F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    targets = []
    for b in range(B):
        T = 12 + 5 * (b % 8)
        targets.append({"lines": torch.rand(T, 6, generator=g), "labels": torch.zeros(T, dtype=torch.int64)})
    depth_gt = torch.rand(B, 1, H, W, generator=g) * 9.5 + 0.3
    seg_gt = (torch.rand(B, 1, H, W, generator=g) > 0.5).long()
    return images, targets, depth_gt, seg_gt
