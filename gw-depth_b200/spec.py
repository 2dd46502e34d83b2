"""Parameter inventory of the GW-Depth model, generated from hyper-parameters.

The drop-in must load and save the reference's checkpoints unchanged ("checkpoint / resume"),
so the model exposes exactly the reference's state_dict: the same dotted names, shapes, dtypes and order, the same
split between parameters and buffers, and the same frozen (requires_grad=False) backbone stem + layer1.  This
module enumerates that inventory; `tests/test_boundary.py` checks it against the list recorded from the reference
(tests/golden/state_dict_spec.json).

Entries are (name, shape, dtype, kind, trainable) with kind in {"param", "buffer"}.
"""
F32, I64 = "float32", "int64"


def _lin(out, name, n_out, n_in, bias=True):
    out.append((name + ".weight", (n_out, n_in), F32, "param", True))
    if bias:
        out.append((name + ".bias", (n_out,), F32, "param", True))


def _norm(out, name, n):
    out.append((name + ".weight", (n,), F32, "param", True))
    out.append((name + ".bias", (n,), F32, "param", True))


def _conv(out, name, n_out, n_in, k, bias=False, trainable=True):
    out.append((name + ".weight", (n_out, n_in, k, k), F32, "param", trainable))
    if bias:
        out.append((name + ".bias", (n_out,), F32, "param", trainable))


def _frozen_bn(out, name, n):  # src/models/backbone.py:19-33: four buffers, no parameters
    for leaf in ("weight", "bias", "running_mean", "running_var"):
        out.append(("%s.%s" % (name, leaf), (n,), F32, "buffer", False))


def _mha(out, name, e):  # src/models/multi_head_attention.py:382-470 (packed in-projection)
    out.append((name + ".in_proj_weight", (3 * e, e), F32, "param", True))
    out.append((name + ".in_proj_bias", (3 * e,), F32, "param", True))
    _lin(out, name + ".out_proj", e, e)


def detr_spec(out, cfg):
    e, ff = cfg["hidden_dim"], cfg["dim_feedforward"]
    for i in range(cfg["enc_layers"]):
        p = "transformer.encoder.layers.%d" % i
        _mha(out, p + ".self_attn", e)
        _lin(out, p + ".linear1", ff, e)
        _lin(out, p + ".linear2", e, ff)
        _norm(out, p + ".norm1", e)
        _norm(out, p + ".norm2", e)
    for i in range(cfg["dec_layers"]):
        p = "transformer.decoder.layers.%d" % i
        _mha(out, p + ".self_attn", e)
        _mha(out, p + ".multihead_attn", e)
        _lin(out, p + ".linear1", ff, e)
        _lin(out, p + ".linear2", e, ff)
        for n in ("norm1", "norm2", "norm3"):
            _norm(out, "%s.%s" % (p, n), e)
    _norm(out, "transformer.decoder.norm", e)
    _lin(out, "class_embed", 2, e)
    out.append(("query_embed.weight", (cfg["num_queries"], e), F32, "param", True))
    _conv(out, "input_proj", e, 2048, 1, bias=True)
    line_dim = 6 if cfg["with_center"] else 4
    for i, n_out in enumerate((e, e, line_dim)):
        _lin(out, "lines_embed.layers.%d" % i, n_out, e)


def resnet50_spec(out):
    """torchvision resnet50 under IntermediateLayerGetter (src/models/backbone.py:58-92); stem and layer1 are frozen"""
    p = "backbone.0.body."
    _conv(out, p + "conv1", 64, 3, 7, trainable=False)
    _frozen_bn(out, p + "bn1", 64)
    inplanes = 64
    for li, (planes, nblocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        train = li > 1
        for b in range(nblocks):
            q = "%slayer%d.%d." % (p, li, b)
            _conv(out, q + "conv1", planes, inplanes, 1, trainable=train)
            _frozen_bn(out, q + "bn1", planes)
            _conv(out, q + "conv2", planes, planes, 3, trainable=train)
            _frozen_bn(out, q + "bn2", planes)
            _conv(out, q + "conv3", planes * 4, planes, 1, trainable=train)
            _frozen_bn(out, q + "bn3", planes * 4)
            if b == 0:
                _conv(out, q + "downsample.0", planes * 4, inplanes, 1, trainable=train)
                _frozen_bn(out, q + "downsample.1", planes * 4)
            inplanes = planes * 4


def _mlp(out, name, d, hidden):
    _lin(out, name + ".fc1", hidden, d)
    _lin(out, name + ".fc2", d, hidden)


def _window_attn_common(out, p, dim, heads, ws, cls):
    out.append((p + ".diff_mu", (1, 1, dim), F32, "param", True))
    out.append((p + ".diff_logsigma", (1, 1, dim), F32, "param", True))
    if cls:
        out.append((p + ".border_mu", (1, 1, dim), F32, "param", True))
        out.append((p + ".border_logsigma", (1, 1, dim), F32, "param", True))
    out.append((p + ".relative_position_bias_table", ((2 * ws - 1) ** 2, heads), F32, "param", True))
    out.append((p + ".relative_position_index", (ws * ws, ws * ws), I64, "buffer", False))
    _lin(out, p + ".qkv", 3 * dim, dim)
    _lin(out, p + ".proj", dim, dim)


def _line_block(out, p, dim, heads, ws):  # SwinTransformerBlock + WindowAttention, multiscale_transformerr.py:202-265,601-644
    _norm(out, p + ".norm1", dim)
    _window_attn_common(out, p + ".attn", dim, heads, ws, False)
    _lin(out, p + ".attn.ref_qk", 2 * dim, dim)
    _conv(out, p + ".attn.ref_attn_diffusion", heads, heads, 3, bias=True)
    _norm(out, p + ".norm2", dim)
    _mlp(out, p + ".mlp", dim, 2 * dim)


def _class_block(out, p, dim, heads, ws, td):  # ... + WindowClassAttention, multiscale_transformerr.py:375-452,624-632
    _norm(out, p + ".norm1", dim)
    _window_attn_common(out, p + ".attn", dim, heads, ws, True)
    _lin(out, p + ".attn.cls_dth_q", td, td)
    _lin(out, p + ".attn.cls_seg_q", td, td)
    _lin(out, p + ".attn.global_k", dim + 2 * td, dim + 2 * td)
    _lin(out, p + ".attn.global_v", dim + 2 * td, dim + 2 * td)
    _lin(out, p + ".attn.proj_dth", td, td)
    _lin(out, p + ".attn.proj_seg", td, td)
    _norm(out, p + ".norm2", dim)
    _mlp(out, p + ".mlp", dim, 2 * dim)
    _norm(out, p + ".norm_seg1", td)
    _norm(out, p + ".norm_depth1", td)
    _mlp(out, p + ".mlp_seg", td, 2 * td)
    _norm(out, p + ".norm_seg2", td)
    _mlp(out, p + ".mlp_depth", td, 2 * td)
    _norm(out, p + ".norm_depth2", td)


def _conv_ln(out, name, n_out, n_in, k=3):
    _conv(out, name + ".conv", n_out, n_in, k)
    _norm(out, name + ".layer_norm", n_out)


def _pyramid(out, p, k):  # PyramidLayer, src/models/points/points_sample.py:45-92
    _conv_ln(out, p + ".firstconv.0", k, k)
    _conv_ln(out, p + ".firstconv.2", 2 * k, k)
    for lname, nblk in (("layer1", 1), ("layer2", 2), ("layer3", 2), ("layer4", 1)):
        for b in range(nblk):
            _conv_ln(out, "%s.%s.%d.conv1.0" % (p, lname, b), 2 * k, 2 * k)
            _conv_ln(out, "%s.%s.%d.conv2" % (p, lname, b), 2 * k, 2 * k)
    for i in range(1, 5):
        _conv_ln(out, "%s.branch%d.1" % (p, i), 2 * k, 2 * k)
    _conv_ln(out, p + ".lastconv.0", 4 * k, 10 * k)
    _conv(out, p + ".lastconv.2", k, 4 * k, 1)


def _depth_head(out, name, n_in, td):
    _lin(out, name + ".0", td, n_in)
    _lin(out, name + ".1", 1, td)


def dense_encoder_spec(out, cfg):  # ReferTransformer.__init__, multiscale_transformerr.py:1025-1138
    p = "dense_encoder."
    D, heads, ws, td = cfg["dense_trans_dim"], cfg["dense_trans_heads"], cfg["window"], cfg["class_token_dim"]
    out.append((p + "depth_token", (1, 1, td), F32, "param", True))
    out.append((p + "seg_token", (1, 1, td), F32, "param", True))
    for i in range(cfg["dense_trans_layers"][0]):
        _line_block(out, "%sdense_transformer.blocks.%d" % (p, i), D, heads, ws)
    _depth_head(out, p + "depth_pred32", D, td)
    backbone_ch = {1: 1024, 2: 512, 3: 256}
    for si, depth in enumerate(cfg["class_trans_layers"], start=1):
        C = D >> si
        if si > 1:
            scale_name = {2: "8", 3: "4"}[si]
            for kind in ("depth", "seg"):
                q = "%sold_%s_token_proj%s" % (p, kind, scale_name)
                _mlp(out, q, td, 2 * td)
                _norm(out, q + ".norm", td)
        _lin(out, "%sproj_class%d" % (p, si), C, 2 * C)
        _conv(out, "%sproj_backbn%d.conv" % (p, si), C, backbone_ch[si], 3, bias=True)
        for i in range(depth):
            _class_block(out, "%sclass_transformer%d.blocks.%d" % (p, si, i), C, heads, ws, td)
        if si == 1:
            _depth_head(out, p + "class_transformer1.pre_depth_pred", C + td, td)
            _depth_head(out, p + "depth_pred16", C + td, td)
        if si < 3:
            Cn, K = D >> (si + 1), cfg["interval_sample_num"][si - 1]
            q = "%spoint_based_pred%d" % (p, si)
            _lin(out, q + ".pre_proj", Cn, Cn + td)
            _lin(out, q + ".refer_proj", 2 * Cn, Cn)
            _pyramid(out, q + ".pyramid", K)
    _depth_head(out, p + "depth_pred4", (D >> 3) + td, td)


def depth_decoder_spec(out, cfg):  # DensePrediction.__init__, src/models/dense_upsample.py:114-147
    p, td, C = "depth_decoder.", cfg["class_token_dim"], 64
    _lin(out, p + "depth_token_fuse.fc1", C + 1 + td, C + 1 + td)
    _lin(out, p + "depth_token_fuse.fc2", td, C + 1 + td)
    _lin(out, p + "seg_token_fuse.fc1", C + td, C + td)
    _lin(out, p + "seg_token_fuse.fc2", td, C + td)
    for kind, n_out in (("depth", 1), ("seg", 2)):
        _conv(out, "%supconv1_%s.conv" % (p, kind), td, td, 3)
        _norm(out, p + "norm_" + kind, td)
        _conv(out, "%sconv1_%s.0" % (p, kind), td, td, 3)
        _conv(out, "%supconv2_%s.conv" % (p, kind), td // 2, td, 3)
        _conv(out, "%sconv2_%s.0" % (p, kind), td // 2, td // 2, 3)
        _conv(out, p + ("get_depth.0" if kind == "depth" else "get_seg"), n_out, td // 2, 3)


DEFAULT_HP = dict(hidden_dim=256, dim_feedforward=2048, enc_layers=6, dec_layers=6, num_queries=100, with_center=True,
                  dense_trans_dim=512, dense_trans_heads=16, dense_trans_layers=(4,), class_trans_layers=(2, 2, 1),
                  class_token_dim=64, window=7, interval_sample_num=(30, 80, 160))


def model_spec(cfg=None):
    """the full inventory in the reference's registration order (src/models/glassrgbd.py:45-72)"""
    cfg = dict(DEFAULT_HP, **(cfg or {}))
    out = []
    detr_spec(out, cfg)
    resnet50_spec(out)
    _conv(out, "dense_input_proj", 512, 2048, 1, bias=True)
    dense_encoder_spec(out, cfg)
    depth_decoder_spec(out, cfg)
    return out
