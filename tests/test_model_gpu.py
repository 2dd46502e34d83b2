"""GPU parity of the whole forward through the public API (build_model / model(samples)) against the CPU oracle and the
fixtures the reference produced, plus size-independent properties at the benchmark size.

Tolerances: activations are bf16 (8-bit mantissa) through ~150 layers, so the comparison against the fp32 oracle is
bounded by bf16 round-off, not by kernel bugs; `test_error_is_bf16_roundoff` shows the same error level when the ORACLE
itself is run with bf16-rounded weights/activations.  Discrete selections (top-20 lines, uncertainty samples) are
pinned to the oracle's for the tight comparison and reported un-pinned separately."""
import pytest
import torch
import torch.nn.functional as F

from helpers import golden, oracle, synth, synth_weights

pytestmark = pytest.mark.gpu

# mean / max relative error bounds vs the fp32 oracle (bf16 path, selections pinned).  Measured on the B200 at 2x224x320,
# 1x480x640, 16x480x640 and 1x960x1280: depth mean 1.2-1.4 % / max 10-13 %, logits max 1.6-2.5 %, end points 2.6-4.6 %, seg mean
# 7.7-8.2 % (logits around zero).  north_star's 1e-2 on the depth mean is NOT met: rounding the WEIGHTS alone to bf16 moves the
# oracle's own depth by 1.34 % (tools/bf16_sensitivity.py), so the bound sits at 1.5 x that distance.
TOL = {"depth_mean": 2e-2, "depth_max": 0.16, "logits_max": 4e-2, "lines_max": 8e-2, "seg_mean": 0.12}


def _model():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    net, criterions, post = M.build_model(M.default_args(device="cuda"))
    net.load_state_dict(synth_weights())
    return net.cuda().eval(), criterions, M


_cache = {}


def model():
    if "m" not in _cache:
        _cache["m"] = _model()
    return _cache["m"]


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    return float(d.mean() / b.abs().mean().clamp_min(1e-9)), float(d.max() / b.abs().max().clamp_min(1e-9))


def pinned_from(trace):
    return {"line_ids": trace["line_ids"].cuda(), "sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}


@pytest.mark.parametrize("B,H,W,fixture", [(2, 224, 320, "fwd_224x320_b2.npz"), (1, 480, 640, "fwd_480x640_b1.npz")])
def test_forward_matches_oracle_and_reference_fixture(B, H, W, fixture):
    net, _, _ = model()
    images, _, _, _ = synth.synth_batch(B, H, W, seed=0)
    trace = {}
    ref = oracle.forward(synth_weights(), images, trace=trace)
    with torch.no_grad():
        out = net(images.cuda(), _pinned=pinned_from(trace))
    assert set(out) == {"pred_logits", "pred_lines", "aux_outputs", "pred_depth", "pred_seg"}
    assert len(out["aux_outputs"]) == 5 and len(out["pred_depth"]) == 4
    for a, b in zip(out["pred_depth"], ref["pred_depth"]):
        assert a.shape == b.shape
    assert out["pred_seg"].shape == ref["pred_seg"].shape
    assert rel(out["pred_logits"], ref["pred_logits"])[1] < TOL["logits_max"]
    assert rel(out["pred_lines"], ref["pred_lines"])[1] < TOL["lines_max"]
    for i in range(4):
        m, x = rel(out["pred_depth"][i], ref["pred_depth"][i])
        assert m < (TOL["depth_mean"] if i == 3 else 6e-2) and x < 0.3, (i, m, x)
    m, x = rel(out["pred_depth"][3], ref["pred_depth"][3])
    assert m < TOL["depth_mean"] and x < TOL["depth_max"], (m, x)
    assert rel(out["pred_seg"], ref["pred_seg"])[0] < TOL["seg_mean"]
    # the same against what the UNMODIFIED reference produced (committed fixture)
    g = golden(fixture)
    s = int(g["dense_stride"])
    m, x = rel(out["pred_depth"][3][..., ::s, ::s], torch.from_numpy(g["pred_depth_3"]))
    assert m < TOL["depth_mean"] and x < TOL["depth_max"], (m, x)
    assert rel(out["pred_logits"], torch.from_numpy(g["pred_logits"]))[1] < TOL["logits_max"]


def test_ragged_batch_with_padding_mask():
    """images of different sizes in one batch (src/util/misc.py:291-313 pads them, the mask reaches the position codes of
    all four levels and the key-padding masks of the DETR attention): public API with a LIST of images vs the oracle,
    which tests/test_oracle_vs_reference.py checks against the unmodified reference on exactly this kind of batch"""
    net, _, M = model()
    images, _, _, _ = synth.synth_batch(2, 224, 320, seed=1)
    a, b = images[0], images[1][:, :160, :256].contiguous()
    padded = torch.zeros(2, 3, 224, 320)
    mask = torch.ones(2, 224, 320, dtype=torch.bool)
    padded[0], mask[0] = a, False
    padded[1, :, :160, :256] = b
    mask[1, :160, :256] = False
    trace = {}
    ref = oracle.forward(synth_weights(), padded, mask, trace=trace)
    nt = M.nested_tensor_from_tensor_list([a.cuda(), b.cuda()])
    assert nt.padded is True and torch.equal(nt.mask.cpu(), mask)
    with torch.no_grad():
        out = net(nt, _pinned=pinned_from(trace))
        plain = net(padded.cuda(), _pinned=pinned_from(trace))        # same pixels, no mask
    assert rel(out["pred_logits"], ref["pred_logits"])[1] < TOL["logits_max"]
    assert rel(out["pred_lines"], ref["pred_lines"])[1] < TOL["lines_max"]
    m, x = rel(out["pred_depth"][3], ref["pred_depth"][3])
    assert m < TOL["depth_mean"] and x < TOL["depth_max"], (m, x)
    assert rel(out["pred_seg"], ref["pred_seg"])[0] < TOL["seg_mean"]
    # the mask must matter for the padded image and must not for the full-size one's logits beyond round-off
    d_masked = rel(out["pred_logits"][1], ref["pred_logits"][1])[1]
    d_plain = rel(plain["pred_logits"][1], ref["pred_logits"][1])[1]
    assert d_plain > 3 * d_masked, (d_plain, d_masked)


def test_padded_batch_graph_replay_and_output_ownership():
    """a padded batch replays a CUDA graph too (the padding mask is a static input of the graph) and equals the kernel-by-kernel
    launch; the public forward hands out tensors the next call does not overwrite"""
    net, _, M = model()
    images, _, _, _ = synth.synth_batch(2, 224, 320, seed=7)
    a, b = images[0], images[1][:, :192, :288].contiguous()
    nt = M.nested_tensor_from_tensor_list([a.cuda(), b.cuda()])
    nt2 = M.nested_tensor_from_tensor_list([images[1].cuda(), images[0][:, :160, :320].contiguous().cuda()])
    with torch.no_grad():
        eager = net(nt, _trace={})                 # a trace request takes the un-graphed path
        g1 = net(nt)
        keep = g1["pred_depth"][3].clone()
        g2 = net(nt2)                              # same shape, other images AND another mask: replays the same graph
        e2 = net(nt2, _trace={})
    assert torch.equal(g1["pred_depth"][3], keep), "the second call overwrote the first call's outputs"
    for k in ("pred_logits", "pred_lines", "pred_seg"):
        assert torch.allclose(g1[k].float(), eager[k].float(), rtol=2e-3, atol=2e-4), k
        assert torch.allclose(g2[k].float(), e2[k].float(), rtol=2e-3, atol=2e-4), k
    assert torch.allclose(g1["pred_depth"][3], eager["pred_depth"][3], rtol=2e-3, atol=2e-4)
    assert torch.allclose(g2["pred_depth"][3], e2["pred_depth"][3], rtol=2e-3, atol=2e-4)
    assert not torch.allclose(g1["pred_depth"][3], g2["pred_depth"][3], rtol=1e-2, atol=1e-2)


def test_dense_center_configuration():
    """--with_dense_center (BASELINE configs[3], the richest flag set that runs in the reference): three points
    per selected line -> 60 reference tokens in the line-window attention instead of 40"""
    _, _, M = model()
    net, _, _ = M.build_model(M.default_args(device="cuda", with_dense_center=True))
    net.load_state_dict(synth_weights())
    net.cuda().eval()
    images, _, _, _ = synth.synth_batch(2, 224, 320, seed=2)
    trace = {}
    ref = oracle.forward(synth_weights(), images, cfg={"with_dense_center": True}, trace=trace)
    base = oracle.forward(synth_weights(), images, cfg={"with_dense_center": False})
    with torch.no_grad():
        out = net(images.cuda(), _pinned=pinned_from(trace))
    m, x = rel(out["pred_depth"][3], ref["pred_depth"][3])
    assert m < TOL["depth_mean"] and x < TOL["depth_max"], (m, x)
    assert rel(out["pred_seg"], ref["pred_seg"])[0] < TOL["seg_mean"]
    assert rel(out["pred_depth"][0], ref["pred_depth"][0])[0] < 6e-2
    # the centre points must matter: the two configurations differ by more than the bf16 distance to the right one
    assert rel(base["pred_depth"][3], ref["pred_depth"][3])[0] > 1.3 * m


def test_forward_matches_oracle_second_weight_draw():
    """the same parity on a SECOND random draw of all 970 tensors (more than one weight set), through
    load_state_dict on the live module (the cached kernel plan must notice the new weights)"""
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    sd = synth_weights(seed=1)
    net, _, _ = M.build_model(M.default_args(device="cuda"))
    net.load_state_dict(synth_weights())
    net.cuda().eval()
    B, H, W = 2, 224, 320
    images, _, _, _ = synth.synth_batch(B, H, W, seed=4)
    with torch.no_grad():
        first = net(images.cuda())["pred_depth"][3].clone()       # plan built on the first weight set
    net.load_state_dict(sd)
    trace = {}
    ref = oracle.forward(sd, images, trace=trace)
    with torch.no_grad():
        out = net(images.cuda(), _pinned=pinned_from(trace))
    assert float((out["pred_depth"][3] - first).abs().max()) > 1e-3           # the new weights are in use
    assert rel(out["pred_logits"], ref["pred_logits"])[1] < TOL["logits_max"]
    assert rel(out["pred_lines"], ref["pred_lines"])[1] < TOL["lines_max"]
    m, x = rel(out["pred_depth"][3], ref["pred_depth"][3])
    assert m < TOL["depth_mean"] and x < 0.3, (m, x)
    assert rel(out["pred_seg"], ref["pred_seg"])[0] < TOL["seg_mean"]


def test_unpinned_selections_agree_with_oracle():
    """without pinning, the fp32-kept selection inputs must reproduce most of the oracle's choices"""
    net, _, _ = model()
    B, H, W = 2, 224, 320
    images, _, _, _ = synth.synth_batch(B, H, W, seed=0)
    trace, mine = {}, {}
    ref = oracle.forward(synth_weights(), images, trace=trace)
    with torch.no_grad():
        out = net(images.cuda(), _trace=mine)
    same = 0
    for b in range(B):
        same += len(set(mine["line_ids"][b].tolist()) & set(trace["line_ids"][b].tolist()))
    assert same >= 0.85 * B * 20, "top-20 line sets diverge: %d of %d" % (same, B * 20)
    m, x = rel(out["pred_depth"][3], ref["pred_depth"][3])
    assert m < 5e-2, (m, x)          # un-pinned: a flipped sample point moves the anchor mixture


def test_error_is_bf16_roundoff():
    """the CUDA path's distance to the fp32 oracle is the distance of a bf16-rounded run of the oracle itself"""
    net, _, _ = model()
    B, H, W = 1, 224, 320
    images, _, _, _ = synth.synth_batch(B, H, W, seed=0)
    sd = synth_weights()
    trace = {}
    ref = oracle.forward(sd, images, trace=trace)
    rnd = lambda t: t.bfloat16().float()  # noqa: E731
    sdb = {k: (rnd(v) if v.is_floating_point() and v.dim() > 1 else v) for k, v in sd.items()}
    names = ["linear", "conv2d", "layer_norm", "gelu", "relu", "elu"]
    orig = {n: getattr(F, n) for n in names}
    try:
        for n in names:
            setattr(F, n, (lambda f: (lambda *a, **k: rnd(f(*a, **k))))(orig[n]))
        pin = {"line_ids": trace["line_ids"], "sample1": (trace["sample1"], trace["sample1_idx"]),
               "sample2": (trace["sample2"], trace["sample2_idx"])}
        emu = oracle.forward(sdb, images, pinned=pin)
    finally:
        for n in names:
            setattr(F, n, orig[n])
    with torch.no_grad():
        out = net(images.cuda(), _pinned=pinned_from(trace))
    e_emu = rel(emu["pred_depth"][3], ref["pred_depth"][3])[0]
    e_cuda = rel(out["pred_depth"][3], ref["pred_depth"][3])[0]
    assert e_cuda < 2.0 * e_emu + 5e-3, (e_cuda, e_emu)


def _random_init_state_dict(source, seed=0):
    """random-init weights, as north_star words the parity bar: `package` = build_model()'s own initialisation (xavier / truncated
    normal, model._init_parameters), `reference` = the UNMODIFIED reference's build_model() under the same seed (staged copy under
    baseline/_ref or the mounted tree; skipped where neither exists)"""
    torch.manual_seed(seed)
    if source == "package":
        import gwdepth_b200  # noqa: F401
        from gwdepth_b200 import model as M
        net = M.build_model(M.default_args(device="cpu", dropout=0.0))[0]
        return {k: v.detach().clone() for k, v in net.state_dict().items()}, None
    import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("no copy of the reference on this box")
    ref_model = ref_shims.build_reference()[0].eval()
    return {k: v.detach().clone() for k, v in ref_model.state_dict().items()}, ref_model


# north_star: "outputs must match the reference PyTorch implementation on the same synthetic inputs and random-init weights: depth
# within 1e-2 relative in bf16".  With RANDOM-INIT weights (either initialiser) the bf16 path meets that bar; the hash-seeded
# high-gain weights of the other tests (TOL above) are the harder case and do not.  Stated bounds for the rest: every depth value
# within 6 % of its reference value, end points within 5e-3 absolute (they live in [0, 1]), logits within 8 % of the largest logit.
NORTH_STAR = {"depth_mean_rel": 1e-2, "depth_pointwise_max_rel": 6e-2, "lines_abs": 5e-3, "logits_max_rel": 8e-2}


@pytest.mark.parametrize("source,B,H,W", [("package", 2, 224, 320), ("package", 1, 480, 640), ("reference", 2, 224, 320)])
def test_random_init_weights_meet_the_north_star_depth_bar(source, B, H, W):
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    sd, ref_model = _random_init_state_dict(source)
    images, _, _, _ = synth.synth_batch(B, H, W, seed=0)
    trace = {}
    ref = oracle.forward(sd, images, trace=trace)
    if ref_model is not None:          # the oracle IS the reference on these weights: compare against the reference's own output
        with torch.no_grad():
            theirs = ref_model(images)
        assert rel(ref["pred_depth"][3], theirs["pred_depth"][3])[1] < 1e-5 and rel(ref["pred_logits"], theirs["pred_logits"])[1] < 1e-4
        ref = {k: theirs[k] for k in ("pred_logits", "pred_lines", "pred_depth", "pred_seg")}
    net = M.build_model(M.default_args(device="cuda", dropout=0.0))[0]
    net.load_state_dict(sd)
    net.cuda().eval()
    with torch.no_grad():
        out = net(images.cuda(), _pinned=pinned_from(trace))
    for i in range(4):
        a, b = out["pred_depth"][i].float().cpu().reshape(-1), ref["pred_depth"][i].float().reshape(-1)
        assert torch.isfinite(a).all()
        mean_rel = float((a - b).abs().mean() / b.abs().mean())
        point_rel = float(((a - b).abs() / b.abs().clamp_min(1e-3)).max())
        print("random-init (%s) %dx%dx%d depth level %d: mean-rel %.5f, largest point-wise rel %.5f" % (source, B, H, W, i, mean_rel, point_rel))
        assert mean_rel < NORTH_STAR["depth_mean_rel"] and point_rel < NORTH_STAR["depth_pointwise_max_rel"], (i, mean_rel, point_rel)
    lines_abs = float((out["pred_lines"].float().cpu() - ref["pred_lines"]).abs().max())
    logits_rel = rel(out["pred_logits"], ref["pred_logits"])[1]
    print("random-init (%s) %dx%dx%d end points max abs %.5f, logits max-rel %.5f" % (source, B, H, W, lines_abs, logits_rel))
    assert lines_abs < NORTH_STAR["lines_abs"] and logits_rel < NORTH_STAR["logits_max_rel"], (lines_abs, logits_rel)


def test_criterion_on_device_outputs():
    """SetCriterion with the CUDA cost-matrix kernel vs the oracle criterion on the SAME predictions"""
    net, criterions, _ = model()
    B, H, W = 2, 224, 320
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=0)
    with torch.no_grad():
        out = net(images.cuda())
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    losses = criterions[0](out, tg)
    cpu_out = {"pred_logits": out["pred_logits"].cpu(), "pred_lines": out["pred_lines"].cpu(),
               "aux_outputs": [{k: v.cpu() for k, v in a.items()} for a in out["aux_outputs"]]}
    ref_losses, _ = oracle.set_criterion(cpu_out, [t["lines"] for t in targets])
    assert set(losses) == set(ref_losses)
    for k in ref_losses:
        assert abs(float(losses[k]) - float(ref_losses[k])) <= 2e-4 * max(1.0, abs(float(ref_losses[k]))), k
    # dense losses through the reference-shaped criteria (engine_glassrgbd.py:65-90)
    mask = (depth_gt >= 0.2) & (depth_gt < 10.0)
    ref_d = oracle.depth_losses([d.cpu() for d in out["pred_depth"]], depth_gt)
    for i, pd in enumerate(out["pred_depth"]):
        size = pd.shape[-2:]
        g = F.interpolate(depth_gt, size=size, mode="nearest").cuda()
        m = F.interpolate(mask.to(torch.uint8), size=size, mode="nearest").to(torch.bool).cuda()
        w = (0.25, 0.25, 0.25, 1.0)[i]
        got = float(criterions[1](pd, g, m)) * w
        assert abs(got - float(ref_d[i])) <= 1e-3 * max(1.0, abs(float(ref_d[i]))), (i, got, float(ref_d[i]))
    got_s = float(criterions[2](out["pred_seg"], seg_gt.squeeze(1).cuda())) * 2.0
    assert abs(got_s - float(oracle.seg_loss(out["pred_seg"].cpu(), seg_gt))) < 1e-3


def test_benchmark_size_properties():
    """BASELINE configs[1] size (16 x 480 x 640): finite, in range, and every image's result is independent of its
    batch neighbours (the property that makes batch sharding across GPUs exact)."""
    net, _, _ = model()
    B, H, W = 16, 480, 640
    images, _, _, _ = synth.synth_batch(B, H, W, seed=3)
    x = images.cuda()
    tr = {}
    with torch.no_grad():
        full = net(x, _trace=tr)
        # same discrete selections for the solo run: a 1-ulp difference in a near-tied line logit would otherwise
        # legitimately pick another reference line
        pin = {"line_ids": tr["line_ids"][5:6], "sample1": tr["sample1"][5:6], "sample2": tr["sample2"][5:6]}
        solo = net(x[5:6], _pinned=pin)
    d = full["pred_depth"][3]
    assert d.shape == (B, 1, H, W) and torch.isfinite(d).all() and float(d.min()) > 0 and float(d.max()) < 10.0
    assert full["pred_seg"].shape == (B, 2, H, W) and torch.isfinite(full["pred_seg"]).all()
    assert full["pred_logits"].shape == (B, 100, 2) and full["pred_lines"].shape == (B, 100, 6)
    assert float(full["pred_lines"].min()) >= 0 and float(full["pred_lines"].max()) <= 1
    # cuDNN chooses other algorithms for the backbone at batch 1 (different bf16 rounding of C2..C5) and the network
    # amplifies a 1-ulp feature difference ~3x, so the two runs agree to bf16 round-off, not bit-for-bit
    m, xm = rel(full["pred_depth"][3][5:6], solo["pred_depth"][3])
    assert m < TOL["depth_mean"], (m, xm)
    assert rel(full["pred_logits"][5:6], solo["pred_logits"])[1] < TOL["logits_max"]


@pytest.mark.parametrize("B,H,W,seed", [(16, 480, 640, 3), (1, 960, 1280, 4)])
def test_benchmark_and_large_sizes_match_oracle(B, H, W, seed):
    """oracle parity AT the benchmark configuration (BASELINE configs[1]: 16 x 480 x 640) and at the large eval size of config 5
    (960 x 1280, L = 1 200 tokens): every output of every image against the fp32 CPU oracle, selections pinned"""
    net, _, _ = model()
    images, _, _, _ = synth.synth_batch(B, H, W, seed=seed)
    trace = {}
    ref = oracle.forward(synth_weights(), images, trace=trace)
    with torch.no_grad():
        out = net(images.cuda(), _pinned=pinned_from(trace))
    per_image = [rel(out["pred_depth"][3][b], ref["pred_depth"][3][b]) for b in range(B)]
    m, x = rel(out["pred_depth"][3], ref["pred_depth"][3])
    print("depth %dx%dx%d: mean-rel %.4f max-rel %.4f (worst image mean %.4f); logits max-rel %.4f, lines %.4f, seg mean-rel %.4f" % (
        B, H, W, m, x, max(p[0] for p in per_image), rel(out["pred_logits"], ref["pred_logits"])[1],
        rel(out["pred_lines"], ref["pred_lines"])[1], rel(out["pred_seg"], ref["pred_seg"])[0]))
    assert m < TOL["depth_mean"] and x < TOL["depth_max"], (m, x)
    assert max(p[0] for p in per_image) < 1.5 * TOL["depth_mean"]
    assert rel(out["pred_logits"], ref["pred_logits"])[1] < TOL["logits_max"]
    assert rel(out["pred_lines"], ref["pred_lines"])[1] < TOL["lines_max"]
    assert rel(out["pred_seg"], ref["pred_seg"])[0] < TOL["seg_mean"]
    for i in range(3):
        assert rel(out["pred_depth"][i], ref["pred_depth"][i])[0] < 6e-2, i


def test_infer_stream_matches_forward():
    """the pipelined serving loop (three streams, double buffers) returns exactly what model(x) returns, batch by batch"""
    net, _, _ = model()
    batches = [synth.synth_batch(2, 224, 320, seed=300 + i)[0].pin_memory() for i in range(5)]
    want = []
    with torch.no_grad():
        for hb in batches:
            out = net(hb.cuda())
            want.append({"pred_logits": out["pred_logits"].clone().cpu(), "pred_lines": out["pred_lines"].clone().cpu(),
                         "pred_depth": out["pred_depth"][-1].clone().cpu(), "pred_seg": out["pred_seg"].clone().cpu()})
    got = [{k: v.clone() for k, v in res.items()} for res in net.infer_stream(iter(batches))]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in w:
            # same kernels, same inputs; only the fp64 atomics of the diffusion statistics may reorder
            assert torch.allclose(g[k].float(), w[k].float(), rtol=2e-3, atol=2e-4), k


def test_infer_stream_accepts_raw_uint8_images():
    """uint8 [B,H,W,3] host batches through the serving loop == normalising on the host (ToTensor + Normalize) first"""
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    net, _, _ = model()
    g = torch.Generator().manual_seed(9)
    mean, std = torch.tensor(ops.IMAGE_MEAN), torch.tensor(ops.IMAGE_STD)
    raws = [torch.randint(0, 256, (2, 128, 160, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(3)]
    floats = [r.permute(0, 3, 1, 2).float().div(255).sub(mean[None, :, None, None]).div(std[None, :, None, None]).contiguous().pin_memory()
              for r in raws]
    a = [{k: v.clone() for k, v in out.items()} for out in net.infer_stream(iter(raws))]
    b = [{k: v.clone() for k, v in out.items()} for out in net.infer_stream(iter(floats))]
    for x, y in zip(a, b):
        for k in x:
            assert torch.equal(x[k], y[k]), k
