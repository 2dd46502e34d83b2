"""CPU: checkpoint tooling (gw-depth_b200/checkpoint.py) against the key surgery of src/main_glassrgbd.py:104-193 on synthetic
checkpoints: a DETR-R50-shaped file (class_embed [92,256], bbox_embed, query_embed dropped), a DataParallel-prefixed GW-Depth
file with the old `bbox_embed` naming, optimizer / scheduler restore, and a save -> load round trip."""
import torch

from helpers import synth_weights

import gwdepth_b200  # noqa: F401
from gwdepth_b200 import checkpoint as C, model as M


def _model():
    net, _, _ = M.build_model(M.default_args(device="cpu"))
    return net


def test_detr_pretrained_drops_heads_and_queries():
    sd = synth_weights()
    detr = {k: v.clone() for k, v in sd.items() if k.startswith(("transformer.", "input_proj.", "backbone."))}
    detr["class_embed.weight"], detr["class_embed.bias"] = torch.randn(92, 256), torch.randn(92)
    for i, (o, n) in enumerate(((256, 256), (256, 256), (4, 256))):
        detr["bbox_embed.layers.%d.weight" % i], detr["bbox_embed.layers.%d.bias" % i] = torch.randn(o, n), torch.randn(o)
    detr["query_embed.weight"] = torch.randn(100, 256)
    net = _model()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    rep = C.load_detr_pretrained(net, {"model": detr})
    after = net.state_dict()
    assert set(rep["loaded"]) == {k for k in detr if k.startswith(("transformer.", "input_proj.", "backbone."))}
    assert not rep["unexpected"] and not rep["shape_mismatch"]
    for k in ("class_embed.weight", "query_embed.weight", "lines_embed.layers.0.weight", "dense_input_proj.weight"):
        assert torch.equal(after[k], before[k]) and k in rep["missing"]          # untouched, reported as new parameters
    assert torch.equal(after["transformer.encoder.layers.0.linear1.weight"], sd["transformer.encoder.layers.0.linear1.weight"])
    # --layer1_num != 3 also skips input_proj (:112-113)
    net2 = _model()
    rep2 = C.load_detr_pretrained(net2, {"model": detr}, layer1_num=2)
    assert "input_proj.weight" not in rep2["loaded"]


def test_resume_strips_module_prefix_and_renames_bbox_embed():
    sd = synth_weights()
    old = {}
    for k, v in sd.items():
        if k.startswith("lines_embed."):
            old["bbox_embed." + k.split(".", 1)[1]] = v.clone()       # the reference renames by dropping the FIRST component
        else:
            old["module." + k] = v.clone()
    old["module.some_removed_head.weight"] = torch.zeros(3)
    net = _model()
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 200)
    opt2 = torch.optim.AdamW(params, lr=3e-5)
    ckpt = {"model": old, "optimizer": opt2.state_dict(), "lr_scheduler": torch.optim.lr_scheduler.StepLR(opt2, 50).state_dict(), "epoch": 7}
    rep, start = C.load_resume(net, ckpt, optimizer=opt, lr_scheduler=sch, lr_drop=120)
    assert start == 8 and sch.step_size == 120 and opt.param_groups[0]["lr"] == 3e-5
    assert rep["unexpected"] == ["some_removed_head.weight"] and not rep["missing"] and not rep["shape_mismatch"]
    got = net.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # --eval / --no_opt leave the optimizer alone (:159)
    _, start = C.load_resume(net, ckpt, optimizer=opt, lr_scheduler=sch, evaluate=True)
    assert start is None


def test_frozen_letr_and_save_round_trip(tmp_path):
    sd = synth_weights()
    net = _model()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    rep = C.load_frozen_letr(net, {"model": {"module." + k: v for k, v in sd.items()}})
    assert all(any(t in k for t in ("encoder", "decoder", "class_embed", "lines_embed")) for k in rep["loaded"])
    got = net.state_dict()
    assert torch.equal(got["transformer.decoder.norm.weight"], sd["transformer.decoder.norm.weight"])
    assert torch.equal(got["backbone.0.body.layer2.0.conv1.weight"], before["backbone.0.body.layer2.0.conv1.weight"])
    path = str(tmp_path / "checkpoint.pth")
    net.load_state_dict(sd)
    blob = C.save_checkpoint(path, net, epoch=3, args={"lr": 1e-4})
    assert set(blob["model"]) == set(sd) and blob["epoch"] == 3 and list(blob["model"]) == list(net.state_dict())
    net2 = _model()
    rep, _ = C.load_resume(net2, torch.load(path, weights_only=False))
    assert not rep["missing"] and not rep["unexpected"]
    assert all(torch.equal(net2.state_dict()[k], sd[k]) for k in sd)
