"""CPU, build container only: run the UNMODIFIED reference in-process next to the oracle on a padded
(ragged) batch -- a case the committed fixtures do not cover.  Skipped where /root/reference is absent."""
import pytest
import torch

from helpers import oracle, rel_err, synth, synth_weights, load_spec

ref_shims = pytest.importorskip("ref_shims")
pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not mounted")


def test_ragged_batch_with_padding_mask():
    model, criterions, post, args = ref_shims.build_reference()
    model.eval()
    sd = synth_weights()
    model.load_state_dict({k: v for k, v in sd.items() if not k.endswith("relative_position_index")}, strict=False)
    g = torch.Generator().manual_seed(3)
    a = torch.randn(3, 192, 256, generator=g)
    b = torch.randn(3, 160, 224, generator=g)
    with torch.no_grad():
        ref = model([a, b])
    images = torch.zeros(2, 3, 192, 256)
    mask = torch.ones(2, 192, 256, dtype=torch.bool)
    images[0] = a
    mask[0] = False
    images[1, :, :160, :224] = b
    mask[1, :160, :224] = False
    out = oracle.forward(sd, images, mask)
    assert rel_err(out["pred_logits"], ref["pred_logits"]) < 2e-5
    assert rel_err(out["pred_lines"], ref["pred_lines"]) < 2e-5
    for x, y in zip(out["pred_depth"], ref["pred_depth"]):
        assert rel_err(x, y) < 1e-4
    assert rel_err(out["pred_seg"], ref["pred_seg"]) < 1e-4


def test_spec_matches_reference_build():
    model, _, _, _ = ref_shims.build_reference()
    spec = load_spec()
    assert [k for k, _, _ in spec["keys"]] == list(model.state_dict().keys())
    assert spec["trainable"] == sorted(n for n, p in model.named_parameters() if p.requires_grad)
