"""CPU: the drop-in boundary -- state_dict layout, build_model surface, C-ABI exports."""
import os
import re
import subprocess

import pytest
import torch

from helpers import ROOT, load_spec, synth_weights

import gwdepth_b200  # noqa: F401
from gwdepth_b200 import capi, model as M, spec


def test_spec_equals_reference_state_dict():
    ref = load_spec()
    mine = spec.model_spec()
    assert [(k, list(s), d) for k, s, d, _, _ in mine] == [(k, list(s), d) for k, s, d in ref["keys"]]
    assert sorted(k for k, _, _, kind, _ in mine if kind == "param") == ref["params"]
    assert sorted(k for k, _, _, kind, tr in mine if kind == "param" and tr) == ref["trainable"]


def test_build_model_surface_and_checkpoint_layout():
    args = M.default_args(device="cpu")
    model, criterions, post = M.build_model(args)
    ref = load_spec()
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in ref["keys"]]
    assert [list(v.shape) for v in sd.values()] == [s for _, s, _ in ref["keys"]]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == ref["trainable"]
    # the optimizer split of src/main_glassrgbd.py:59-65 relies on "backbone" appearing in backbone parameter names
    assert any("backbone" in n for n, _ in model.named_parameters())
    assert len(criterions) == 4 and criterions[3] is None and set(post) == {"line"}
    crit = criterions[0]
    assert set(crit.weight_dict) == {"loss_ce", "loss_line"} | {"%s_%d" % (k, i) for k in ("loss_ce", "loss_line") for i in range(5)}
    # reference checkpoints load key-for-key
    missing = model.load_state_dict(synth_weights(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # no CPU fallback: the forward must fail loudly off-GPU
    model.eval()
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 64, 64))


def test_nested_tensor_padding():
    a, b = torch.ones(3, 4, 6), torch.ones(3, 5, 3)
    nt = M.nested_tensor_from_tensor_list([a, b])
    assert nt.tensors.shape == (2, 3, 5, 6)
    assert nt.mask[0, :4, :6].sum() == 0 and nt.mask[0, 4:].all() and nt.mask[1, :, 3:].all()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gwd_b200.h")).read()
    declared = set(re.findall(r"\b(gwd_[a-z0-9_]+)\s*\(", hdr)) - {"gwd_pack_conv3x3_weight"}
    assert declared == set(capi.SIGNATURES)
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libgwd_b200.so not built (python __graft_entry__.py build)")
    out = subprocess.check_output(["nm", "-D", capi.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T gwd_" in l}
    assert declared <= exported
    lib = capi.lib()            # loads without a GPU; no compute call is made here
    assert lib.gwd_version() >= 100 and capi.launch_count() == 0
