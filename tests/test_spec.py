"""Host-side contract: the state_dict key set / shapes equal the reference's (golden dump from /root/reference)."""
import json
import os

from gpv1_b200.config import load_config
from gpv1_b200.model.spec import gpv_specs, never_gets_grad

GOLD = os.path.join(os.path.dirname(__file__), "golden", "state_dict_spec.json")


def test_specs_match_reference_state_dict():
    g = json.load(open(GOLD))
    cfg = load_config()
    specs = gpv_specs(cfg.model, g["V"])
    mine = {s.name: list(s.shape) for s in specs}
    ref = {k: v[0] for k, v in g["keys"].items()}
    assert set(mine) == set(ref), (sorted(set(mine) - set(ref))[:5], sorted(set(ref) - set(mine))[:5])
    for k in ref:
        assert mine[k] == ref[k], k
    params = {s.name for s in specs if s.kind in ("param", "frozen")}
    assert params == set(g["parameters"])
    trainable = {s.name for s in specs if s.kind == "param"}
    assert trainable == set(g["requires_grad"])
    live = {n for n in trainable if not never_gets_grad(n)}
    assert 380 < len(live) < 420


def test_config_interpolation():
    cfg = load_config(overrides=["training.lr_backbone=0.5", "model.max_text_len=5"])
    assert cfg.model.detr.lr_backbone == 0.5 and cfg.model.max_text_len == 5
    assert cfg.model.losses.Localization.num_classes == 1
    assert abs(cfg.losses.CaptionLoss.loss_wts.loss_caption - 0.05) < 1e-12
    assert list(cfg.model.losses.keys()) == ["CaptionLoss", "VqaLoss", "ClsLoss", "Localization"]
