"""CPU: the fp32 oracle (oracle/torch_oracle.py) against the REFERENCE's numbers committed under tests/golden/ (written by
oracle/make_golden.py from the unmodified /root/reference modules).  This is what pins the checker the GPU parity tests
rely on; it needs neither a GPU nor /root/reference."""
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def state():
    from oracle import torch_oracle as TO
    g = json.load(open(os.path.join(GOLD, "gpv_specs.json")))
    return TO.make_state([tuple(s) for s in g["specs"]], seed=0)


@pytest.mark.parametrize("name", ["train_small", "train_padded"])
def test_oracle_loss_and_outputs_equal_reference(state, name):
    from oracle import torch_oracle as TO
    from oracle.make_golden import crop_list, make_inputs
    fix = torch.load(os.path.join(GOLD, f"gpv_{name}.pt"), weights_only=False)
    m = fix["meta"]
    images, qids, ans, targets = make_inputs(m["B"], m["H"], m["W"], m["Tl"], m["S"], m["seed"], m["tasks"])
    mask = None
    if m.get("sizes"):
        images, mask = TO.nested(crop_list(images, m["sizes"]))
    with torch.no_grad():
        out = TO.gpv_forward(state, images, qids, ans, None, mask=mask)
        loss, ld = TO.gpv_criterion(out, targets)
    assert abs(loss.item() - fix["loss"].item()) <= 1e-4 * abs(fix["loss"].item())
    for key in ("pred_relevance_logits", "pred_boxes", "answer_logits"):
        ref = fix[key].float()
        assert (out[key] - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 2e-3 * (fix[key].dtype == torch.float16), key
    for (q, t), (rq, rt) in zip(ld["_indices"], fix["indices"]):
        assert torch.equal(q, rq) and torch.equal(t, rt)
