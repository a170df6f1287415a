"""Developer tool (GPU box): compare the engine's intermediate activations with the fp32 oracle, stage by stage."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpv1_b200.config import load_config  # noqa: E402
from gpv1_b200.model import GPV  # noqa: E402
from oracle import torch_oracle as TO  # noqa: E402
from oracle.make_golden import make_inputs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.float().cpu().reshape(-1), b.float().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item(), b.abs().max().item()


def main(name="train_small"):
    cuda = torch.device("cuda:0")
    g = json.load(open(os.path.join(GOLD, "gpv_specs.json")))
    V = g["V"]
    P = TO.make_state([tuple(s) for s in g["specs"]], seed=0)
    vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(V - 4)]
    model = GPV(load_config().model, vocab=vocab, vocab_embed=P["answer_head.vocab_embed"].numpy())
    model.load_state_dict(P, strict=True)
    model.to(cuda)
    fix = torch.load(os.path.join(GOLD, f"gpv_{name}.pt"), weights_only=False)
    m = fix["meta"]
    images, qids, ans, targets = make_inputs(m["B"], m["H"], m["W"], m["Tl"], m["S"], m["seed"], m["tasks"])
    tr = {}
    with torch.no_grad():
        oout, omem = TO.gpv_encode(P, images, qids, trace=tr)
    eng = model.engine
    s = eng.encode(images.to(cuda), qids.to(cuda), save=True)
    B, Q = m["B"], 100
    S = s["S"]

    def show(tag, mine, ref):
        print(f"{tag:28s} rel-L2 {rel(mine, ref)[0]:.4f}  max-err {rel(mine, ref)[1]:.4f}  |ref|max {rel(mine, ref)[2]:.3f}")

    show("c5", s["c5"].permute(0, 3, 1, 2), tr["c5"])
    for i in range(6):
        x_out = s["enc"][i + 1][0][0] if i < 5 else s["mem"]
        show(f"enc{i}", x_out.view(B, S, -1), tr[f"enc{i}"])
    for i in range(6):
        t_out = s["dec"][i + 1][0][0] if i < 5 else s["t_final"]
        show(f"dec{i}", t_out.view(B, Q, -1), tr[f"dec{i}"])
    show("hs", s["detr_hs"][:, 2048:].reshape(B, Q, -1), tr["hs"])
    show("boxes", s["boxes"].view(B, Q, 8)[:, :, :4], oout["pred_boxes"])
    show("roi_raw", s["roi_raw"].view(B, Q, -1), tr["roi_raw"])
    show("detr_hs joined", s["detr_hs_joined"].view(B, Q, -1), oout["detr_hs"][0])
    show("bert", s["qe_b"].view(B, m["Tl"], -1), tr["bert"])
    for i in range(3):
        lang = s["co"][i + 1][0] if i < 2 else s["lang"]
        vis = s["co"][i + 1][1] if i < 2 else s["vis"]
        show(f"co{i}_lang", lang.view(B, m["Tl"], -1), tr[f"co{i}_lang"])
        show(f"co{i}_vis", vis.view(B, Q, -1), tr[f"co{i}_vis"])
    show("relevance logits", s["logits"].view(B, Q, 8)[:, :, :2], oout["pred_relevance_logits"])
    show("memory", s["memory"].view(B, -1, 768), omem)


if __name__ == "__main__":
    main(*sys.argv[1:])
