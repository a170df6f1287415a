"""Developer tool (GPU box): per-parameter gradient error of the CUDA path vs the fp32 oracle, worst first."""
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


def main(name="train_small", top=60, mode="all"):
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
    for t in targets:
        if mode == "text":
            t.pop("boxes", None), t.pop("labels", None)
        if mode == "loc":
            t.pop("answer", None)
    dt = [{k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in t.items()} for t in targets]
    loss = model(images.to(cuda), qids.to(cuda), ans.to(cuda), dt)
    loss.backward()
    grads = {n: p.grad.float().cpu() for n, p in model.named_parameters() if p.grad is not None}
    Pg = {kk: (v.clone().requires_grad_(True) if kk in grads else v.clone()) for kk, v in P.items()}
    ol = TO.gpv_forward(Pg, images, qids, ans, targets)
    ol.backward()
    print("loss", loss.item(), ol.item(), fix["loss"].item())
    rows = []
    for n, gm in grads.items():
        og = Pg[n].grad
        if og is None:
            rows.append((float("inf"), n, gm.norm().item(), 0.0))
            continue
        rows.append((((gm - og).norm() / (og.norm() + 1e-12)).item(), n, gm.norm().item(), og.norm().item()))
    rows.sort(key=lambda r: -r[0])
    for r in rows[: int(top)]:
        print(f"{r[0]:9.4f}  mine {r[2]:.5f}  ref {r[3]:.5f}  {r[1]}")
    import statistics
    groups = {}
    for r in rows:
        n = r[1]
        key = ".".join(n.split(".")[:5]) if n.startswith("detr.backbone") else ".".join(n.split(".")[:4]) if n.startswith("detr.transformer") else ".".join(n.split(".")[:2])
        groups.setdefault(key, []).append(r[0])
    for kk in sorted(groups):
        print(f"  {kk:50s} median {statistics.median(groups[kk]):.4f}  max {max(groups[kk]):.4f}")
    print("median rel err", statistics.median(r[0] for r in rows), "n", len(rows))


if __name__ == "__main__":
    main(*sys.argv[1:])
