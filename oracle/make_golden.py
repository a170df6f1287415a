"""ORACLE tooling: pin oracle/torch_oracle.py against the UNMODIFIED reference (imported from /root/reference through
oracle/ref_harness.py) and write the fixtures tests/golden/gpv_*.pt that travel to the GPU box.

    python -m oracle.make_golden            # run in the build container (needs /root/reference)

For each case the same seeded weights (torch_oracle.make_state over the reference's own state_dict keys) are loaded
into the reference GPV with load_state_dict(strict=True); reference and oracle then run on identical inputs.  The
script asserts oracle == reference (fp32 tolerances below) and stores the REFERENCE's numbers:
  train case    total loss, the six loss terms, matcher indices, answer/relevance logits, boxes, and for every
                parameter with a gradient: its L2 norm and 8 sampled entries
  greedy case   token ids and the 20-position answer logits of GPV.forward(images, queries, None)
  beam case     GPV.forward_beam_search sequences (as ids) and probabilities
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness, torch_oracle as TO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
V = 512


def specs_from_model(model):
    sd = model.state_dict()
    params = dict(model.named_parameters())
    out = []
    for k, v in sd.items():
        kind = "buffer" if k not in params else ("param" if params[k].requires_grad else "frozen")
        out.append((k, tuple(v.shape), kind))
    return out


def make_inputs(B, H, W, Tl, S, seed, tasks, qpad=None):
    """qpad: per-sample number of trailing [PAD] = 0 query tokens (what BertTokenizer(padding=True) appends to the shorter
    queries of a batch, bert.py:12-15); drawn after everything else so that fixtures made without it are unchanged."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, H, W, generator=g)
    qids = torch.randint(1000, 30000, (B, Tl), generator=g)
    if qpad is not None:
        for b, npad in enumerate(qpad):
            if npad:
                qids[b, Tl - npad:] = 0
    ans = torch.randint(4, V, (B, S), generator=g)
    ans[:, 0] = 1
    ans[:, -1] = 2
    targets = []
    for b in range(B):
        t = {"task": tasks[b % len(tasks)]}
        if t["task"] != "CocoDetection":
            t["answer"] = "x"
        nb = int(torch.randint(1, 9, (1,), generator=g))
        if t["task"] in ("CocoDetection", "CocoCaptioning") or b % 2 == 0:
            cxcy = 0.25 + 0.5 * torch.rand(nb, 2, generator=g)
            wh = 0.05 + 0.3 * torch.rand(nb, 2, generator=g)
            t["boxes"] = torch.cat((cxcy, wh), -1)
            t["labels"] = torch.zeros(nb, dtype=torch.long)
        t["answer_token_ids"] = ans[b, 1:]
        targets.append(t)
    return images, qids, ans, targets


def sample_idx(numel, n=8):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (n,), generator=g)


def close(a, b, rtol, atol, what):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= atol + rtol * ref, f"{what}: oracle deviates from the reference by {err} (max |ref| {ref})"
    return err


def crop_list(images, sizes):
    """Mixed-size batch: image b is the top-left sizes[b] = (h, w) crop of the seeded full-size image."""
    return [images[b, :, :h, :w].clone() for b, (h, w) in enumerate(sizes)]


def train_case(model, P, name, B, H, W, Tl, S, seed, tasks, sizes=None, qpad=None):
    images, qids, ans, targets = make_inputs(B, H, W, Tl, S, seed, tasks, qpad)
    model.zero_grad()
    t0 = time.time()
    # outputs (teacher forced), then the loss + grads through the reference's own criterion
    mask = None
    if sizes is not None:                      # the reference pads the list itself (detr_roi_head.py:59-60)
        ref_in = crop_list(images, sizes)
        images, mask = TO.nested(ref_in)
    else:
        ref_in = images
    out = model(ref_in, qids, ans, None)
    loss, ld = model.criterion(out, targets)
    loss.backward()
    t_ref = time.time() - t0
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    ind = model.criterion.localization_criterion.matcher(
        {"pred_relevance_logits": out["pred_relevance_logits"][[i for i, t in enumerate(targets) if "boxes" in t]],
         "pred_boxes": out["pred_boxes"][[i for i, t in enumerate(targets) if "boxes" in t]]},
        [t for t in targets if "boxes" in t])

    trainable = {k for k, p in model.named_parameters() if p.requires_grad}
    Pg = {k: (v.clone().requires_grad_(True) if k in trainable else v.clone()) for k, v in P.items()}
    oout = TO.gpv_forward(Pg, images, qids, ans, None, mask=mask)
    oloss, old = TO.gpv_criterion(oout, targets)
    oloss.backward()
    errs = {"loss": abs(oloss.item() - loss.item())}
    assert errs["loss"] <= 1e-4 * abs(loss.item()), errs
    for k in ("pred_relevance_logits", "pred_boxes", "answer_logits", "detr_hs"):
        errs[k] = close(oout[k], out[k], 1e-4, 1e-5, k)
    for (q, t), (oq, ot) in zip(ind, old["_indices"]):
        assert torch.equal(q, oq) and torch.equal(t, ot), "matcher indices differ"
    for k, gr in grads.items():
        og = Pg[k].grad
        assert og is not None, f"oracle has no grad for {k}"
        close(og, gr, 2e-3, 2e-5, "grad " + k)
    for k, v in Pg.items():
        if v.requires_grad and v.grad is not None and v.grad.abs().max() > 0:
            assert k in grads, f"oracle has a grad the reference lacks: {k}"
    fix = {
        "meta": {"B": B, "H": H, "W": W, "Tl": Tl, "S": S, "seed": seed, "tasks": tasks, "V": V, "weights_seed": 0, "sizes": sizes, "qpad": qpad,
                 "ref_seconds": t_ref, "oracle_max_err": errs},
        "loss": loss.detach(), "losses": {k: (v.detach() if torch.is_tensor(v) else v) for k, v in ld.items() if v is not None and k != "class_error"},
        "indices": [(q.clone(), t.clone()) for q, t in ind],
        "pred_relevance_logits": out["pred_relevance_logits"].detach(), "pred_boxes": out["pred_boxes"].detach(),
        "answer_logits": out["answer_logits"].detach().half() if out["answer_logits"].numel() > 2_000_000 else out["answer_logits"].detach(),
        "detr_hs_sample": out["detr_hs"].detach().flatten()[sample_idx(out["detr_hs"].numel(), 256)],
        "grad_norm": {k: v.norm().item() for k, v in grads.items()},
        "grad_sample": {k: v.flatten()[sample_idx(v.numel())].clone() for k, v in grads.items()},
    }
    torch.save(fix, os.path.join(GOLD, f"gpv_{name}.pt"))
    print(f"[{name}] reference loss {loss.item():.6f}  oracle err {errs}  ref fwd+bwd {t_ref:.1f}s  grads {len(grads)}")


def greedy_case(model, P, name, B, H, W, Tl, seed, max_text_len):
    images, qids, _, _ = make_inputs(B, H, W, Tl, 4, seed, ["CocoVqa"])
    model.cfg.max_text_len = max_text_len
    with torch.no_grad():
        out = model(images, qids, None)
        oout = TO.gpv_forward(P, images, qids, None, max_text_len=max_text_len)
    close(oout["answer_logits"], out["answer_logits"], 1e-4, 1e-5, "greedy logits")
    ids = out["answer_logits"].topk(1, -1).indices[0, :, :, 0]
    assert torch.equal(ids, oout["answer_logits"].topk(1, -1).indices[0, :, :, 0])
    fix = {"meta": {"B": B, "H": H, "W": W, "Tl": Tl, "seed": seed, "V": V, "max_text_len": max_text_len, "weights_seed": 0},
           "answer_logits": out["answer_logits"], "ids": ids, "pred_boxes": out["pred_boxes"],
           "pred_relevance_logits": out["pred_relevance_logits"]}
    torch.save(fix, os.path.join(GOLD, f"gpv_{name}.pt"))
    print(f"[{name}] greedy ids[0] = {ids[0].tolist()}")


def beam_case(model, P, name, B, H, W, Tl, seed, K, max_text_len):
    images, qids, _, _ = make_inputs(B, H, W, Tl, 4, seed, ["CocoVqa"])
    model.cfg.max_text_len = max_text_len
    # the reference returns words cut at __stop__/__pad__; map back to ids for an exact comparison
    with torch.no_grad():
        out = model.forward_beam_search(images, qids, K)
        _, memory = TO.gpv_encode(P, images, qids)
        seqs, score = TO.beam_search(P, memory, K, max_text_len)
    w2i = model.word_to_idx
    ref_ids = [[[w2i[w] for w in out["answers"][b][k]] for k in range(K)] for b in range(B)]
    stop = {w2i["__stop__"], w2i["__pad__"]}
    for b in range(B):
        for k in range(K):
            mine = []
            for t in seqs[b, k].tolist():
                if t in stop:
                    break
                mine.append(t)
            assert mine == ref_ids[b][k], (b, k, mine, ref_ids[b][k])
            assert abs(score[b, k].exp().item() - out["answer_probs"][b][k]) <= 1e-4 * max(out["answer_probs"][b][k], 1e-30) + 1e-12
    fix = {"meta": {"B": B, "H": H, "W": W, "Tl": Tl, "seed": seed, "V": V, "K": K, "max_text_len": max_text_len, "weights_seed": 0},
           "answers_ids": ref_ids, "answer_probs": out["answer_probs"], "seqs": seqs, "log_prob": score}
    torch.save(fix, os.path.join(GOLD, f"gpv_{name}.pt"))
    print(f"[{name}] beam probs[0] = {out['answer_probs'][0]}")


def answers_case(model):
    """GPV.encode_answers / token_ids_to_words of the reference (gpv.py:377-441) on answers that exercise lower-casing,
    padding to the batch maximum, truncation to max_text_len, out-of-vocabulary words and the empty answer."""
    model.cfg.max_text_len = 20
    vocab = model.vocab
    words = [w for w in vocab if not w.startswith("__")]
    long_answer = " ".join(words[i % len(words)] for i in range(3, 3 + 25))      # > max_text_len - 2 words
    cases = {
        "mixed": [{"answer": f"{words[0]} {words[5]}"}, {"answer": words[7].upper()}, {"answer": ""}, {"answer": f"{words[2]} zzzunknownzzz {words[9]}"}],
        "single": [{"answer": words[11]}],
        "long": [{"answer": long_answer}, {"answer": words[1]}],
    }
    out = {}
    for name, targets in cases.items():
        padded, ids = model.encode_answers(targets)
        out[name] = {"targets": targets, "padded": [list(p) for p in padded], "ids": ids.cpu().tolist(),
                     "words": model.token_ids_to_words(ids.cpu())}
    json.dump({"V": V, "max_text_len": int(model.cfg.max_text_len), "cases": out}, open(os.path.join(GOLD, "gpv_answers.json"), "w"), indent=1)
    print("[answers]", {k: [len(r) for r in v["ids"]] for k, v in out.items()})


def main():
    assert ref_harness.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    model, cfg = ref_harness.build_reference_gpv(V=V, seed=0, eval_mode=True)
    specs = specs_from_model(model)
    json.dump({"V": V, "specs": [[n, list(s), k] for n, s, k in specs]}, open(os.path.join(GOLD, "gpv_specs.json"), "w"))
    P = TO.make_state(specs, seed=0)
    model.load_state_dict(P, strict=True)
    model.eval()
    which = sys.argv[1:] or ["train_small", "train_mixed", "greedy", "beam", "train_full", "train_padded", "train_qpad", "beam5", "answers"]
    if "train_small" in which:
        train_case(model, P, "train_small", B=2, H=224, W=288, Tl=6, S=7, seed=11, tasks=["CocoCaptioning"])
    if "train_mixed" in which:
        train_case(model, P, "train_mixed", B=5, H=192, W=256, Tl=9, S=5, seed=12,
                   tasks=["CocoCaptioning", "CocoVqa", "CocoDetection", "CocoClassification", "CocoVqa"])
    if "greedy" in which:
        greedy_case(model, P, "greedy", B=2, H=224, W=288, Tl=6, seed=13, max_text_len=8)
    if "beam" in which:
        beam_case(model, P, "beam", B=2, H=224, W=288, Tl=6, seed=14, K=3, max_text_len=5)
    if "train_padded" in which:
        train_case(model, P, "train_padded", B=3, H=192, W=256, Tl=7, S=5, seed=16, tasks=["CocoCaptioning", "CocoVqa", "CocoDetection"],
                   sizes=[(192, 224), (160, 256), (128, 200)])
    if "train_qpad" in which:     # mixed-length queries: [PAD] keys masked inside BERT (bert.py:12-21)
        train_case(model, P, "train_qpad", B=3, H=192, W=256, Tl=9, S=5, seed=17, tasks=["CocoVqa", "CocoCaptioning", "CocoClassification"],
                   qpad=[0, 3, 5])
    if "beam5" in which:          # SURVEY 8d config 4 parity shape: B=4, K=5, max_text_len=5
        beam_case(model, P, "beam5", B=4, H=192, W=256, Tl=6, seed=18, K=5, max_text_len=5)
    if "answers" in which:
        answers_case(model)
    if "train_full" in which:
        train_case(model, P, "train_full", B=2, H=480, W=640, Tl=20, S=20, seed=15, tasks=["CocoCaptioning"])


if __name__ == "__main__":
    main()
