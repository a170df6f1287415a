"""Inference entry point: the reference's inference.py / inference_beam_search.py on the B200 path.

    python -m gpv1_b200.inference ckpt=/path/model.pth inputs.img=/path/img.jpg inputs.query="What is this?" num_output_boxes=5

Same flow as the reference (inference.py:14-104): build `GPV(cfg.model).cuda().eval()`, load a checkpoint whose keys carry
DDP's `module.` prefix (57-62), normalise the image with the ImageNet statistics (64-67), run `model(images, queries,
None)` (greedy) or `model.forward_beam_search(images, queries, beam_size)` and decode: relevance = softmax of
`pred_relevance_logits`, boxes sorted by relevance, answer = top-1 token per position cut at `__stop__` / `__pad__`
(24-49).  Here the image may be handed over as raw uint8 HWC pixels (the normalisation is fused into the stem kernel),
and `model.inference_graphs = True` replays the whole call from one CUDA graph on repeated shapes.
"""
import sys

import numpy as np
import torch

from .config import load_config
from .model import GPV


def read_image(path):
    """-> uint8 [H,W,3] RGB (inference_util.read_image without the resize)."""
    if path.endswith(".npy"):
        img = np.load(path)
    else:
        try:
            from PIL import Image
            img = np.asarray(Image.open(path).convert("RGB"))
        except ImportError:
            import cv2
            img = cv2.imread(path)[:, :, ::-1]
    if img.dtype != np.uint8:
        img = (255 * img).astype(np.uint8)                     # inference.py:17
    if img.ndim == 2:
        img = np.tile(img[:, :, None], (1, 1, 3))
    return np.ascontiguousarray(img[:, :, :3])


def preprocess(inputs):
    """inputs: list of (uint8 HWC image, query string or token-id row) -> (list of uint8 [1,H,W,3] tensors or one batch,
    queries), the counterpart of inference.py:14-22 + collate_fn; images of one size are stacked into a uint8 NHWC batch."""
    imgs = [torch.from_numpy(np.ascontiguousarray(i)) for i, _ in inputs]
    queries = [q for _, q in inputs]
    if len({tuple(i.shape) for i in imgs}) == 1:
        return torch.stack(imgs), queries
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    return [((i.permute(2, 0, 1).float() / 255.0) - mean) / std for i in imgs], queries   # mixed sizes: padded by the model


def detokenize(tokens):
    """Stand-in for nltk's TreebankWordDetokenizer (inference.py:47) when nltk is absent."""
    try:
        from nltk.tokenize.treebank import TreebankWordDetokenizer
        return TreebankWordDetokenizer().detokenize(tokens)
    except ImportError:
        out = ""
        for t in tokens:
            out += t if (not out or t in ",.!?;:'s" or t.startswith("'")) else " " + t
        return out


def decode_outputs(outputs, model):
    """inference.py:24-49."""
    relevance = outputs["pred_relevance_logits"].float().softmax(-1).detach().cpu().numpy()
    pred_boxes = outputs["pred_boxes"].float().detach().cpu().numpy()
    ids = torch.topk(outputs["answer_logits"][-1], k=1, dim=-1).indices.detach().cpu().numpy()
    pred_answers = model.token_ids_to_words(ids[:, :, 0])
    decoded = []
    for b in range(len(pred_answers)):
        order = sorted(range(relevance.shape[1]), key=lambda q: relevance[b, q, 0], reverse=True)
        answer = []
        for tok in pred_answers[b]:
            if tok in ("__stop__", "__pad__"):
                break
            answer.append(tok)
        decoded.append({"answer": detokenize(answer), "boxes": pred_boxes[b][order].astype(np.float32),
                        "relevance": relevance[b, order, 0].astype(np.float32)})
    return decoded


def load_model(cfg, ckpt=None, device="cuda:0"):
    model = GPV(cfg.model).to(device).eval()
    if ckpt:
        loaded = torch.load(ckpt, map_location=device)["model"]                # inference.py:57-62
        sd = model.state_dict()
        for k in sd:
            sd[k] = loaded[f"module.{k}"] if f"module.{k}" in loaded else loaded[k]
        model.load_state_dict(sd)
    for p in model.parameters():
        p.requires_grad = False
    return model


def main(argv=None):
    cfg = load_config(overrides=list(sys.argv[1:] if argv is None else argv))
    model = load_model(cfg, getattr(cfg, "ckpt", None))
    img = read_image(cfg.inputs.img)
    images, queries = preprocess([(img, cfg.inputs.query)])
    beam = int(getattr(cfg, "beam_size", 0) or 0)
    if beam > 1:                                                               # inference_beam_search.py:68
        out = model.forward_beam_search(images, queries, beam)
        for kk, (words, prob) in enumerate(zip(out["answers"][0], out["answer_probs"][0])):
            print(f"beam {kk}: p={prob:.4g}  {detokenize(words)}")
        return out
    outputs = model(images, queries, None)
    pred = decode_outputs(outputs, model)[0]
    n = int(getattr(cfg, "num_output_boxes", 5))
    pred["boxes"], pred["relevance"] = pred["boxes"][:n], pred["relevance"][:n]
    for kind, value in pred.items():
        print("-" * 80)
        print(kind)
        print("-" * 80)
        print(value)
    return pred


if __name__ == "__main__":
    main()
