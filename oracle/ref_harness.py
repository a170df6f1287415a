"""ORACLE tooling (test infrastructure only): run the UNMODIFIED reference from /root/reference in this container.

Only usable where /root/reference exists (the build container).  Nothing that runs on the GPU box imports this
module; it is used by oracle/make_golden.py and oracle/pin_matcher.py to validate the restatements and to generate
the fixtures under tests/golden/, and by bench.py --impl reference when the reference tree is present.

Shims -- all outside the reference tree:
  1. torchvision.__version__ parsed as float('0.2') by utils/detr_misc.py:20-22 -> report '0.9.0' during import
  2. stub modules: hydra, nltk(+tokenize, treebank), boto3, botocore.exceptions, torch._six
  3. exp.gpv.models.gpv.Bert -> random-init transformers.BertModel fed synthetic token ids (no hub access)
  4. backbone.is_main_process -> False so torchvision resnet50 is built with pretrained=False
  5. torch.Tensor.cuda -> identity on CPU (gpv.py hard-codes .cuda(device))
  6. YAML config loader with ${a.b} interpolation and attribute access
  7. synthetic vocab.json + vocab_embed.npy [V,768]
"""
import json
import os
import re
import sys
import types

import numpy as np
import torch
import yaml

REF_ROOT = os.environ.get("GPV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "exp", "gpv", "models"))


# ----------------------------------------------------------------------------------------------- config
class AttrDict(dict):
    """dict with attribute access and .items(), enough for the reference's `cfg.a.b` / `cfg.losses.items()` uses."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_attr(o):
    if isinstance(o, dict):
        return AttrDict({k: _to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_to_attr(v) for v in o]
    return o


_NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


def _coerce(v):
    if isinstance(v, str) and _NUM.match(v):
        f = float(v)
        return int(f) if f.is_integer() and "e" not in v.lower() and "." not in v else f
    return v


def _lookup(root, path):
    cur = root
    for p in path.split("."):
        cur = cur[p]
    return cur


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        m = re.fullmatch(r"\$\{([^}]+)\}", node)
        if m:
            return _resolve(_lookup(root, m.group(1)), root)
        out = re.sub(r"\$\{([^}]+)\}", lambda mm: str(_resolve(_lookup(root, mm.group(1)), root)), node)
        return _coerce(out)
    return node


def load_cfg(yaml_path, overrides=None):
    """Minimal Hydra/OmegaConf stand-in: YAML + ${a.b} interpolation + dotted overrides."""
    with open(yaml_path) as f:
        raw = yaml.safe_load(f)
    raw.pop("defaults", None)
    raw.pop("hydra", None)
    for k, v in (overrides or {}).items():
        cur = raw
        parts = k.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = v
    return _to_attr(_resolve(raw, raw))


# ----------------------------------------------------------------------------------------------- shims
def make_vocab(V, seed=0):
    words = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(V - 4)]
    g = torch.Generator().manual_seed(seed)
    embed = (0.1 * torch.randn(V, 768, generator=g)).numpy().astype(np.float32)
    return words, embed


def install_shims():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    def _main(*a, **k):
        def deco(fn):
            return fn
        return deco

    stub("hydra", main=_main)
    tok = stub("nltk.tokenize", word_tokenize=lambda s: s.split())

    class _Detok:
        def detokenize(self, words):
            return " ".join(words)

    tb = stub("nltk.tokenize.treebank", TreebankWordDetokenizer=_Detok)
    n = stub("nltk", tokenize=tok, word_tokenize=tok.word_tokenize)
    n.tokenize.treebank = tb
    stub("boto3")
    be = stub("botocore.exceptions", ClientError=Exception)
    stub("botocore", exceptions=be)
    stub("torch._six", inf=float("inf"), string_classes=(str,))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    import torchvision
    real = torchvision.__version__
    torchvision.__version__ = "0.9.0"
    try:
        import utils.detr_misc  # noqa: F401  (reference module)
    finally:
        torchvision.__version__ = real


class SyntheticBert(torch.nn.Module):
    """Stands in for exp/gpv/models/bert.py:Bert -- same `.model` (HF BertModel, bert-base config, random init),
    but `queries` are already-tokenised id tensors [B,T] (or a list of lists): no tokenizer files offline."""

    def __init__(self, cfg=None):
        super().__init__()
        from transformers import BertConfig, BertModel
        self.model = BertModel(BertConfig())

    def forward(self, token_ids, device=None):
        ids = torch.as_tensor(token_ids, dtype=torch.long, device=self.model.embeddings.word_embeddings.weight.device)
        # what BertTokenizer(padding=True) returns for ids padded with [PAD] = 0 (bert.py:12-15)
        out = self.model(input_ids=ids, attention_mask=(ids != 0).long())
        return out[0], {"input_ids": ids}


def build_reference_gpv(V=512, seed=0, tmpdir="/tmp/gpv_oracle", overrides=None, eval_mode=True):
    """Instantiate the reference GPV (random init, seeded) from configs/exp/gpv.yaml with a synthetic vocab."""
    install_shims()
    os.makedirs(tmpdir, exist_ok=True)
    words, embed = make_vocab(V, seed)
    vp, ep = os.path.join(tmpdir, f"vocab_{V}.json"), os.path.join(tmpdir, f"vocab_embed_{V}.npy")
    with open(vp, "w") as f:
        json.dump(words, f)
    np.save(ep, embed)
    ov = {"model.vocab": vp, "model.vocab_embed": ep}
    ov.update(overrides or {})
    cfg = load_cfg(os.path.join(REF_ROOT, "configs", "exp", "gpv.yaml"), ov)
    import exp.gpv.models.backbone as rb
    import exp.gpv.models.gpv as rg
    rb.is_main_process = lambda: False
    rg.Bert = SyntheticBert
    torch.manual_seed(seed)
    model = rg.GPV(cfg.model)
    if eval_mode:
        model.eval()
    return model, cfg


def reference_matcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0):
    install_shims()
    from utils.matcher import HungarianMatcher
    return HungarianMatcher(cost_class, cost_bbox, cost_giou)
