"""Names, shapes and trainability of every tensor in the reference GPV `state_dict()` (836 entries).

The reference checkpoints are loaded by name (inference.py:57-62, train_distr.py:264-272), so the B200 model keeps
the exact keys: exp/gpv/models/gpv.py:58-119 (top level), detr_roi_head.py:21-42, backbone.py:82-97 (torchvision
resnet50 with FrozenBatchNorm2d), transformer.py:18-36, vilbert.py:696-870 (BertConnectionLayer), bert.py:8-9
(HF BertModel), answer_head.py:8-24, set_criterion.py:40-42 (empty_weight buffer).
"""
from collections import namedtuple

Spec = namedtuple("Spec", "name shape kind")  # kind: "param" (trainable) | "frozen" (Parameter, no grad) | "buffer"

RESNET50_LAYERS = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]  # (planes, blocks, stride of first block)


def resnet_blocks():
    """Yield (layer_idx, block_idx, inplanes, planes, stride, has_downsample) for torchvision resnet50 (v1.5)."""
    inplanes = 64
    for li, (planes, blocks, stride) in enumerate(RESNET50_LAYERS, start=1):
        for bi in range(blocks):
            s = stride if bi == 0 else 1
            ds = bi == 0 and (s != 1 or inplanes != planes * 4)
            yield li, bi, inplanes, planes, s, ds
            inplanes = planes * 4


def _bn(prefix, c):
    return [Spec(f"{prefix}.{n}", (c,), "buffer") for n in ("weight", "bias", "running_mean", "running_var")]


def _lin(prefix, out_f, in_f, kind="param"):
    return [Spec(f"{prefix}.weight", (out_f, in_f), kind), Spec(f"{prefix}.bias", (out_f,), kind)]


def _ln(prefix, d, kind="param"):
    return [Spec(f"{prefix}.weight", (d,), kind), Spec(f"{prefix}.bias", (d,), kind)]


def _mha(prefix, d):
    return [Spec(f"{prefix}.in_proj_weight", (3 * d, d), "param"), Spec(f"{prefix}.in_proj_bias", (3 * d,), "param")] + \
        _lin(f"{prefix}.out_proj", d, d)


def gpv_specs(cfg, V):
    """cfg: the `model:` block of configs/exp/gpv.yaml (attribute access). V: vocabulary size."""
    d = cfg.detr.hidden_dim
    D = cfg.hidden_dim
    ff = cfg.detr.dim_feedforward
    out = [Spec("vision_token", (D,), "param"), Spec("lang_token", (D,), "param"), Spec("relevance_tokens", (2, D), "param"),
           Spec("pos_enc", (1, cfg.max_pos_enc_len, cfg.text_decoder.hidden_dim), "frozen")]
    # ---- detr.backbone.0.body (backbone.py:61-63: only layer2-4 conv weights train)
    bb = "detr.backbone.0.body"
    out.append(Spec(f"{bb}.conv1.weight", (64, 3, 7, 7), "frozen"))
    out += _bn(f"{bb}.bn1", 64)
    for li, bi, inp, planes, stride, ds in resnet_blocks():
        kind = "param" if li >= 2 else "frozen"
        p = f"{bb}.layer{li}.{bi}"
        out.append(Spec(f"{p}.conv1.weight", (planes, inp, 1, 1), kind))
        out += _bn(f"{p}.bn1", planes)
        out.append(Spec(f"{p}.conv2.weight", (planes, planes, 3, 3), kind))
        out += _bn(f"{p}.bn2", planes)
        out.append(Spec(f"{p}.conv3.weight", (planes * 4, planes, 1, 1), kind))
        out += _bn(f"{p}.bn3", planes * 4)
        if ds:
            out.append(Spec(f"{p}.downsample.0.weight", (planes * 4, inp, 1, 1), kind))
            out += _bn(f"{p}.downsample.1", planes * 4)
    # ---- detr transformer (transformer.py)
    for i in range(cfg.detr.num_encoder_layers):
        p = f"detr.transformer.encoder.layers.{i}"
        out += _mha(f"{p}.self_attn", d) + _lin(f"{p}.linear1", ff, d) + _lin(f"{p}.linear2", d, ff)
        out += _ln(f"{p}.norm1", d) + _ln(f"{p}.norm2", d)
    for i in range(cfg.detr.num_decoder_layers):
        p = f"detr.transformer.decoder.layers.{i}"
        out += _mha(f"{p}.self_attn", d) + _mha(f"{p}.multihead_attn", d)
        out += _lin(f"{p}.linear1", ff, d) + _lin(f"{p}.linear2", d, ff)
        out += _ln(f"{p}.norm1", d) + _ln(f"{p}.norm2", d) + _ln(f"{p}.norm3", d)
    out += _ln("detr.transformer.decoder.norm", d)
    out += _lin("detr.class_embed", cfg.detr.num_classes + 1, d)
    out += _lin("detr.bbox_embed.layers.0", d, d) + _lin("detr.bbox_embed.layers.1", d, d) + _lin("detr.bbox_embed.layers.2", 4, d)
    out.append(Spec("detr.query_embed.weight", (cfg.detr.num_queries, d), "param"))
    out += [Spec("detr.input_proj.weight", (d, 2048, 1, 1), "param"), Spec("detr.input_proj.bias", (d,), "param")]
    out += _lin("detr_joiner", cfg.detr_joiner.out_dim, cfg.detr_joiner.detr_dim)
    # ---- bert.model (HF BertModel, bert-base-uncased config; never receives gradients: gpv.py:142-145)
    b = "bert.model"
    out += [Spec(f"{b}.embeddings.word_embeddings.weight", (30522, 768), "param"),
            Spec(f"{b}.embeddings.position_embeddings.weight", (512, 768), "param"),
            Spec(f"{b}.embeddings.token_type_embeddings.weight", (2, 768), "param")]
    out += _ln(f"{b}.embeddings.LayerNorm", 768)
    for i in range(12):
        p = f"{b}.encoder.layer.{i}"
        for n in ("query", "key", "value"):
            out += _lin(f"{p}.attention.self.{n}", 768, 768)
        out += _lin(f"{p}.attention.output.dense", 768, 768) + _ln(f"{p}.attention.output.LayerNorm", 768)
        out += _lin(f"{p}.intermediate.dense", 3072, 768) + _lin(f"{p}.output.dense", 768, 3072) + _ln(f"{p}.output.LayerNorm", 768)
    out += _lin(f"{b}.pooler.dense", 768, 768)
    out += _lin("bert_joiner", cfg.bert_joiner.out_dim, cfg.bert_joiner.bert_dim)
    # ---- co-attention (vilbert.py:859-870)
    ca = cfg.co_att
    H = ca.bi_hidden_size
    for i in range(ca.num_layers):
        p = f"co_att_transformer.{i}"
        for n in ("query1", "key1", "value1"):
            out += _lin(f"{p}.biattention.{n}", H, ca.v_hidden_size)
        for n in ("query2", "key2", "value2"):
            out += _lin(f"{p}.biattention.{n}", H, ca.hidden_size)
        out += _lin(f"{p}.biOutput.dense1", ca.v_hidden_size, H) + _ln(f"{p}.biOutput.LayerNorm1", ca.v_hidden_size)
        out += _lin(f"{p}.biOutput.q_dense1", ca.v_hidden_size, H)
        out += _lin(f"{p}.biOutput.dense2", ca.hidden_size, H) + _ln(f"{p}.biOutput.LayerNorm2", ca.hidden_size)
        out += _lin(f"{p}.biOutput.q_dense2", ca.hidden_size, H)
        out += _lin(f"{p}.v_intermediate.dense", ca.v_intermediate_size, ca.v_hidden_size)
        out += _lin(f"{p}.v_output.dense", ca.v_hidden_size, ca.v_intermediate_size) + _ln(f"{p}.v_output.LayerNorm", ca.v_hidden_size)
        out += _lin(f"{p}.t_intermediate.dense", ca.intermediate_size, ca.hidden_size)
        out += _lin(f"{p}.t_output.dense", ca.hidden_size, ca.intermediate_size) + _ln(f"{p}.t_output.LayerNorm", ca.hidden_size)
    out += _lin("relevance_predictor", cfg.detr.num_classes + 1, D)
    # ---- text decoder: nn.TransformerDecoderLayer(d_model, nhead, dropout) -> dim_feedforward 2048 (gpv.py:37-43)
    td = cfg.text_decoder.hidden_dim
    for i in range(cfg.text_decoder.num_layers):
        p = f"text_decoder.layers.{i}"
        out += _mha(f"{p}.self_attn", td) + _mha(f"{p}.multihead_attn", td)
        out += _lin(f"{p}.linear1", 2048, td) + _lin(f"{p}.linear2", td, 2048)
        out += _ln(f"{p}.norm1", td) + _ln(f"{p}.norm2", td) + _ln(f"{p}.norm3", td)
    out.append(Spec("answer_head.vocab_embed", (V, cfg.bert_joiner.bert_dim), "frozen"))
    out += _lin("answer_head.classifier_transform", cfg.bert_joiner.out_dim, cfg.bert_joiner.bert_dim)
    out += _lin("answer_input_embedings.transform", cfg.bert_joiner.out_dim, cfg.bert_joiner.bert_dim)
    out.append(Spec("answer_input_embedings.embedding_layer.weight", (V, cfg.bert_joiner.bert_dim), "frozen"))
    out.append(Spec("criterion.localization_criterion.set_criterion.empty_weight", (cfg.detr.num_classes + 1,), "buffer"))
    return out


# Parameters that exist and have requires_grad=True in the reference but never receive a gradient in GPV.forward
# (DDP needs find_unused_parameters=True for them, train_distr.py:192-193): all of BERT (no_grad, gpv.py:142),
# the indicator tokens (never used), and BertBiOutput.q_dense1/2 (vilbert.py:835-843, unused in forward 845-856).
def never_gets_grad(name):
    return name.startswith("bert.") or name in ("vision_token", "lang_token") or ".biOutput.q_dense" in name


# Backward completes the gradient arena in this order; the data-parallel all-reduce is bucketed along it
# (parallel.py) so that every bucket but the last is reduced underneath the remaining backward kernels.
N_STAGES = 7


def grad_stage(name):
    if name.startswith(("text_decoder.", "answer_head.", "answer_input_embedings.")):
        return 0
    if name.startswith(("co_att_transformer.", "relevance_predictor.", "bert_joiner.")) or name == "relevance_tokens":
        return 1
    if name.startswith("detr.backbone."):
        for li, st in ((4, 4), (3, 5), (2, 6)):
            if f".layer{li}." in name:
                return st
        raise KeyError(name)
    if name.startswith(("detr.transformer.encoder.", "detr.input_proj.")):
        return 3
    if name.startswith(("detr.", "detr_joiner.")):
        return 2
    raise KeyError(name)
