"""Host-side text plumbing: answer word tokenisation (gpv.py:377-430 uses nltk.word_tokenize) and the BERT query
tokenizer (bert.py:8-21 uses transformers.BertTokenizer('bert-base-uncased'), padding=True)."""
import os
import re

_TOK = re.compile(r"__\w+__|\w+|[^\w\s]")


def word_tokenize(s):
    try:
        from nltk.tokenize import word_tokenize as wt
        return wt(s)
    except Exception:
        return _TOK.findall(s)


def load_tokenizer(vocab_path=None):
    """Returns f(list[str]) -> list[list[int]] (padded with 0 like BertTokenizer(padding=True))."""
    path = vocab_path or os.environ.get("GPV_BERT_VOCAB")
    tok = None
    try:
        from transformers import BertTokenizer
        tok = BertTokenizer(path) if path else BertTokenizer.from_pretrained("bert-base-uncased", local_files_only=True)
    except Exception as e:  # no vocabulary available offline
        raise RuntimeError("no BERT vocabulary found (set cfg.bert_vocab or GPV_BERT_VOCAB to a vocab.txt, or pass "
                           "`queries` as a LongTensor of token ids)") from e

    def encode(sentences):
        return tok(sentences, padding=True)["input_ids"]
    return encode
