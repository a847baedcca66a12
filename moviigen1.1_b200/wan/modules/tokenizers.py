"""Tokenizer wrapper — out of scope with the text encoder (see t5.py)."""


class HuggingfaceTokenizer:
    def __init__(self, *a, **k):
        raise NotImplementedError("HuggingfaceTokenizer belongs to the out-of-scope text-encoder path")
