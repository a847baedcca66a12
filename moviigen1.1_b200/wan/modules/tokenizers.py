"""Prompt tokenizer front-end of the text encoder (reference: wan/modules/tokenizers.py:38-82).

Host-side string handling around `transformers.AutoTokenizer` (the umT5 sentencepiece vocabulary shipped in the
checkpoint directory under `google/umt5-xxl`): text cleaning, padding / truncation to `seq_len`, ids + mask out.
`ftfy` (mojibake repair, tokenizers.py:13) is used when importable; without it the text is passed through unchanged
and a warning says so once.
"""
import html
import logging
import re
import string

__all__ = ["HuggingfaceTokenizer"]

try:
    import ftfy
except ImportError:  # not in every image; not arithmetic
    ftfy = None
_warned = []


def basic_clean(text):
    """tokenizers.py:12-15."""
    if ftfy is not None:
        text = ftfy.fix_text(text)
    elif not _warned:
        _warned.append(1)
        logging.warning("ftfy is not installed: prompts are not mojibake-repaired before tokenisation")
    return html.unescape(html.unescape(text)).strip()


def whitespace_clean(text):
    """tokenizers.py:18-21."""
    return re.sub(r"\s+", " ", text).strip()


def canonicalize(text, keep_punctuation_exact_string=None):
    """tokenizers.py:24-35."""
    drop = str.maketrans("", "", string.punctuation)
    text = text.replace("_", " ")
    if keep_punctuation_exact_string:
        text = keep_punctuation_exact_string.join(part.translate(drop)
                                                  for part in text.split(keep_punctuation_exact_string))
    else:
        text = text.translate(drop)
    return re.sub(r"\s+", " ", text.lower()).strip()


class HuggingfaceTokenizer:
    def __init__(self, name, seq_len=None, clean=None, **kwargs):
        if clean not in (None, "whitespace", "lower", "canonicalize"):
            raise ValueError("unknown cleaning mode %r" % (clean,))
        from transformers import AutoTokenizer
        self.name, self.seq_len, self.clean = name, seq_len, clean
        self.tokenizer = AutoTokenizer.from_pretrained(name, **kwargs)
        self.vocab_size = self.tokenizer.vocab_size

    def __call__(self, sequence, **kwargs):
        return_mask = kwargs.pop("return_mask", False)
        opts = {"return_tensors": "pt"}
        if self.seq_len is not None:
            opts.update(padding="max_length", truncation=True, max_length=self.seq_len)
        opts.update(kwargs)
        if isinstance(sequence, str):
            sequence = [sequence]
        if self.clean:
            sequence = [self._clean(u) for u in sequence]
        enc = self.tokenizer(sequence, **opts)
        return (enc.input_ids, enc.attention_mask) if return_mask else enc.input_ids

    def _clean(self, text):
        if self.clean == "whitespace":
            return whitespace_clean(basic_clean(text))
        if self.clean == "lower":
            return whitespace_clean(basic_clean(text)).lower()
        return canonicalize(basic_clean(text))
