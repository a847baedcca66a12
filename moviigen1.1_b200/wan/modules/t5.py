"""umT5-XXL text encoder — OUT OF SCOPE for this tier (runs twice per video, not per
step).  The names exist so `wan.modules` imports like the reference; constructing them raises.  WanT2V accepts
pre-computed text embeddings instead (see WanT2V.generate(context=..., context_null=...))."""


class _OutOfScope:
    def __init__(self, *a, **k):
        raise NotImplementedError(
            "%s: the umT5 text encoder is outside the B200 hot path built here; pass pre-computed text embeddings "
            "to WanT2V.generate(context=..., context_null=...)" % type(self).__name__)


class T5Model(_OutOfScope):
    pass


class T5Encoder(_OutOfScope):
    pass


class T5Decoder(_OutOfScope):
    pass


class T5EncoderModel(_OutOfScope):
    pass
