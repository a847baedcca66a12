"""generate.py:21 imports QwenPromptExpander; the LLM prompt rewriter is out of scope."""


class QwenPromptExpander:
    def __init__(self, *a, **k):
        raise NotImplementedError("prompt extension (Qwen LLM) is outside the B200 hot path; run without --use_prompt_extend")
