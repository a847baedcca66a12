from .attention import flash_attention
from .model import WanModel
from .t5 import T5Decoder, T5Encoder, T5EncoderModel, T5Model
from .tokenizers import HuggingfaceTokenizer
from .vae import WanVAE

__all__ = ["WanVAE", "WanModel", "T5Model", "T5Encoder", "T5Decoder", "T5EncoderModel", "HuggingfaceTokenizer",
           "flash_attention"]
